"""GPU parity at the sizes BASELINE.json names (-m gpu): the benchmarked configurations, not scaled-down shapes.

C2 1 M docs x 5000 mixed queries (the metric's configuration), C3 10 M docs top-100 stop-word unions, C4's shape
(3-term AND) at 20 M docs over 4 doc-id-range shards merged on the device, C5 10 M docs with 64 namespaces and
facet filters. Each case checks the TopDocs form (no match counts: the lead-driven kernels with MaxScore /
block-max pruning, i.e. exactly what bench.py times) and the counting form (exhaustive) against the oracle:
scores within 1e-5 relative, same documents except inside ties, match counts bit-exact.
Sizes can be scaled down for a quick look with FG_FULLSIZE_SCALE=<divisor>."""
import os

import numpy as np
import pytest

from fugu_b200 import _native as nat
from fugu_b200 import synth
from tests.util import check_topk, plan_queries

pytestmark = pytest.mark.gpu
SCALE = max(1, int(os.environ.get("FG_FULLSIZE_SCALE", "1")))
THREADS = max(4, len(os.sched_getaffinity(0)))


@pytest.fixture(scope="module")
def ctx():
    c = nat.Context(0)
    yield c
    c.close()


def _check(index, desc, batch, o=None, label=""):
    from oracle import orc

    o_hits, o_n, o_c = o if o is not None else orc.search(desc, batch, threads=THREADS)
    t_hits, t_n, _ = index.search(batch, want_counts=False)   # TopDocs form: pruned
    c_hits, c_n, c_c = index.search(batch, want_counts=True)  # counting form: every matching doc visited
    assert np.array_equal(t_n, o_n), f"{label}: n_hits differ (TopDocs form) for queries {np.nonzero(t_n != o_n)[0][:10]}"
    assert np.array_equal(c_c, o_c), f"{label}: match counts differ for queries {np.nonzero(c_c != o_c)[0][:10]}"
    for qi in range(batch.n_queries):
        n, k = int(o_n[qi]), int(batch.q["k"][qi])
        check_topk(t_hits[qi, :n], o_hits[qi, :n], k, ctx=f"{label} TopDocs form, query {qi}")
        check_topk(c_hits[qi, :n], o_hits[qi, :n], k, ctx=f"{label} counting form, query {qi}")
    return o_hits, o_n, o_c


def test_c2_one_million_docs_5000_mixed_queries(ctx):
    base = synth.CONFIGS[2]
    cfg = synth.Config(cfg=2, n_docs=base.n_docs // SCALE, vocab=base.vocab, n_queries=base.n_queries, k=base.k, name_pct=base.name_pct)
    corpus = synth.Corpus.for_config(cfg)
    desc = nat.HostIndexDesc(cfg.n_docs, synth.build_fields(corpus, 0, cfg.n_docs))
    index = nat.Index(ctx, desc)
    batch = plan_queries(synth.gen_queries(cfg), vocab=cfg.vocab, n_text_fields=2)
    _check(index, desc, batch, label="C2")
    info = index.info()
    assert info.n_columns > 0 and info.n_bitmaps > 0
    index.close()


def test_c3_ten_million_docs_stopword_unions_top100(ctx):
    base = synth.CONFIGS[3]
    cfg = synth.Config(cfg=3, n_docs=base.n_docs // SCALE, vocab=base.vocab, n_queries=200, k=base.k)
    corpus = synth.Corpus.for_config(cfg)
    desc = nat.HostIndexDesc(cfg.n_docs, synth.build_fields(corpus, 0, cfg.n_docs))
    index = nat.Index(ctx, desc)
    batch = plan_queries(synth.gen_queries(cfg), vocab=cfg.vocab, n_text_fields=1)
    _check(index, desc, batch, label="C3")
    index.close()


def test_c4_shape_20m_docs_four_shards_device_merge(ctx):
    """3-term AND (C4's queries) over 4 doc-id-range shards with global statistics, per-shard top-k merged by
    fg_merge_topk_device == the oracle on the unsharded 20 M-doc corpus."""
    from oracle import orc
    from tests.util import DevBuf

    n_docs, R = 20_000_000 // SCALE, 4
    cfg = synth.Config(cfg=4, n_docs=n_docs, vocab=synth.CONFIGS[4].vocab, n_queries=200, k=10)
    corpus = synth.Corpus.for_config(cfg)
    batch = plan_queries(synth.gen_queries(cfg), vocab=cfg.vocab, n_text_fields=1)
    bounds = [n_docs * r // R for r in range(R + 1)]
    shard_fields = [synth.build_fields(corpus, bounds[r], bounds[r + 1]) for r in range(R)]
    gdf = sum(np.diff(sf[0]["term_offsets"]).astype(np.int64) for sf in shard_fields).astype(np.uint32)
    tot = sum(sf[0]["total_num_tokens"] for sf in shard_fields)
    nq, k = batch.n_queries, batch.kmax
    g_hits = np.zeros((R, nq, k), nat.HIT_DT)
    g_n = np.zeros((R, nq), np.uint32)
    counts = np.zeros(nq, np.int64)
    for r in range(R):
        shard_fields[r][0]["global_doc_freq"] = gdf
        shard_fields[r][0]["total_num_tokens"] = tot
        desc = nat.HostIndexDesc(bounds[r + 1] - bounds[r], shard_fields[r], doc_id_base=bounds[r], global_n_docs=n_docs)
        index = nat.Index(ctx, desc)
        g_hits[r], g_n[r], _ = index.search(batch, want_counts=False)  # the pruned form, per shard
        _, _, c = index.search(batch, want_counts=True)
        counts += c
        index.close()
    d_g = DevBuf(None, g_hits.view(np.int32).reshape(R, nq, k, 2))
    d_gn = DevBuf(None, g_n.view(np.int32))
    d_out, d_on = DevBuf((nq, k, 2)), DevBuf(nq)
    nat.merge_topk_device(ctx, d_g.ptr, d_gn.ptr, R, nq, k, k, d_out.ptr, d_on.ptr)
    ctx.synchronize()
    raw = d_out.numpy().view(np.uint32)
    got = np.zeros((nq, k), nat.HIT_DT)
    got["score"], got["doc"] = raw[..., 0].view(np.float32), raw[..., 1]
    on = d_on.numpy().view(np.uint32)
    del shard_fields
    whole = nat.HostIndexDesc(n_docs, synth.build_fields(corpus, 0, n_docs))
    o_hits, o_n, o_c = orc.search(whole, batch, threads=THREADS)
    assert counts.tolist() == o_c.astype(np.int64).tolist()
    assert on.tolist() == o_n.tolist()
    for q in range(nq):
        check_topk(got[q, :on[q]], o_hits[q, :o_n[q]], k=k, ctx=f"C4 shape, query {q}")


def test_c5_ten_million_docs_64_namespaces_facet_filters(ctx):
    from fugu_b200.dataset import Dataset

    base = synth.CONFIGS[5]
    cfg = synth.Config(cfg=5, n_docs=base.n_docs // SCALE, vocab=base.vocab, n_queries=500, k=base.k, n_ns=base.n_ns)
    corpus = synth.Corpus.for_config(cfg)
    fields = synth.build_fields(corpus, 0, cfg.n_docs)
    desc = nat.HostIndexDesc(cfg.n_docs, fields)
    ds = Dataset(ctx)
    words = [f"w{i + 1}" for i in range(cfg.vocab)]
    ds.adopt(desc, [words, [], [corpus.facet_path(i) for i in range(corpus.facet_vocab())]])
    qs = synth.gen_queries(cfg)
    batch, status = ds.plan_batch([q["query"] for q in qs], [q["filters"] for q in qs], 0, 10)
    assert (status == 0).all()
    o_hits, o_n, o_c = _check(ds.index(), desc, batch, label="C5")
    assert (o_c > 0).sum() > 100  # the filters actually select documents
    ds.close()
