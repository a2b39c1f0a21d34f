"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the oracle."""
import numpy as np
import pytest

from fugu_b200 import _native as nat
from fugu_b200 import synth
from tests.util import check_batch_against_oracle, plan_queries

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = nat.Context(0)
    yield c
    c.close()


def _setup(ctx, cfg, n_docs=None, vocab=None):
    n_docs = n_docs or cfg.n_docs
    vocab = vocab or cfg.vocab
    corpus = synth.Corpus.for_config(cfg, vocab=vocab)
    fields = synth.build_fields(corpus, 0, n_docs)
    desc = nat.HostIndexDesc(n_docs, fields)
    return corpus, desc, nat.Index(ctx, desc)


def test_config1_two_term_and(ctx):
    """BASELINE config 0: 10k docs, 1k two-term AND queries, top-10."""
    cfg = synth.CONFIGS[1]
    corpus, desc, index = _setup(ctx, cfg)
    batch = plan_queries(synth.gen_queries(cfg), vocab=cfg.vocab, n_text_fields=1)
    check_batch_against_oracle(index, desc, batch)
    index.close()


def test_config2_small_mixed(ctx):
    """Config 2 shape at 50k docs: mixed 1-4 term AND/OR, two-field expansion (text + name)."""
    cfg = synth.Config(cfg=2, n_docs=50_000, vocab=20_000, n_queries=500, k=10, name_pct=10)
    corpus, desc, index = _setup(ctx, cfg)
    batch = plan_queries(synth.gen_queries(cfg), vocab=cfg.vocab, n_text_fields=2)
    check_batch_against_oracle(index, desc, batch)
    index.close()


def test_config3_small_or_top100(ctx):
    """Config 3 shape at 100k docs: stop-word-heavy OR of 2-6 terms, top-100 (KS=4 queue)."""
    cfg = synth.Config(cfg=3, n_docs=100_000, vocab=20_000, n_queries=100, k=100)
    corpus, desc, index = _setup(ctx, cfg)
    batch = plan_queries(synth.gen_queries(cfg), vocab=cfg.vocab, n_text_fields=1)
    check_batch_against_oracle(index, desc, batch)
    index.close()


def test_golden_cases_through_dataset_search(ctx):
    """The reference-facing path: ObjectRecords -> Dataset.upsert/delete/commit -> Dataset.search
    (query string, filters, page, per_page) on the GPU == committed golden hits (ids + scores)."""
    from tests.util import check_topk, golden_dataset

    g, ds, ix = golden_dataset(ctx)
    ds.commit()
    n_ok = 0
    for c in g["cases"]:
        try:
            res = ds.search(c["query"], c["filters"], c["page"], c["per_page"])
        except nat.FgError as e:
            assert c.get("error") == {nat.FG_ERR_INVALID: "invalid", nat.FG_ERR_UNSUPPORTED: "unsupported"}[e.code], c
            continue
        assert "error" not in c, c
        got = np.zeros(len(res), nat.HIT_DT)
        got["doc"] = [r.doc for r in res]
        got["score"] = [r.score for r in res]
        want = np.zeros(len(c["hits"]), nat.HIT_DT)
        want["doc"] = [ix.ids.index(i) if i not in ix.by_id else ix.by_id[i] for i, _ in c["hits"]]
        want["score"] = [s for _, s in c["hits"]]
        # a page is a slice of the top-(page+1)*per_page list: ties may be cut at either end
        assert len(got) == len(want), c
        for a, b in zip(got["score"], want["score"]):
            assert abs(a - b) <= 1e-5 * max(abs(a), abs(b), 1e-30), c
        if c["page"] == 0:
            check_topk(got, want, k=c["per_page"], ctx=repr(c["query"]))
            assert [r.id for r in res] == [ix.ids[d] for d in got["doc"]]
        n_ok += 1
    assert n_ok >= 50
    # batched form agrees with the one-at-a-time form
    qs = [c["query"] for c in g["cases"] if "error" not in c and c["page"] == 0]
    fl = [c["filters"] for c in g["cases"] if "error" not in c and c["page"] == 0]
    hits, nh, cnt, status = ds.search_batch(qs, fl, 0, 10)
    assert (status == 0).all()
    want_cnt = [c["match_count"] for c in g["cases"] if "error" not in c and c["page"] == 0]
    assert cnt.tolist() == want_cnt
    ds.close()


def test_upsert_delete_snapshot_semantics(ctx):
    """Upsert of an existing id deletes the old doc (alive bit) but statistics keep counting it
    (A.4), as the python twin models; a new commit swaps the snapshot."""
    from fugu_b200.dataset import Dataset, ObjectRecord
    from oracle import oracle_py as op

    ds, ix = Dataset(ctx), op.PyIndex()
    recs = [("a", "one two three"), ("b", "two three four four"), ("c", "five six"), ("a", "two two two seven")]
    for i, t in recs:
        ds.upsert([ObjectRecord(id=i, text=t)], commit=False)
        ix.upsert(i, t)
    ds.commit()
    for q in ["two", "three", "seven", "one", ""]:
        want, n = op.search(ix, q, [], 0, 10)
        got = ds.search(q, [], 0, 10)
        assert [r.id for r in got] == [ix.ids[d] for d, _ in want], q
        for r, (_, s) in zip(got, want):
            assert abs(r.score - s) <= 1e-5 * max(abs(s), 1e-30)
    ds.delete("b")
    ix.delete("b")
    want, n = op.search(ix, "three", [], 0, 10)
    assert [r.id for r in ds.search("three")] == [ix.ids[d] for d, _ in want] == []
    ds.close()


def test_sharded_search_and_device_merge(ctx):
    """Doc-id-range shards with global statistics + fg_merge_topk_device == single index (the
    multi-GPU data path of SURVEY.md 8(e), emulated on one device as two shard snapshots)."""
    from oracle import orc
    from tests.util import DevBuf, check_topk, gpu_search_device

    cfg = synth.Config(cfg=2, n_docs=40_000, vocab=8_000, n_queries=200, k=10, name_pct=10)
    corpus = synth.Corpus.for_config(cfg)
    whole = nat.HostIndexDesc(cfg.n_docs, synth.build_fields(corpus, 0, cfg.n_docs))
    batch = plan_queries(synth.gen_queries(cfg), vocab=cfg.vocab, n_text_fields=2)
    o_hits, o_n, o_c = orc.search(whole, batch, threads=4)
    R = 3
    bounds = [cfg.n_docs * r // R for r in range(R + 1)]
    shard_fields = [synth.build_fields(corpus, bounds[r], bounds[r + 1]) for r in range(R)]
    for f in range(2):
        gdf = sum(np.diff(sf[f]["term_offsets"]).astype(np.int64) for sf in shard_fields).astype(np.uint32)
        tot = sum(sf[f]["total_num_tokens"] for sf in shard_fields)
        for sf in shard_fields:
            sf[f]["global_doc_freq"] = gdf
            sf[f]["total_num_tokens"] = tot
    nq, k = batch.n_queries, batch.kmax
    g_hits = np.zeros((R, nq, k), nat.HIT_DT)
    g_n = np.zeros((R, nq), np.uint32)
    counts = np.zeros(nq, np.int64)
    for r in range(R):
        desc = nat.HostIndexDesc(bounds[r + 1] - bounds[r], shard_fields[r], doc_id_base=bounds[r], global_n_docs=cfg.n_docs)
        index = nat.Index(ctx, desc)
        h, n, c, _, _ = gpu_search_device(index, batch)
        g_hits[r], g_n[r] = h, n
        counts += c
        index.close()
    d_g = DevBuf(None, g_hits.view(np.int32).reshape(R, nq, k, 2))
    d_gn = DevBuf(None, g_n.view(np.int32))
    d_out, d_on = DevBuf((nq, k, 2)), DevBuf(nq)
    nat.merge_topk_device(ctx, d_g.ptr, d_gn.ptr, R, nq, k, k, d_out.ptr, d_on.ptr)
    ctx.synchronize()
    raw = d_out.numpy().view(np.uint32)
    got = np.zeros((nq, k), nat.HIT_DT)
    got["score"] = raw[..., 0].view(np.float32)
    got["doc"] = raw[..., 1]
    on = d_on.numpy()
    assert counts.tolist() == o_c.astype(np.int64).tolist()
    assert on.tolist() == o_n.tolist()
    for q in range(nq):
        check_topk(got[q, :on[q]], o_hits[q, :o_n[q]], k=k, ctx=f"query {q}")


def test_accounting_matches_oracle_definition(ctx):
    """fg_batch_stats (exact accounting mode) == the oracle's independent algorithmic-byte count."""
    from oracle import orc
    from tests.util import gpu_search_device

    cfg = synth.Config(cfg=2, n_docs=60_000, vocab=10_000, n_queries=300, k=10, name_pct=10)
    corpus, desc, index = _setup(ctx, cfg)
    batch = plan_queries(synth.gen_queries(cfg), vocab=cfg.vocab, n_text_fields=2)
    by, sc = orc.algorithmic_bytes(desc, batch, threads=4)
    *_, st = gpu_search_device(index, batch, flags=nat.FG_EXEC_EXACT_ACCOUNTING)
    assert st.scored_postings == int(sc.sum())
    # blocks straddling work-item boundaries are only counted by the item that owns their start;
    # filter-clause blocks first needed in a later round count as re-decodes -> a lower bound within 3%
    assert st.bytes_blocks <= int(by.sum())
    assert st.bytes_blocks >= 0.97 * int(by.sum()), (st.bytes_blocks, int(by.sum()))
    index.close()


def test_edge_cases(ctx):
    """Empty / ragged inputs: empty batch, empty posting lists, tail blocks (n % 128 != 0), k larger
    than the match count, missing terms, a term present in every doc, boosts, MustNot."""
    from oracle import orc
    from tests.util import check_batch_against_oracle

    rng = np.random.default_rng(11)
    n_docs = 5000
    lists = [np.arange(n_docs, dtype=np.uint32),                                   # every doc (gaps of 0 bits)
             np.sort(rng.choice(n_docs, 1, replace=False)).astype(np.uint32),     # single posting
             np.sort(rng.choice(n_docs, 127, replace=False)).astype(np.uint32),   # just under one block
             np.sort(rng.choice(n_docs, 128, replace=False)).astype(np.uint32),   # exactly one block
             np.sort(rng.choice(n_docs, 129, replace=False)).astype(np.uint32),   # one block + 1
             np.zeros(0, np.uint32),                                               # empty list
             np.array([0, n_docs - 1], np.uint32),                                 # first and last doc
             np.sort(rng.choice(n_docs, 2500, replace=False)).astype(np.uint32)]
    offs = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.uint64)
    docs = np.concatenate(lists)
    tfs = rng.integers(1, 300, len(docs)).astype(np.uint32)
    fn = rng.integers(0, 120, n_docs).astype(np.uint8)
    desc = nat.HostIndexDesc(n_docs, [{"term_offsets": offs, "doc_ids": docs, "term_freqs": tfs, "fieldnorm_ids": fn,
                                       "total_num_tokens": int(n_docs * 50)}])
    index = nat.Index(ctx, desc)
    S, M, N = nat.FG_OCCUR_SHOULD, nat.FG_OCCUR_MUST, nat.FG_OCCUR_MUST_NOT
    T = lambda t, b=1.0: (0, t, b)
    qs = [{"k": 10, "clauses": [(S, [T(t)])]} for t in range(8)]
    qs += [{"k": 128, "clauses": [(S, [T(4)])]}, {"k": 100, "clauses": [(S, [T(1)]), (S, [T(6)])]},
           {"k": 10, "clauses": [(M, [T(0)]), (M, [T(7)]), (M, [T(4)])]},
           {"k": 10, "clauses": [(M, [T(0)]), (N, [T(7)])]},
           {"k": 10, "clauses": [(S, [T(2), T(3), T(4)]), (N, [T(0)])]},
           {"k": 10, "clauses": [(M, [T(5)]), (S, [T(0)])]},
           {"k": 10, "clauses": [(S, [T(nat.FG_TERM_MISSING)])]},
           {"k": 10, "clauses": [(M, [T(7, 2.0)]), (S, [T(3, 0.5)]), (S, [T(6, 3.0)])]},
           {"k": 10, "clauses": []},
           {"k": 5, "clauses": [(M, [T(nat.FG_TERM_ALL)])]},
           {"k": 10, "clauses": [(M, [T(nat.FG_TERM_ALL)]), (M, [T(3)])]}]
    batch = nat.HostBatch(qs[:-2] + qs[-1:])
    check_batch_against_oracle(index, desc, batch)
    # pure AllQuery: first k alive docs, score 1.0 (the oracle leaves this to the host layer)
    h, n, c = index.search(nat.HostBatch([qs[-2]]))
    assert n[0] == 5 and c[0] == n_docs and h[0]["doc"].tolist() == [0, 1, 2, 3, 4] and (h[0]["score"] == 1.0).all()
    # empty batch
    h, n, c = index.search(nat.HostBatch([]))
    assert len(n) == 0
    # k == 0 is an error (TopDocs::with_limit asserts); k > 1024 is a deep page (test_deep_pagination_beyond_1024),
    # except on the window kernels, which keep the 1024 limit
    with pytest.raises(nat.FgError) as e:
        index.search(nat.HostBatch([{"k": 0, "clauses": [(S, [T(0)])]}]))
    assert e.value.code == nat.FG_ERR_INVALID
    h, n, c = index.search(nat.HostBatch([{"k": 1025, "clauses": [(S, [T(0)])]}]), want_counts=False)
    assert 0 < n[0] <= 1025
    with pytest.raises(nat.FgError) as e:
        index.prepare(nat.HostBatch([{"k": 1025, "clauses": [(S, [T(0)])]}]), nat.FG_PREP_LEGACY)
    assert e.value.code == nat.FG_ERR_UNSUPPORTED
    index.close()


def test_config5_facet_filters(ctx):
    """Config 5 shape at 60k docs: 2-term OR text + 1-2 facet filters (Must group, OR inside) over
    namespace/organization/data facets with ancestor terms (src/object.rs:81-111), planned by the
    C++ host layer from query + filter strings."""
    from fugu_b200.dataset import Dataset
    from oracle import orc
    from tests.util import check_topk, gpu_search_device

    cfg = synth.Config(cfg=5, n_docs=60_000, vocab=5_000, n_queries=300, k=10, n_ns=64)
    corpus = synth.Corpus.for_config(cfg)
    fields = synth.build_fields(corpus, 0, cfg.n_docs)
    assert len(fields) == 3
    desc = nat.HostIndexDesc(cfg.n_docs, fields)
    ds = Dataset(ctx)
    words = [f"w{i + 1}" for i in range(cfg.vocab)]
    ds.adopt(desc, [words, [], [corpus.facet_path(i) for i in range(corpus.facet_vocab())]])
    qs = synth.gen_queries(cfg)
    batch, status = ds.plan_batch([q["query"] for q in qs], [q["filters"] for q in qs], 0, 10)
    assert (status == 0).all()
    o_hits, o_n, o_c = orc.search(desc, batch, threads=4)
    g_hits, g_n, g_c, _, _ = gpu_search_device(ds.index(), batch)
    assert np.array_equal(g_c, o_c) and np.array_equal(g_n, o_n)
    assert (o_c > 0).sum() > 100  # the filters actually select documents
    for qi in range(batch.n_queries):
        check_topk(g_hits[qi, :g_n[qi]], o_hits[qi, :o_n[qi]], k=10, ctx=f"query {qi}: {qs[qi]}")
    # and through the string API, page 1
    h, nh, cnt, st = ds.search_batch([q["query"] for q in qs[:50]], [q["filters"] for q in qs[:50]], 1, 5)
    assert (st == 0).all() and np.array_equal(cnt, o_c[:50])
    ds.close()


def test_config4_three_term_and_with_deletes(ctx):
    """Config 4 shape (3-term AND) at 150k docs with 5% of the docs deleted (alive bitset)."""
    from tests.util import check_batch_against_oracle

    cfg = synth.Config(cfg=4, n_docs=150_000, vocab=20_000, n_queries=300, k=10)
    corpus = synth.Corpus.for_config(cfg)
    fields = synth.build_fields(corpus, 0, cfg.n_docs)
    rng = np.random.default_rng(3)
    alive = np.full((cfg.n_docs + 31) // 32, 0xFFFFFFFF, np.uint32)
    for d in rng.choice(cfg.n_docs, cfg.n_docs // 20, replace=False):
        alive[d >> 5] &= ~np.uint32(1 << (d & 31))
    desc = nat.HostIndexDesc(cfg.n_docs, fields, alive_bitset=alive)
    index = nat.Index(ctx, desc)
    batch = plan_queries(synth.gen_queries(cfg), vocab=cfg.vocab, n_text_fields=1)
    check_batch_against_oracle(index, desc, batch)
    index.close()


def test_deep_pagination_k_up_to_1024(ctx):
    """limit = page*per_page + per_page beyond 128 (32-row register queue): k = 300 and k = 1000."""
    cfg = synth.Config(cfg=2, n_docs=30_000, vocab=5_000, n_queries=40, k=10, name_pct=10)
    corpus, desc, index = _setup(ctx, cfg)
    qs = synth.gen_queries(cfg)
    for i, q in enumerate(qs):
        q["k"] = 300 if i % 2 else 1000
    batch = plan_queries(qs, vocab=cfg.vocab, n_text_fields=2)
    check_batch_against_oracle(index, desc, batch, bitmaps=False)
    index.close()


def test_deep_pagination_beyond_1024(ctx):
    """The reference puts no bound on page * per_page (src/db/search.rs:154-160; handlers/search.rs:370-374 clamps only
    per_page): limits above 1024 run without per-warp queues (append + radix select + sort). k = 1500 / 5000 / 40000
    (more than there are matches for most queries), through the device ABI in a batch of their own, and through
    fgh_search_batch mixed with ordinary pages."""
    from tests import util

    small = util.EMULATED  # (CPU suite time)
    cfg = synth.Config(cfg=2, n_docs=6_000 if small else 30_000, vocab=1_000 if small else 5_000, n_queries=12 if small else 24, k=10, name_pct=10)
    corpus, desc, index = _setup(ctx, cfg)
    qs = synth.gen_queries(cfg)
    for i, q in enumerate(qs):
        q["k"] = (1500, 5000, 40_000)[i % 3]
    batch = plan_queries(qs, vocab=cfg.vocab, n_text_fields=2)
    check_batch_against_oracle(index, desc, batch, bitmaps=False, legacy=False)
    index.close()
    # host layer: page 30 of 100 per page (limit 3100) next to first pages, one request
    from fugu_b200.dataset import Dataset, QuerySet

    ds = Dataset(ctx)
    words = [f"w{i + 1}" for i in range(cfg.vocab)]
    ds.adopt(desc, [words, words])
    strings = [q["query"] for q in qs[:12]]
    pages = [30 if i % 4 == 0 else 0 for i in range(12)]
    deep = QuerySet(strings, None, 0, 100)
    deep.pages[:] = np.array(pages, np.uint32)
    hits, nh, cnt, status = ds.search_batch(deep, want_counts=False)
    assert (status == 0).all()
    for i, s_ in enumerate(strings):
        one = ds.search_batch(QuerySet([s_], None, pages[i], 100), want_counts=False)
        assert nh[i] == one[1][0] and np.array_equal(hits[i, :nh[i]], one[0][0, :nh[i]]), (i, s_)
        # the page is rows [offset, offset + per_page) of the full ranking
        allr = ds.search_batch(QuerySet([s_], None, 0, 100 * (pages[i] + 1)), want_counts=False)
        want = allr[0][0, 100 * pages[i]:allr[1][0]]
        assert nh[i] == len(want) and np.array_equal(hits[i, :nh[i]]["doc"], want["doc"]), (i, s_)
    ds.close()


def test_concurrent_callers_share_one_index(ctx):
    """The ABI is thread-safe and re-entrant (axum handlers call Dataset::search concurrently,
    src/server/server_main.rs:50): 8 host threads hammer one snapshot; every result equals the
    single-threaded answer."""
    import threading

    from fugu_b200.dataset import Dataset

    cfg = synth.Config(cfg=2, n_docs=40_000, vocab=8_000, n_queries=240, k=10, name_pct=10)
    corpus = synth.Corpus.for_config(cfg)
    desc = nat.HostIndexDesc(cfg.n_docs, synth.build_fields(corpus, 0, cfg.n_docs))
    ds = Dataset(ctx)
    words = [f"w{i + 1}" for i in range(cfg.vocab)]
    ds.adopt(desc, [words, words])
    qs = [q["query"] for q in synth.gen_queries(cfg)]
    want = ds.search_batch(qs, None, 0, 10)
    errors = []

    def worker(t):
        try:
            for rep in range(3):
                part = qs[t::8]
                h, n, c, st = ds.search_batch(part, None, 0, 10)
                assert (st == 0).all()
                assert np.array_equal(c, want[2][t::8]) and np.array_equal(n, want[1][t::8])
                for i in range(len(part)):
                    check_topk_local(h[i, :n[i]], want[0][t::8][i, :n[i]])
                # single-request form as well
                r = ds.search(part[0], [], 0, 10)
                assert len(r) == int(n[0])
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    from tests.util import check_topk

    def check_topk_local(a, b):
        check_topk(a, b, k=10, ctx="concurrent")

    th = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errors, errors[:3]
    ds.close()


def test_many_leaf_union_in_hash_mode(ctx):
    """8-word OR over two fields = 16 sparse insert leaves (more leaves than hash-round block quota)."""
    cfg = synth.Config(cfg=2, n_docs=60_000, vocab=30_000, n_queries=1, k=10, name_pct=30)
    corpus, desc, index = _setup(ctx, cfg)
    rng = np.random.default_rng(9)
    qs = [{"query": " ".join(f"w{int(r)}" for r in rng.choice(np.arange(200, 3000), 8, replace=False)), "filters": [], "k": 10}
          for _ in range(40)]
    qs += [{"query": " ".join(f"w{int(r)}" for r in rng.choice(np.arange(1, 40), 8, replace=False)), "filters": [], "k": 10}
           for _ in range(10)]
    batch = plan_queries(qs, vocab=cfg.vocab, n_text_fields=2)
    check_batch_against_oracle(index, desc, batch)
    index.close()


def test_dense_tf_columns_every_plan_shape(ctx):
    """Terms in >= 1/16 of the docs get a dense tf column (1 B/doc) and are applied in the slot scan
    instead of block phases. Every role a column leaf can play, mixed with block leaves, text columns
    (fieldnorms) next to facet columns (constant norm), tf = 255, boosts, deletes -- against the oracle
    and against the block path (FG_PREP_NO_COLUMNS)."""
    rng = np.random.default_rng(5)
    n_docs = 40_003  # ragged: the last window ends off a multiple of 4
    def lst(p):
        return np.sort(rng.choice(n_docs, int(n_docs * p), replace=False)).astype(np.uint32)
    text_lists = [lst(0.9), lst(0.5), lst(0.2), lst(0.07), lst(0.03), lst(0.004), lst(0.0005),
                  np.arange(n_docs, dtype=np.uint32)]
    facet_lists = [lst(0.25), lst(0.5), lst(0.01)]
    def field(lists, freqs, norms):
        offs = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.uint64)
        docs = np.concatenate(lists)
        fd = {"term_offsets": offs, "doc_ids": docs, "total_num_tokens": int(n_docs * 40) if norms else 0,
              "term_freqs": None, "fieldnorm_ids": None}
        if freqs:
            tf = rng.integers(1, 9, len(docs)).astype(np.uint32)
            tf[rng.integers(0, len(docs), 50)] = 255
            fd["term_freqs"] = tf
        if norms:
            fd["fieldnorm_ids"] = rng.integers(1, 90, n_docs).astype(np.uint8)
        return fd
    alive = np.full((n_docs + 31) // 32, 0xFFFFFFFF, np.uint32)
    for d in rng.choice(n_docs, 500, replace=False):
        alive[d >> 5] &= ~np.uint32(1 << (d & 31))
    fields = [field(text_lists, True, True), field(facet_lists, False, False)]
    S, M, N = nat.FG_OCCUR_SHOULD, nat.FG_OCCUR_MUST, nat.FG_OCCUR_MUST_NOT
    T = lambda t, b=1.0: (0, t, b)
    F = lambda t, b=1.0: (1, t, b)
    shapes = [
        [(S, [T(0)])], [(S, [T(7)])], [(S, [T(0)]), (S, [T(1)]), (S, [T(2)])],
        [(S, [T(0), T(5)]), (S, [T(3), T(6)])],                       # column + block leaves in one clause
        [(S, [T(4)]), (S, [T(5)])],                                  # no column at all
        [(M, [T(0)])], [(M, [T(0)]), (M, [T(1)])], [(M, [T(1)]), (M, [T(2)]), (M, [T(0)])],
        [(M, [T(5)]), (M, [T(0)])], [(M, [T(0)]), (M, [T(5)])],      # block lead, column filter
        [(M, [T(4)]), (M, [T(1), T(6)]), (M, [T(0)])],
        [(M, [T(0), T(5)]), (M, [T(1), T(4)])],                      # mixed clauses: column lead, NOFILT block filter
        [(M, [T(2)]), (M, [T(4)]), (M, [T(3)])],                     # column lead, block clause, column clause
        [(M, [T(1)]), (S, [T(4)])], [(M, [T(4)]), (S, [T(0)])], [(M, [T(0)]), (S, [T(1)]), (S, [T(5)])],
        [(S, [T(0)]), (N, [T(4)])], [(S, [T(4)]), (N, [T(0)])], [(S, [T(1)]), (N, [T(2)])],
        [(M, [T(1)]), (N, [T(0)])], [(M, [T(5)]), (N, [T(1)]), (S, [T(2)])], [(M, [T(1)]), (M, [T(4)]), (N, [T(3)])],
        [(S, [T(0, 2.5)]), (S, [T(3, 0.5)])], [(S, [T(0, -1.0)]), (S, [T(4)])],   # negative boost: masked union
        [(M, [T(4)]), (M, [F(0), F(2)])], [(S, [T(1)]), (S, [T(5)]), (M, [F(1)])],  # facet Must group (constant norm)
        [(M, [F(0), F(1)])], [(M, [T(0)]), (M, [F(1)]), (N, [F(0)])],
        [(M, [T(nat.FG_TERM_ALL)]), (M, [T(0)])],
    ]
    qs = [{"k": k, "clauses": sh} for sh in shapes for k in (10, 100)]
    batch = nat.HostBatch(qs)
    for al in (None, alive):
        desc = nat.HostIndexDesc(n_docs, fields, alive_bitset=al)
        index = nat.Index(ctx, desc)
        assert index.info().n_columns == 7  # text: 0.9 0.5 0.2 0.07 and the all-docs list; facets: 0.25 0.5
        check_batch_against_oracle(index, desc, batch)
        index.close()


def test_async_submit_collect_and_pipelined_requests(ctx):
    """fg_batch_submit / fg_batch_collect (asynchronous host-buffer form): several batches in flight at
    once return what the blocking call returns; fgh_search_batch gives the same answers whether it
    runs a request as one batch or pipelines it in chunks; counts are optional (TopDocs does not count)."""
    import os

    from fugu_b200.dataset import Dataset
    from tests.util import check_topk

    cfg = synth.Config(cfg=2, n_docs=60_000, vocab=10_000, n_queries=900, k=10, name_pct=10)
    corpus = synth.Corpus.for_config(cfg)
    desc = nat.HostIndexDesc(cfg.n_docs, synth.build_fields(corpus, 0, cfg.n_docs))
    ds = Dataset(ctx)
    words = [f"w{i + 1}" for i in range(cfg.vocab)]
    ds.adopt(desc, [words, words])
    index = ds.index()
    qs = synth.gen_queries(cfg)
    parts = [plan_queries(qs[i::3], vocab=cfg.vocab, n_text_fields=2) for i in range(3)]
    want = [index.search(b) for b in parts]
    pbs = [index.prepare(b) for b in parts]
    for i, pb in enumerate(pbs):
        pb.submit(want_counts=(i != 1))
    for i in (2, 0, 1):  # collected out of order
        h, n, c = pbs[i].collect()
        assert np.array_equal(n, want[i][1])
        assert (c is None) if i == 1 else np.array_equal(c, want[i][2])
        for q in range(parts[i].n_queries):
            check_topk(h[q, :n[q]], want[i][0][q, :n[q]], k=10, ctx=f"batch {i} query {q}")
        pbs[i].close()
    with pytest.raises(nat.FgError):  # counts were not requested at submit time
        pb = index.prepare(parts[0]); pb.submit(want_counts=False)
        try:
            nat.check(nat.lib().fg_batch_collect(pb.h, nat._ptr(np.zeros((parts[0].n_queries, 10), nat.HIT_DT)),
                                                 nat._ptr(np.zeros(parts[0].n_queries, np.uint32)),
                                                 nat._ptr(np.zeros(parts[0].n_queries, np.uint32))))
        finally:
            pb.close()
    strings = [q["query"] for q in qs]
    ref = None
    for chunks in ("1", "2", "5"):
        os.environ["FG_PIPELINE_CHUNKS"] = chunks
        try:
            got = ds.search_batch(strings, None, 0, 10)
            nocnt = ds.search_batch(strings, None, 0, 10, want_counts=False)
        finally:
            os.environ.pop("FG_PIPELINE_CHUNKS")
        assert (got[3] == 0).all() and nocnt[2] is None and np.array_equal(nocnt[1], got[1])
        if ref is None:
            ref = got
            continue
        assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])
        for q in range(len(strings)):
            check_topk(got[0][q, :got[1][q]], ref[0][q, :ref[1][q]], k=10, ctx=f"chunks {chunks} query {q}")
    ds.close()


def test_search_while_commits_land(ctx):
    """A search that runs while another thread upserts + commits must not fail (the planner resolved
    terms against a dictionary newer than the snapshot it executed on and the lowering rejected the ordinals).
    Terms a snapshot does not know are empty scorers; a long query fails alone, not its siblings."""
    import threading

    from fugu_b200.dataset import Dataset, ObjectRecord

    ds = Dataset(ctx)
    ds.upsert([ObjectRecord(id=f"d{i}", text=f"alpha beta common{i % 7} w{i}") for i in range(200)], commit=True)
    stop, errors = threading.Event(), []

    def searcher():
        try:
            while not stop.is_set():
                for q in ("alpha beta", "common3 OR fresh17", "fresh3 AND alpha", "w5 w6 w7"):
                    ds.search(q, [], 0, 10)
                h, n, c, st = ds.search_batch(["alpha", "fresh1 fresh2", "beta AND common1"], None, 0, 10, want_counts=False)
                assert (st == 0).all()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    th = [threading.Thread(target=searcher) for _ in range(3)]
    for x in th:
        x.start()
    for r in range(25):
        ds.upsert([ObjectRecord(id=f"n{r}_{i}", text=f"alpha fresh{r} fresh{i} newterm{r}x{i}") for i in range(20)], commit=True)
    stop.set()
    for x in th:
        x.join()
    assert not errors, errors[:3]
    # one query over the leaf limit fails alone
    long_q = " ".join(f"w{i}" for i in range(1, 40))  # 39 words x 2 fields > 32 live leaves
    h, n, c, st = ds.search_batch(["alpha beta", long_q, "beta"], None, 0, 10, want_counts=False)
    assert st[0] == 0 and st[2] == 0 and st[1] == nat.FG_ERR_UNSUPPORTED and n[0] > 0 and n[1] == 0 and n[2] > 0
    ds.close()


def test_bulk_copy_staged_variant_equals_default(ctx, monkeypatch):
    """FG_LEAD_TMA=1 selects the kernel variant that stages lead-block payloads in shared memory with 1-D bulk
    copies (cp.async.bulk + mbarrier, two slots per warp): same results as the default (plain loads)."""
    monkeypatch.setenv("FG_LEAD_TMA", "1")
    c2 = nat.Context(0)  # the switch is read when a context is created
    try:
        cfg = synth.Config(cfg=2, n_docs=30_000, vocab=6_000, n_queries=200, k=10, name_pct=10)
        corpus = synth.Corpus.for_config(cfg)
        desc = nat.HostIndexDesc(cfg.n_docs, synth.build_fields(corpus, 0, cfg.n_docs))
        index = nat.Index(c2, desc)
        batch = plan_queries(synth.gen_queries(cfg), vocab=cfg.vocab, n_text_fields=2)
        check_batch_against_oracle(index, desc, batch, legacy=False)
        index.close()
    finally:
        c2.close()


def test_multi_item_plans_forced_small_items(ctx, monkeypatch):
    """Work-item boundaries: with tiny item sizes every query is cut into many doc-id ranges (16-aligned
    cuts, blocks straddling two items, streamed leaves starting mid-list, per-item partial top-k lists and
    counts merged per query) -- the shape a 1 M-doc batch has, at a size the oracle checks in seconds."""
    for k_, v_ in (("FG_ITEM_BYTES", "8192"), ("FG_ITEM_BYTES_HASH", "4096"), ("FG_ITEM_BYTES_MASKED", "8192"),
                   ("FG_COL_COST_DIV", "1"), ("FG_COL_COST_DIV_PHASES", "1")):
        monkeypatch.setenv(k_, v_)
    cfg = synth.Config(cfg=2, n_docs=120_008, vocab=20_000, n_queries=400, k=10, name_pct=10)
    corpus, desc, index = _setup(ctx, cfg)
    batch = plan_queries(synth.gen_queries(cfg), vocab=cfg.vocab, n_text_fields=2)
    res = check_batch_against_oracle(index, desc, batch)
    assert res["stats"].n_work_items > batch.n_queries  # (heavy queries are cut into up to 14 / 29 items)
    index.close()


def test_gated_column_scan_equals_exhaustive(ctx):
    """Pruning never changes a result. Default lowering (lead-driven kernels): MaxScore / block-max pruning on
    (TopDocs form) vs off (FG_EXEC_NO_PRUNE) vs the oracle, the counters prove blocks were skipped. Legacy
    lowering (FG_PREP_LEGACY): sparse-hit gating of the column scan, same three-way check. Deletes included."""
    from oracle import orc
    from tests.util import DevBuf, check_topk

    cfg = synth.Config(cfg=2, n_docs=50_000, vocab=10_000, n_queries=1, k=10, name_pct=10)
    corpus = synth.Corpus.for_config(cfg)
    fields = synth.build_fields(corpus, 0, cfg.n_docs)
    rng = np.random.default_rng(21)
    alive = np.full((cfg.n_docs + 31) // 32, 0xFFFFFFFF, np.uint32)
    for d in rng.choice(cfg.n_docs, cfg.n_docs // 50, replace=False):
        alive[d >> 5] &= ~np.uint32(1 << (d & 31))
    qs = []
    for i in range(36):
        cols = rng.choice(np.arange(1, 40), 1 + i % 3, replace=False)      # frequent terms: dense tf columns
        sparse = rng.choice(np.arange(150, 4000), 1 + (i // 3) % 3, replace=False)  # streamed leaves, from a few per window to one in many windows
        words = [f"w{int(r)}" for r in list(cols) + list(sparse)]
        rng.shuffle(words)
        qs.append({"query": " ".join(words), "filters": [], "k": 10 if i % 4 else 100})
    batch = plan_queries(qs, vocab=cfg.vocab, n_text_fields=2)
    from tests import util as _u

    for bits in ((alive,) if (_u.EMULATED and not _u.FULL) else (None, alive)):  # (CPU suite time: one variant under emulation)
        desc = nat.HostIndexDesc(cfg.n_docs, fields, alive_bitset=bits)
        index = nat.Index(ctx, desc)
        assert index.info().n_columns >= 40
        o_hits, o_n, _ = orc.search(desc, batch, threads=4)
        nq, ks = batch.n_queries, batch.kmax
        for prep in (0, nat.FG_PREP_LEGACY):
            res = {}
            for name, flags in (("gated", nat.FG_EXEC_COUNTERS), ("exhaustive", nat.FG_EXEC_COUNTERS | nat.FG_EXEC_NO_PRUNE)):
                d_hits, d_n = DevBuf((nq, ks, 2)), DevBuf(nq)
                pb = index.prepare(batch, prep)
                pb.execute(d_hits.ptr, d_n.ptr, None, None, k_stride=ks, flags=flags)  # no match counts: the TopDocs form
                st = pb.stats()
                raw = d_hits.numpy().view(np.uint32).reshape(nq, ks, 2)
                out = np.zeros((nq, ks), nat.HIT_DT)
                out["score"], out["doc"] = raw[:, :, 0].view(np.float32), raw[:, :, 1]
                if prep & nat.FG_PREP_LEGACY:
                    res[name] = (out, d_n.numpy().view(np.uint32), st.colscan_chunks, st.colscan_chunks_skipped)
                else:  # lead blocks whose block maximum was tested / of which skipped
                    res[name] = (out, d_n.numpy().view(np.uint32), st.lead_blocks_seen, st.lead_blocks_seen - st.lead_blocks)
                pb.close()
            g, e = res["gated"], res["exhaustive"]
            assert e[3] == 0 and e[2] > 0
            if prep & nat.FG_PREP_LEGACY:
                assert g[2] == e[2]
            # how much is skipped depends on how fast the threshold warms up (work items of a query run one after
            # the other under emulation, concurrently on a GPU): only require that the pruning engaged
            assert 0 < g[3] <= g[2], f"pruning skipped {g[3]} of {g[2]}"
            assert np.array_equal(g[1], o_n) and np.array_equal(e[1], o_n)
            for qi in range(nq):
                n = int(o_n[qi])
                k = int(batch.q["k"][qi])
                check_topk(g[0][qi, :n], o_hits[qi, :n], k, ctx=f"pruned query {qi} {qs[qi]['query']}")
                check_topk(e[0][qi, :n], o_hits[qi, :n], k, ctx=f"exhaustive query {qi}")
                # same kernels, same operation order: pruned and exhaustive agree far inside the oracle tolerance
                check_topk(g[0][qi, :n], e[0][qi, :n], k, tol=1e-6, ctx=f"pruned vs exhaustive query {qi}")
        # with match counts every doc is visited and the counts are exact
        h_hits, h_n, h_c = index.search(batch)
        _, _, o_c = orc.search(desc, batch, threads=4)
        assert np.array_equal(h_c, o_c)
        index.close()


def test_index_append_equals_full_upload(ctx, monkeypatch):
    """Row f3: a snapshot grown by fg_index_append (base + two segments, with deletes in the last one) answers exactly
    like the oracle on the whole corpus -- every engine, lead exhaustive / pruned / block path / window kernels -- and
    only the segments' bytes cross PCIe. Small column / bitmap thresholds so that the base owns both kinds of lookup
    structures and their extension is exercised (terms keep the structure they had)."""
    monkeypatch.setenv("FG_BITMAP_MIN_DF", "64")
    cfg = synth.Config(cfg=2, n_docs=30_000, vocab=3_000, n_queries=240, k=10, name_pct=10)
    corpus = synth.Corpus.for_config(cfg)
    cuts = [0, 17_003, 24_130, cfg.n_docs]  # (not multiples of 128 / 256 / 32: partial blocks, chunks and bitset words)
    whole = nat.HostIndexDesc(cfg.n_docs, synth.build_fields(corpus, 0, cfg.n_docs))
    batch = plan_queries(synth.gen_queries(cfg), vocab=cfg.vocab, n_text_fields=2)
    base = nat.Index(ctx, nat.HostIndexDesc(cuts[1], synth.build_fields(corpus, 0, cuts[1])))
    assert base.info().n_columns > 0 and base.info().n_bitmaps > 0
    seg1 = nat.HostIndexDesc(cuts[2] - cuts[1], synth.build_fields(corpus, cuts[1], cuts[2]))
    mid = base.append(seg1)
    # the intermediate snapshot is a complete index of the first two parts
    part = nat.HostIndexDesc(cuts[2], synth.build_fields(corpus, 0, cuts[2]))
    check_batch_against_oracle(mid, part, batch, legacy=False)
    # second segment, with deletes (an upsert is delete + add, src/db/document.rs:38-48)
    rng = np.random.default_rng(5)
    alive = np.full((cfg.n_docs + 31) // 32, 0xFFFFFFFF, np.uint32)
    for d in rng.choice(cfg.n_docs, 900, replace=False):
        alive[d >> 5] &= ~np.uint32(1 << (d & 31))
    seg2 = nat.HostIndexDesc(cuts[3] - cuts[2], synth.build_fields(corpus, cuts[2], cuts[3]))
    full = mid.append(seg2, alive)
    whole_alive = nat.HostIndexDesc(cfg.n_docs, synth.build_fields(corpus, 0, cfg.n_docs), alive_bitset=alive)
    check_batch_against_oracle(full, whole_alive, batch)
    i_full, i_ref = full.info(), nat.Index(ctx, whole).info()
    # (only bitmap terms have their partial last block re-encoded; other terms get new blocks behind their old ones)
    assert i_full.n_postings == i_ref.n_postings and i_full.n_blocks >= i_ref.n_blocks and i_full.n_docs == cfg.n_docs
    # what crossed PCIe for the last segment (a fifth of the corpus: its blocks, re-encoded tail blocks, column bytes)
    # against what a full upload reads from the host (8 B per posting of the flat CSR)
    assert 0 < i_full.appended_bytes_h2d < 8 * i_ref.n_postings // 4
    # the base is untouched and still answers for its own docs
    check_batch_against_oracle(base, nat.HostIndexDesc(cuts[1], synth.build_fields(corpus, 0, cuts[1])), batch, legacy=False)
    for ix in (base, mid, full):
        ix.close()


def test_dataset_commits_append_segments(ctx):
    """fgh_dataset_commit hands new documents over as a segment (fg_index_append) instead of re-uploading the corpus;
    results equal those of a dataset built in one go, and the python twin's; a rebuild follows once the appended part
    outgrows the part uploaded whole."""
    from fugu_b200.dataset import Dataset, ObjectRecord
    from oracle import oracle_py as op

    words = ["alpha", "beta", "gamma", "delta", "omega", "sigma", "kappa", "theta"]
    def rec(i):
        text = " ".join(words[(i * 7 + j * j) % len(words)] for j in range(3 + i % 5)) + f" uniq{i}"
        return ObjectRecord(id=f"d{i}", text=text, metadata={"name": f"{words[i % 8]} report"}, namespace="ns%d" % (i % 3))
    inc, one, ix = Dataset(ctx), Dataset(ctx), op.PyIndex()
    n = 700
    for a, b in [(0, 400), (400, 520), (520, 521), (521, 700)]:
        inc.upsert([rec(i) for i in range(a, b)], commit=True)
    inc.upsert([rec(3)], commit=False)      # an upsert of an existing id: delete + append
    inc.delete("d10")
    inc.commit()
    one.upsert([rec(i) for i in range(n)] + [rec(3)], commit=False)
    one.delete("d10")
    one.commit()
    for i in list(range(n)) + [3]:
        r = rec(i)
        ix.upsert(r.id, r.text, name=r.metadata["name"], facets=[f"/namespace/{r.namespace}"])
    ix.delete("d10")
    assert inc.commit_counts() == (1, 4) and one.commit_counts() == (1, 0)
    for q, f in [("alpha", []), ("beta AND gamma", []), ("omega sigma report", []), ("uniq3", []), ("uniq10", []), ("kappa", ["namespace/ns1"]),
                 ("theta -alpha", []), ("", ["namespace/ns2"])]:
        a, b = inc.search(q, f, 0, 25), one.search(q, f, 0, 25)
        assert [r.id for r in a] == [r.id for r in b], (q, f)
        for x, y in zip(a, b):
            assert abs(x.score - y.score) <= 1e-6 * max(abs(y.score), 1e-30), (q, x, y)
        want, _ = op.search(ix, q, f, 0, 25)
        assert [r.id for r in a] == [ix.ids[d] for d, _ in want], (q, f)
    # doubling the corpus since the last full upload triggers a rebuild
    inc.upsert([rec(i) for i in range(n, 2 * n + 200)], commit=True)
    assert inc.commit_counts()[0] == 2
    inc.close(); one.close()


def test_union_of_boolean_queries(ctx):
    """fg_search_union_of: one query whose Should children are boolean queries themselves -- `(a AND b) OR (c AND d)`,
    `a OR (b AND c)`, a conjunction with an excluded term next to a plain union -- against the oracle's union of the
    children's scorers: same docs, scores within 1e-5, same match counts; pages of 10, 100 and 3000."""
    from oracle import orc
    from tests.util import check_topk

    cfg = synth.Config(cfg=2, n_docs=30_000, vocab=3_000, n_queries=8, k=10, name_pct=10)
    corpus, desc, index = _setup(ctx, cfg)
    S, M, N = nat.FG_OCCUR_SHOULD, nat.FG_OCCUR_MUST, nat.FG_OCCUR_MUST_NOT
    W = lambda t: [(0, t, 1.0), (1, t, 1.0)]  # a bare word over the default fields [text, name]
    cases = [
        [[(M, W(3)), (M, W(40))], [(M, W(7)), (M, W(120))]],                       # (a AND b) OR (c AND d)
        [[(S, W(900))], [(M, W(1)), (M, W(2)), (M, W(5))]],                         # a OR (b AND c AND d)
        [[(S, W(300)), (S, W(45))], [(M, W(10)), (N, W(0))], [(M, W(60)), (M, W(61))]],  # (a b) OR (c -d) OR (e AND f)
        [[(M, W(2500)), (M, W(2900))], [(M, W(2999)), (M, W(2998))]],               # rare terms: few or no matches
        [[(M, [(0, 0, 1.0)]), (M, [(0, 1, 2.0)])], [(S, [(0, 0, 0.5)])]],           # overlapping disjuncts, boosts
    ]
    for ci, disj in enumerate(cases):
        batch = nat.HostBatch([{"k": 1, "clauses": cl} for cl in disj])
        for k in (10, 100, 3000):
            o_hits, o_cnt = orc.search_union_of(desc, batch, k)
            g_hits, g_cnt = index.search_union_of(batch, k)
            assert g_cnt == o_cnt and len(g_hits) == len(o_hits), (ci, k, g_cnt, o_cnt, len(g_hits), len(o_hits))
            check_topk(g_hits, o_hits, k, ctx=f"union-of case {ci}, k = {k}")
    # filter children (fg_search_union_of_filtered): Bool[Must(union of the children), Must(filter)..] -- a document needs
    # an ordinary child and every filter, the filters score once. Against the oracle's Intersection(union, filters).
    fcases = [
        ([[(M, W(3)), (M, W(40))], [(M, W(7)), (M, W(120))], [(S, W(1))]], 1),                   # ((a AND b) OR (c AND d)) AND f
        ([[(S, W(900))], [(M, W(1)), (M, W(2))], [(M, W(0)), (M, [(0, 900, 0.0), (0, 1, 0.0), (0, 2, 0.0)])]], 1),  # the planner's form: f AND any-text^0
        ([[(M, W(10)), (N, W(0))], [(M, W(60)), (M, W(61))], [(S, W(5))], [(S, W(2))]], 2),       # two filters
        ([[(M, W(2500)), (M, W(2900))], [(M, W(3)), (M, W(4))], [(S, W(2999))]], 1),              # rare filter: few or no matches
    ]
    for ci, (disj, nf) in enumerate(fcases):
        batch = nat.HostBatch([{"k": 1, "clauses": cl} for cl in disj])
        for k in (10, 100, 3000):
            o_hits, o_cnt = orc.search_union_of(desc, batch, k, n_filters=nf)
            g_hits, g_cnt = index.search_union_of(batch, k, n_filters=nf)
            assert g_cnt == o_cnt and len(g_hits) == len(o_hits), (ci, k, g_cnt, o_cnt, len(g_hits), len(o_hits))
            check_topk(g_hits, o_hits, k, ctx=f"filtered union-of case {ci}, k = {k}")
    # with deleted documents (an alive bitset on the same postings)
    rng = np.random.default_rng(11)
    alive = np.full((cfg.n_docs + 31) // 32, 0xFFFFFFFF, np.uint32)
    for d in rng.choice(cfg.n_docs, 3000, replace=False):
        alive[d >> 5] &= ~np.uint32(1 << (d & 31))
    dead_ix = index.with_alive(alive)
    dead_desc = nat.HostIndexDesc(cfg.n_docs, synth.build_fields(corpus, 0, cfg.n_docs), alive_bitset=alive)
    for ci, disj in enumerate(cases[:3]):
        batch = nat.HostBatch([{"k": 1, "clauses": cl} for cl in disj])
        o_hits, o_cnt = orc.search_union_of(dead_desc, batch, 50)
        g_hits, g_cnt = dead_ix.search_union_of(batch, 50)
        assert g_cnt == o_cnt and len(g_hits) == len(o_hits), (ci, g_cnt, o_cnt)
        check_topk(g_hits, o_hits, 50, ctx=f"union-of with deletes, case {ci}")
    dead_ix.close()
    index.close()


def test_grammar_replay_kit_through_dataset_search(ctx):
    """The committed grammar replay kit (tests/golden/grammar_replay: what a box with `cargo` would POST to a real fugu)
    through Dataset.upsert / delete / commit / search on the device (and, in the CPU suite, on the emulated library): every case the device path answers gives
    the kit's page; a case the kit marks as an error (HTTP 500 in the reference) is FG_ERR_INVALID; the rest may only be
    FG_ERR_UNSUPPORTED."""
    import json
    import os

    from fugu_b200.dataset import Dataset, ObjectRecord

    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grammar_replay")
    ds = Dataset(ctx)
    ds.upsert([ObjectRecord(id=d["id"], text=d["text"], metadata=d["metadata"], facets=d["facets"])
               for d in json.load(open(os.path.join(here, "ingest.json")))["data"]], commit=True)
    for i in json.load(open(os.path.join(here, "deletes.json"))):
        ds.delete(i, commit=False)
    ds.commit()
    answered = unsupported = 0
    for ln in open(os.path.join(here, "cases.jsonl")):
        c = json.loads(ln)
        b = c["body"]
        try:
            res = ds.search(b["query"], b["filters"], b["page"]["page"], b["page"]["per_page"])
        except nat.FgError as e:
            if e.code == nat.FG_ERR_UNSUPPORTED:
                unsupported += 1
                continue
            assert e.code == nat.FG_ERR_INVALID and "error" in c, (b, str(e))
            continue
        assert "error" not in c, b
        want = c["hits"]
        assert len(res) == len(want), b
        for j, (r, (wid, ws)) in enumerate(zip(res, want)):
            assert abs(r.score - ws) <= 1e-5 * max(abs(ws), 1e-30), (b, j, r.score, ws)
            if r.id != wid:  # only inside a run of tied scores (or a tie cut by the end of the page)
                tie = [x for x in range(len(want)) if abs(want[x][1] - ws) <= 4e-5 * abs(ws)]
                assert r.id in [want[x][0] for x in tie] or max(tie) == len(want) - 1, (b, j, r.id, wid)
        answered += 1
    assert answered >= 250 and answered + unsupported >= 355, (answered, unsupported)
    ds.close()
