"""Dev tool (GPU): snapshot refresh cost at C2 scale. Base = the first 900 k docs of the 1 M-doc C2 corpus, then the last
100 k docs arrive as one new segment: fg_index_append vs a full fg_index_upload of all 1 M docs, and the answers of the
appended snapshot against the fully uploaded one (TopDocs form, 5000 queries) and against the oracle (200 queries)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fugu_b200 import _native as nat  # noqa: E402
from fugu_b200 import synth  # noqa: E402
from tests.util import check_topk, plan_queries  # noqa: E402

cfg = synth.Config(cfg=2, n_docs=1_000_000, vocab=200_000, n_queries=5000, k=10, name_pct=10)
corpus = synth.Corpus.for_config(cfg)
n0 = 900_000
whole = nat.HostIndexDesc(cfg.n_docs, synth.build_fields(corpus, 0, cfg.n_docs))
base_d = nat.HostIndexDesc(n0, synth.build_fields(corpus, 0, n0))
seg_d = nat.HostIndexDesc(cfg.n_docs - n0, synth.build_fields(corpus, n0, cfg.n_docs))
ctx = nat.Context(0)
t = time.perf_counter(); full = nat.Index(ctx, whole); t_full = time.perf_counter() - t
base = nat.Index(ctx, base_d)
for rep in range(2):
    t = time.perf_counter(); app = base.append(seg_d); t_app = time.perf_counter() - t
    if rep == 0:
        app.close()
i = app.info()
print(f"full upload of 1 M docs: {t_full * 1e3:.0f} ms; append of 100 k docs to a 900 k-doc snapshot: {t_app * 1e3:.0f} ms "
      f"({i.appended_bytes_h2d / 1e6:.1f} MB over PCIe vs {8 * full.info().n_postings / 1e6:.0f} MB of CSR for the full upload)")
batch = plan_queries(synth.gen_queries(cfg), vocab=cfg.vocab, n_text_fields=2)
h0, n0_, _ = full.search(batch, want_counts=False)
h1, n1_, _ = app.search(batch, want_counts=False)
assert np.array_equal(n0_, n1_)
bad = 0
for q in range(batch.n_queries):
    n = int(n0_[q])
    try:
        check_topk(h1[q, :n], h0[q, :n], 10, ctx=f"query {q}")
    except AssertionError as e:
        bad += 1
        print(e)
print(f"appended snapshot vs fully uploaded snapshot: {batch.n_queries - bad} of {batch.n_queries} queries identical")
from oracle import orc  # noqa: E402  (checker)

sub = nat.HostBatch.from_arrays(batch.q[:200], batch.c, batch.l)
o_h, o_n, _ = orc.search(whole, sub, threads=16)
g_h, g_n, _ = app.search(sub, want_counts=False)
assert np.array_equal(o_n, g_n)
for q in range(200):
    check_topk(g_h[q, :int(o_n[q])], o_h[q, :int(o_n[q])], 10, ctx=f"oracle, query {q}")
print("appended snapshot vs oracle: 200 of 200 queries OK")
assert bad == 0
