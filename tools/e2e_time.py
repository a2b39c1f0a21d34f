import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fugu_b200 import _native as nat, synth
from fugu_b200.dataset import Dataset, QuerySet
from bench import term_lists
cfg = synth.Config(cfg=2, n_docs=1_000_000, vocab=200_000, n_queries=5000, k=10, name_pct=10)
corpus = synth.Corpus.for_config(cfg)
fields = synth.build_fields(corpus, 0, cfg.n_docs)
desc = nat.HostIndexDesc(cfg.n_docs, fields)
ctx = nat.Context(0)
ds = Dataset(ctx); ds.adopt(desc, term_lists(corpus, cfg, 2))
qs = synth.gen_queries(cfg)
qset = QuerySet([q["query"] for q in qs], None, 0, 10)
for i in range(5):
    t = time.perf_counter(); r = ds.search_batch(qset, want_counts=False); print("search_batch total ms", (time.perf_counter() - t) * 1e3, file=sys.stderr)
