"""Dev tool: throughput of the search kernel vs. list density, for synthetic single/multi-leaf queries."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fugu_b200 import _native as nat, synth
from fugu_b200.dataset import Dataset, QuerySet
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)) + "/..")
from bench import term_lists

def main():
    cfg = synth.Config(cfg=2, n_docs=1_000_000, vocab=200_000, n_queries=1, k=10, name_pct=10)
    corpus = synth.Corpus.for_config(cfg)
    fields = synth.build_fields(corpus, 0, cfg.n_docs)
    desc = nat.HostIndexDesc(cfg.n_docs, fields)
    ctx = nat.Context(0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    ds = Dataset(ctx); ds.adopt(desc, term_lists(corpus, cfg, 2)); index = ds.index()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    offs = fields[0]["term_offsets"]
    cases = json.loads(os.environ.get("MICRO", "null")) or [
        ("text:w1", 600), ("text:w10", 2000), ("text:w30", 4000), ("text:w100", 8000), ("text:w300", 20000), ("text:w3000", 20000),
        ("w1", 600), ("w300", 20000), ("text:w1 text:w2", 400), ("text:w1 text:w2 text:w3 text:w4", 300),
        ("text:w300 text:w301", 10000), ("text:w300 text:w301 text:w302 text:w303", 8000), ("w300 w301 w302 w303", 8000),
        ("text:w1 AND text:w2", 400), ("text:w300 AND text:w301", 10000), ("text:w3000 AND text:w1", 10000)]
    for q, rep in cases:
        qs = QuerySet([q] * rep, None, 0, 10)
        b, st = ds.plan_batch(qs)
        pb = index.prepare(b)
        n = b.n_queries
        d_hits = torch.zeros((n, 10, 2), dtype=torch.int32, device=dev); d_n = torch.zeros(n, dtype=torch.int32, device=dev); d_c = torch.zeros(n, dtype=torch.int32, device=dev)
        ms = []
        for it in range(5):
            flush.fill_(it)
            pb.execute(d_hits.data_ptr(), d_n.data_ptr(), (d_c.data_ptr() if os.environ.get('COUNTS') else None), None, k_stride=10, flags=(nat.FG_EXEC_COUNTERS if it == 0 else 0))
            if it == 0: s0 = pb.stats()
            s = pb.stats()
            if it >= 2: ms.append(s.search_kernel_ms)
        s = s0; by = s.bytes_blocks + s.scored_postings
        t = float(np.mean(ms))
        print(f"{q:42s} x{rep:6d} items={s.n_work_items:6d} {t:8.3f} ms  {by/1e6:8.1f} MB {by/t/1e6:8.1f} GB/s  {s.scored_postings/t/1e6:8.1f} Gpost/s  {rep/t:8.1f} kQPS  redecode={s.bytes_redecode/1e6:.1f}MB", flush=True)
        pb.close()
main()
