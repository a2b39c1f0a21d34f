"""Dev tool: time the search kernel for a few planner settings (env vars are read at prepare time)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fugu_b200 import _native as nat, synth

def main():
    nq = int(os.environ.get("SWEEP_Q", "5000"))
    cfg = synth.Config(cfg=2, n_docs=1_000_000, vocab=200_000, n_queries=nq, k=10, name_pct=10)
    corpus = synth.Corpus.for_config(cfg)
    fields = synth.build_fields(corpus, 0, cfg.n_docs)
    desc = nat.HostIndexDesc(cfg.n_docs, fields)
    ctx = nat.Context(0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    index = nat.Index(ctx, desc)
    queries = synth.gen_queries(cfg)
    batch = synth.lower_queries(queries, vocab=cfg.vocab, n_text_fields=2)
    # query classes
    kinds = {"all": queries,
             "and": [q for q in queries if " AND " in q["query"]],
             "or": [q for q in queries if " AND " not in q["query"] and " " in q["query"]],
             "single": [q for q in queries if " " not in q["query"]]}
    offs = [f["term_offsets"] for f in fields]
    def sumdf(q):
        tot = 0
        for w in q["query"].replace(" AND ", " ").split():
            t = int(w[1:]) - 1
            for o in offs: tot += int(o[t + 1] - o[t])
        return tot
    thr = 1536 * cfg.n_docs / 8192
    kinds["or_dense"] = [q for q in kinds["or"] + kinds["single"] if sumdf(q) >= thr]
    kinds["or_hash"] = [q for q in kinds["or"] + kinds["single"] if sumdf(q) < thr]
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    settings = json.loads(os.environ.get("SWEEP", '[{}]'))
    for s in settings:
        for k_, v in s.items(): os.environ[k_] = str(v)
        for kind, qs in kinds.items():
            if not qs or os.environ.get("SWEEP_KIND", kind) != kind: continue
            b = synth.lower_queries(qs, vocab=cfg.vocab, n_text_fields=2)
            pb = index.prepare(b)
            n = b.n_queries
            d_hits = torch.zeros((n, 10, 2), dtype=torch.int32, device=dev); d_n = torch.zeros(n, dtype=torch.int32, device=dev); d_c = torch.zeros(n, dtype=torch.int32, device=dev)
            ms = []
            for it in range(6):
                flush.fill_(it)
                pb.execute(d_hits.data_ptr(), d_n.data_ptr(), (d_c.data_ptr() if os.environ.get('COUNTS') else None), None, k_stride=10, flags=(nat.FG_EXEC_COUNTERS if it == 0 else 0))
                if it == 0: s0 = pb.stats()
                st = pb.stats()
                if it >= 2: ms.append(st.search_kernel_ms)
            st = s0; by = st.bytes_blocks + st.scored_postings
            print(f"{json.dumps(s):40s} {kind:7s} nq={n:5d} items={st.n_work_items:6d} search={np.mean(ms):8.3f} ms merge={st.merge_kernel_ms:6.3f} ms  {by/1e6:8.1f} MB  {by/np.mean(ms)/1e6:7.1f} GB/s  redecode={st.bytes_redecode/1e6:.1f}MB", flush=True)
            pb.close()
        for k_ in s: os.environ.pop(k_, None)

main()
