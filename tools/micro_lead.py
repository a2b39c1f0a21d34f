"""Dev tool (GPU): kernel time + device counters of the lead-driven kernels per query shape of a config.
usage: micro_lead.py [cfg] [docs] [queries] [shapes...]   shapes: all and or single   env FG_* switches apply"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fugu_b200 import _native as nat  # noqa: E402
from fugu_b200 import synth  # noqa: E402

if os.environ.get("MICRO_LIB"):  # dev: time another build of the library
    nat.LIB_PATH = os.environ["MICRO_LIB"]


def main():
    cfgn = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    base = synth.CONFIGS[cfgn]
    docs = int(sys.argv[2]) if len(sys.argv) > 2 else base.n_docs
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else base.n_queries
    shapes = sys.argv[4:] or ["all"]
    reps = int(os.environ.get("REPS", "5"))
    cfg = synth.Config(cfg=base.cfg, n_docs=docs, vocab=base.vocab, n_queries=nq, k=base.k, name_pct=base.name_pct, n_ns=base.n_ns)
    corpus = synth.Corpus.for_config(cfg)
    t0 = time.time()
    fields = synth.build_fields(corpus, 0, docs, with_facets=False)
    desc = nat.HostIndexDesc(docs, fields)
    ctx = nat.Context(0)
    index = nat.Index(ctx, desc)
    print(f"corpus+upload {time.time() - t0:.1f}s", flush=True)
    qs_all = synth.gen_queries(cfg)

    n_col = index.info().n_columns  # text terms w1 .. w<n_col> own a dense tf column (ranks by descending df)

    def shape(q):
        s = q["query"]
        return "and" if " AND " in s else ("single" if " " not in s.strip() else "or")

    def has_col(q):
        return any(w.startswith("w") and w[1:].isdigit() and int(w[1:]) <= n_col for w in q["query"].split())

    def pick(sh):
        if sh == "all":
            return qs_all
        if sh.endswith("_col"):
            return [q for q in qs_all if shape(q) == sh[:-4] and has_col(q)]
        if sh.endswith("_nocol"):
            return [q for q in qs_all if shape(q) == sh[:-6] and not has_col(q)]
        return [q for q in qs_all if shape(q) == sh]

    if os.environ.get("SORT_QUERIES"):  # experiment: queries that lead with the same (rarest) terms next to each other
        def ranks(q):
            return sorted((int(w[1:]) for w in q["query"].split() if w.startswith("w") and w[1:].isdigit()), reverse=True)
        qs_all = sorted(qs_all, key=ranks)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for sh in shapes:
        qs = pick(sh)
        batch = synth.lower_queries(qs, vocab=cfg.vocab, n_text_fields=len(fields))
        n, k = batch.n_queries, batch.kmax
        d_hits = torch.zeros((n, k, 2), dtype=torch.int32, device="cuda")
        d_n = torch.zeros(n, dtype=torch.int32, device="cuda")
        for name, prep, flags in (("pruned", 0, 0), ("exhaustive", 0, nat.FG_EXEC_NO_PRUNE), ("legacy", nat.FG_PREP_LEGACY, 0)):
            if os.environ.get("ONLY") and name not in os.environ["ONLY"].split(","):
                continue
            pb = index.prepare(batch, prep)
            ms = []
            for i in range(reps):
                flush.fill_(i)
                pb.execute(d_hits.data_ptr(), d_n.data_ptr(), None, None, k_stride=k, flags=flags)
                ms.append(pb.stats().search_kernel_ms)
            pb.execute(d_hits.data_ptr(), d_n.data_ptr(), None, None, k_stride=k, flags=flags | nat.FG_EXEC_COUNTERS)
            st = pb.stats()
            print(f"{sh:7s} {name:10s} n={n:5d} items={st.n_work_items:6d} kernel ms min {min(ms):.3f} med {np.median(ms):.3f} | "
                  f"lead blocks {st.lead_blocks}/{st.lead_blocks_seen} block MB {st.bytes_blocks / 1e6:.1f} meta MB {st.bytes_meta / 1e6:.1f} "
                  f"gathers M {st.scored_postings / 1e6:.1f}", flush=True)
            pb.close()


if __name__ == "__main__":
    main()
