"""Dev tool (CPU, numpy): how much of a C2/C3-shaped batch does lead-driven MaxScore + block-max pruning touch?

For every pure-union query: leaves sorted by descending upper bound ub = w * max_factor; leaf i is walked
as a lead (its docs not contained in an earlier lead) while sum(ub[i:]) >= theta; inside leaf i a 128-posting
block is decoded only if w_i * blockmax_i(b) + sum(ub[i+1:]) >= theta. theta = the exact k-th best score
(best case; the kernels reach it progressively). Intersections: the lead clause is walked, blocks pruned by
w_lead * blockmax + sum(ub others) >= theta.
Prints postings touched vs exhaustive. Not part of the product or the tests.
"""
from __future__ import annotations

import sys
import time

import numpy as np

sys.path.insert(0, ".")
from fugu_b200 import synth  # noqa: E402

K1, B = 1.2, 0.75
import os
ORDER = os.environ.get("ORDER", "df")


def fn_table():
    t = np.zeros(256, np.uint64)
    for b in range(256):
        if b < 24:
            t[b] = b
        else:
            i = b - 24
            bits, sh = i & 7, i >> 3
            t[b] = 24 + (bits if sh == 0 else ((bits | 8) << (sh - 1)))
    return t


def main():
    cfgn = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    n_docs = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else 400
    base = synth.CONFIGS[cfgn]
    cfg = synth.Config(cfg=base.cfg, n_docs=n_docs, vocab=base.vocab, n_queries=nq, k=base.k, name_pct=base.name_pct, n_ns=base.n_ns)
    corpus = synth.Corpus.for_config(cfg)
    t0 = time.time()
    fields = synth.build_fields(corpus, 0, n_docs, with_facets=False)
    print("fields built", time.time() - t0, file=sys.stderr)
    tab = fn_table().astype(np.float32)
    caches = []
    for f in fields:
        avg = np.float32(f["total_num_tokens"]) / np.float32(n_docs)
        caches.append((K1 * (1 - B + B * tab / avg)).astype(np.float32))
    qs = synth.gen_queries(cfg)
    k = cfg.k
    tot_exh = tot_touch = tot_lead_blocks = tot_cand = 0
    by_shape = {}
    for q in qs:
        s = q["query"]
        conj = " AND " in s
        words = [int(w[1:]) for w in s.replace(" AND ", " ").split()]
        clauses = []
        for r in words:
            leaves = []
            for fi, f in enumerate(fields):
                o0, o1 = int(f["term_offsets"][r - 1]), int(f["term_offsets"][r])
                if o1 == o0:
                    continue
                docs = f["doc_ids"][o0:o1]
                tfs = f["term_freqs"][o0:o1].astype(np.float32)
                df = o1 - o0
                idf = np.log(np.float32(1) + (np.float32(n_docs - df) + np.float32(0.5)) / (np.float32(df) + np.float32(0.5)))
                w = np.float32(idf * (1 + K1))
                fac = tfs / (tfs + caches[fi][f["fieldnorm_ids"][docs]])
                leaves.append((docs, w * fac, w, fac))
            clauses.append(leaves)
        acc = np.zeros(n_docs, np.float32)
        if conj:
            ok = np.ones(n_docs, bool)
            for cl in clauses:
                m = np.zeros(n_docs, bool)
                for docs, sc, w, fac in cl:
                    m[docs] = True
                    acc[docs] += sc
                ok &= m
            acc[~ok] = -1
        else:
            for cl in clauses:
                for docs, sc, w, fac in cl:
                    acc[docs] += sc
            acc[acc == 0] = -1
        nm = int((acc > 0).sum())
        theta = np.partition(acc, n_docs - k)[n_docs - k] if nm >= k else -1.0
        exh = sum(len(l[0]) for cl in clauses for l in cl)
        if conj:
            cl_df = [sum(len(l[0]) for l in cl) for cl in clauses]
            lead = int(np.argmin(cl_df))
            ub_others = sum(max((float(l[2] * l[3].max()) for l in cl), default=0) if False else sum(float(l[2] * l[3].max()) for l in cl)
                            for i, cl in enumerate(clauses) if i != lead)
            touch = 0
            lead_leaves = sorted(clauses[lead], key=lambda l: -float(l[2] * l[3].max()))
            for i, (docs, sc, w, fac) in enumerate(lead_leaves):
                rest = sum(float(l[2] * l[3].max()) for l in lead_leaves[i + 1:])
                nb = (len(docs) + 127) // 128
                pad = np.full(nb * 128, 0, np.float32)
                pad[:len(docs)] = sc
                bm = pad.reshape(nb, 128).max(axis=1)
                keep = bm + rest + ub_others >= theta * (1 - 1e-6)
                touch += int(keep.sum()) * 128
                tot_lead_blocks += nb
        else:
            leaves = sorted([l for cl in clauses for l in cl], key=lambda l: (len(l[0]) if ORDER == "df" else -float(l[2] * l[3].max())))
            ubs = [float(l[2] * l[3].max()) for l in leaves]
            touch = 0
            for i, (docs, sc, w, fac) in enumerate(leaves):
                if sum(ubs[i:]) < theta * (1 - 1e-6):
                    break
                rest = sum(ubs[i + 1:])
                nb = (len(docs) + 127) // 128
                pad = np.full(nb * 128, 0, np.float32)
                pad[:len(docs)] = sc
                bm = pad.reshape(nb, 128).max(axis=1)
                keep = bm + rest >= theta * (1 - 1e-6)
                touch += int(keep.sum()) * 128
                tot_lead_blocks += nb
                surv = (pad.reshape(nb, 128)[keep] + rest >= theta * (1 - 1e-6)) & (pad.reshape(nb, 128)[keep] > 0)
                tot_cand += int(surv.sum())
                cand_q = cand_q + int(surv.sum()) if 'cand_q' in dir() else int(surv.sum())
        tot_exh += exh
        tot_touch += touch
        key = ("AND" if conj else "OR", len(words))
        a = by_shape.setdefault(key, [0, 0, 0])
        a[0] += 1; a[1] += exh; a[2] += touch
    print(f"OR candidates surviving the per-posting bound (best-case theta): {tot_cand:,}")
    print(f"queries {len(qs)}  exhaustive postings {tot_exh:,}  touched (best-case theta) {tot_touch:,}  ratio {tot_touch / max(tot_exh, 1):.3f}")
    for key in sorted(by_shape):
        n, e, t = by_shape[key]
        print(f"  {key}: n={n} exhaustive/query {e / n:,.0f} touched/query {t / n:,.0f} ({t / max(e, 1):.3f})")


if __name__ == "__main__":
    main()
