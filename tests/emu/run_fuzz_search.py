"""TEST INFRASTRUCTURE ONLY: differential fuzz of the reference-facing call on a box without a GPU. Random query strings
(bare words, field prefixes, boosts, + / -, AND / OR / AND NOT, parenthesised groups two levels deep with boosts, `*`,
stray punctuation and operator words that send the parser into the reference's escape-and-retry fallback), random facet filters
and pages go through Dataset.search on tests/emu/libfugu_emu.so (C++ parser -> planner -> lowering -> the CUDA kernels
under SIMT emulation) and through the independent Python twin (oracle/oracle_py.py: the same parse evaluated as a tree,
no flattening). Every request the device path answers must give the twin's page (scores within 1e-5, same documents
except inside ties); FG_ERR_UNSUPPORTED is allowed (the caller keeps tantivy for those), anything else is a failure.
usage: run_fuzz_search.py SEED N_QUERIES [N_DOCS]"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from fugu_b200 import _native as nat
from tests import util
nat.LIB_PATH = os.environ.get("FG_EMU_LIB") or os.path.join(ROOT, "tests", "emu", "libfugu_emu.so")  # (FG_EMU_LIB: the `make asan` build); util.EMULATED = True
from fugu_b200.dataset import Dataset, ObjectRecord
from oracle import oracle_py as op

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 300
ND = int(sys.argv[3]) if len(sys.argv) > 3 else 400
rng = random.Random(seed)
V = 40
def word(): return f"w{min(int(rng.paretovariate(0.9)), V)}"
ctx = nat.Context(0)
ds = Dataset(ctx); ix = op.PyIndex()
recs = []
for i in range(ND):
    text = " ".join(word() for _ in range(rng.randint(3, 40)))
    name = " ".join(word() for _ in range(rng.randint(1, 4))) if rng.random() < 0.3 else None
    facets = [f"/ns/n{i % 5}"] + ([f"/kind/k{i % 3}/sub{i % 2}"] if i % 4 == 0 else [])
    recs.append(ObjectRecord(id=f"d{i}", text=text, metadata={"name": name} if name else None, facets=facets))
    ix.upsert(f"d{i}", text, name, facets)
ds.upsert(recs, commit=False)
for i in range(0, ND, 17):
    ds.delete(f"d{i}", commit=False); ix.delete(f"d{i}")
ds.commit()

def term():
    t = word()
    if rng.random() < 0.05: t = rng.choice(["*", "W3", "w2.", "zzz", "w1,", "(w2", "w3)", "w4:", "w1-w2", '"w3"', "w5^", "AND", "OR", "NOT"])
    r = rng.random()
    if r < 0.1: t = "text:" + t
    elif r < 0.2: t = "name:" + t
    if rng.random() < 0.15: t += rng.choice(["^2", "^0.5", "^3.5", "^2", "^0.5", "^0", "^-1", "^1e3", "^.5"])
    return t
def group(depth):
    n = rng.randint(1, 4) if rng.random() < 0.92 else rng.randint(8, 20)  # (long ones: up to and beyond 32 leaves)
    parts = []
    for _ in range(n):
        if depth < 2 and rng.random() < 0.3:
            p = "(" + group(depth + 1) + ")" + (rng.choice(["", "", "^2", "^0.5"]))
        else:
            p = term()
        r = rng.random()
        if r < 0.12: p = "+" + p
        elif r < 0.2: p = "-" + p
        parts.append(p)
    j = rng.random()
    if j < 0.3: return " AND ".join(parts)
    if j < 0.45: return " OR ".join(parts)
    if j < 0.5 and len(parts) > 1: return parts[0] + " AND NOT " + " ".join(parts[1:])
    return " ".join(parts)
stats = {"ok": 0, "unsupported": 0, "invalid": 0, "both_err": 0}
bad = 0
for it in range(nq):
    q = group(0)
    if rng.random() < 0.03: q = rng.choice(["", "  ", "*", "* w1", "w1 AND *"])
    fl = rng.choice([[], [], [], ["/ns/n1"], ["/ns/n2", "/kind/k0/*"], ["*x*"], ["/nope"]])
    page, pp = rng.choice([(0, 10), (0, 20), (1, 5), (0, 100), (3, 7), (11, 100), (0, 1), (0, 10), (0, 20), (0, 32), (0, 33), (0, 128), (0, 129),
                           (7, 128), (0, 300), (0, 1024), (0, 1025)])  # (the top-k queue sizes 32 / 128 / 1024 and one past each)
    try:
        want, _ = op.search(ix, q, fl, page, pp)
        werr = None
    except Exception as e:
        want, werr = None, e
    try:
        res = ds.search(q, fl, page, pp)
    except nat.FgError as e:
        if e.code == nat.FG_ERR_UNSUPPORTED: stats["unsupported"] += 1; continue
        if werr is not None: stats["both_err"] += 1; continue
        stats["invalid"] += 1
        print("GPU path INVALID but twin ok:", repr(q), fl, e); bad += 1
        continue
    if werr is not None:
        print("twin error but GPU ok:", repr(q), fl, werr); bad += 1; continue
    gs = [r.score for r in res]; gd = [r.doc for r in res]
    ws = [s for _, s in want]; wd = [d for d, _ in want]
    ok = len(gs) == len(ws) and all(abs(a - b) <= 1e-5 * max(abs(a), abs(b), 1e-30) for a, b in zip(gs, ws))
    if ok and gd != wd:
        # docs may differ only inside score ties
        for i, (a, b) in enumerate(zip(gd, wd)):
            if a != b:
                tie = [j for j in range(len(ws)) if abs(ws[j] - ws[i]) <= 4e-5 * abs(ws[i])]
                if not (a in [wd[j] for j in tie] or i == len(ws) - 1 or max(tie) == len(ws) - 1): ok = False
    if not ok:
        bad += 1
        print("MISMATCH", repr(q), fl, page, pp, "\n  got ", list(zip(gd, gs))[:6], "\n  want", list(zip(wd, ws))[:6])
    else:
        stats["ok"] += 1
print(stats, "bad", bad)
sys.exit(1 if bad or stats["ok"] < nq // 2 else 0)
