"""TEST INFRASTRUCTURE ONLY: differential fuzz of the ingest side on a box without a GPU. A random sequence of
upserts (new ids and existing ones: delete + append, /root/reference/src/db/document.rs:23-67), deletes and commits runs
on a Dataset over tests/emu/libfugu_emu.so -- first commit = full upload, later ones = appended segments
(fg_index_append), delete-only refreshes (fg_index_with_alive) or a full re-upload once the appended part has outgrown
the rest -- and on the Python twin (oracle/oracle_py.py); after every commit random queries (AND / OR / nested, facet
filter, pages) must give the twin's hits and match counts, in the counting form and in the pruned TopDocs form.
usage: run_fuzz_ingest.py SEED ROUNDS BASE_DOCS"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fugu_b200 import _native as nat
from tests import util
nat.LIB_PATH = os.environ.get("FG_EMU_LIB") or os.path.join(ROOT, "tests", "emu", "libfugu_emu.so")  # (FG_EMU_LIB: the `make asan` build); util.EMULATED = True
from fugu_b200.dataset import Dataset, ObjectRecord
from oracle import oracle_py as op
seed = int(sys.argv[1]); rounds = int(sys.argv[2]); base_docs = int(sys.argv[3])
rng = random.Random(seed)
V = 60
def word(): return f"w{min(int(rng.paretovariate(0.8)), V)}"
ctx = nat.Context(0); ds = Dataset(ctx); ix = op.PyIndex()
next_id = 0; live = []
def make(i):
    text = " ".join(word() for _ in range(rng.randint(2, 30)))
    r_ = rng.random()
    if r_ < 0.04: text = " ".join([word()] * rng.randint(256, 400)) + " " + word()   # tf above 255: no tf column for that term
    elif r_ < 0.06: text = "!!! ??? ..."                                             # a document without a token
    elif r_ < 0.09: text = " ".join(word() for _ in range(rng.randint(500, 3000)))   # long: the upper fieldnorm buckets
    elif r_ < 0.11: text = "w1-w2 W3,w4. É" + word()                                 # punctuation, case, non-ASCII
    name = " ".join(word() for _ in range(rng.randint(1, 3))) if rng.random() < 0.3 else None
    facets = [f"/ns/n{i % 4}"]
    return ObjectRecord(id=f"d{i}", text=text, metadata={"name": name} if name else None, facets=facets), name
def add(n, existing_frac=0.0):
    global next_id
    recs = []
    for _ in range(n):
        if live and rng.random() < existing_frac: i = rng.choice(live)
        else:
            i = next_id; next_id += 1; live.append(i)
        r, name = make(i); recs.append(r); ix.upsert(r.id, r.text, name, r.facets)
    ds.upsert(recs, commit=False)
def check(tag):
    bad = 0
    for _ in range(25):
        k = rng.random()
        ws = [word() for _ in range(rng.randint(1, 3))]
        q = " AND ".join(ws) if k < 0.4 else " ".join(ws)
        if k > 0.85: q = f"({ws[0]} AND {word()}) OR {word()}"
        fl = rng.choice([[], [], ["/ns/n1"]])
        page, pp = rng.choice([(0, 10), (2, 10), (0, 100)])
        want, nm = op.search(ix, q, fl, page, pp)
        h, n, c, st = ds.search_batch([q], [fl], page, pp, want_counts=True)
        if st[0] == nat.FG_ERR_UNSUPPORTED: continue
        gs = h[0, :n[0]]["score"].tolist(); gd = h[0, :n[0]]["doc"].tolist()
        w_s = [s for _, s in want]; wd = [d for d, _ in want]
        ok = st[0] == 0 and len(gs) == len(w_s) and int(c[0]) == nm and all(abs(a - b) <= 1e-5 * max(abs(a), abs(b), 1e-30) for a, b in zip(gs, w_s))
        if ok and gd != wd:
            for j, (a, b) in enumerate(zip(gd, wd)):
                if a != b:
                    tie = [x for x in range(len(w_s)) if abs(w_s[x] - w_s[j]) <= 4e-5 * abs(w_s[j])]
                    if not (a in [wd[x] for x in tie] or max(tie) == len(w_s) - 1): ok = False
        h2, n2, _, st2 = ds.search_batch([q], [fl], page, pp, want_counts=False)   # the pruned TopDocs form
        if ok and not (st2[0] == 0 and n2[0] == n[0] and all(abs(a - b) <= 1e-5 * max(abs(a), abs(b), 1e-30) for a, b in zip(h2[0, :n2[0]]["score"].tolist(), w_s))): ok = False
        if not ok:
            bad += 1; print("MISMATCH", tag, repr(q), fl, page, pp, "cnt", int(c[0]), nm, "\n  got ", list(zip(gd, gs))[:5], "\n  want", list(zip(wd, w_s))[:5])
    return bad
bad = 0
add(base_docs); ds.commit(); bad += check("base")
for r in range(rounds):
    op_ = rng.random()
    if op_ < 0.5: add(rng.randint(1, max(2, len(live) // 2)), existing_frac=0.2)
    elif op_ < 0.7:
        for i in rng.sample(live, min(len(live), rng.randint(1, 20))):
            ds.delete(f"d{i}", commit=False); ix.delete(f"d{i}"); live.remove(i)
    else:
        add(rng.randint(1, 30), existing_frac=0.5)
        for i in rng.sample(live, min(len(live), 5)):
            ds.delete(f"d{i}", commit=False); ix.delete(f"d{i}"); live.remove(i)
    ds.commit()
    bad += check(f"round {r}")
print("commit_counts", ds.commit_counts(), "docs", len(live), "bad", bad)
sys.exit(1 if bad else 0)
