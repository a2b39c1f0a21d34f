"""TEST INFRASTRUCTURE ONLY: differential fuzz of the collective host call over doc-id-range shards on a box without a
GPU. A random corpus (text, name, facets, deleted documents) is cut into R shards with GLOBAL statistics and term
ordinals; the ranks are threads of this process (own fg_ctx, shard dataset, fg_comm) over tests/emu/libfugu_emu.so and
the in-process NCCL stand-in (tests/emu/fake_nccl.cpp). Every rank passes the same random requests -- words, AND / OR,
boosts, nested groups, facet filters, first pages and deep pages -- to fgh_search_batch_sharded; what comes back must be
the page the Python twin (oracle/oracle_py.py) computes on the UNSHARDED corpus, on every rank.
usage: run_fuzz_sharded.py SEED WORLD N_REQUESTS"""
import ctypes as C
import os
import random
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
EMU = os.path.join(ROOT, "tests", "emu")
C.CDLL(os.path.join(EMU, "build", "fake", "libnccl.so.2"), mode=C.RTLD_GLOBAL)

import numpy as np  # noqa: E402

from fugu_b200 import _native as nat  # noqa: E402
from tests import util  # noqa: E402

nat.LIB_PATH = os.environ.get("FG_EMU_LIB") or os.path.join(EMU, "libfugu_emu.so")  # (FG_EMU_LIB: the `make asan` build)
util.EMULATED = True
assert "torch" not in sys.modules

from fugu_b200.dataset import Dataset, QuerySet  # noqa: E402
from oracle import oracle_py as op  # noqa: E402

seed, R, nq = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rng = random.Random(seed)
V, ND = 40, 900


def word():
    return f"w{min(int(rng.paretovariate(0.9)), V)}"


ix = op.PyIndex()
for i in range(ND):
    text = " ".join(word() for _ in range(rng.randint(3, 40)))
    name = " ".join(word() for _ in range(rng.randint(1, 4))) if rng.random() < 0.3 else None
    ix.upsert(f"d{i}", text, name, [f"/ns/n{i % 5}"] + ([f"/kind/k{i % 3}/sub{i % 2}"] if i % 4 == 0 else []))
for i in range(0, ND, 13):
    ix.delete(f"d{i}")
assert ix.n_docs == ND  # (no upsert of an existing id: doc ids are 0..ND)
terms = [sorted(ix.post[f]) for f in range(3)]
bounds = [ND * r // R for r in range(R + 1)]


def shard_desc(r):
    a, b = bounds[r], bounds[r + 1]
    fields = []
    for f in range(3):
        offs, docs, tfs, gdf = [0], [], [], []
        for t in terms[f]:
            pl = ix.post[f][t]
            for d in sorted(x for x in pl if a <= x < b):
                docs.append(d - a)
                tfs.append(pl[d])
            offs.append(len(docs))
            gdf.append(len(pl))
        fd = {"term_offsets": np.array(offs, np.uint64), "doc_ids": np.array(docs, np.uint32), "total_num_tokens": ix.total_tokens[f],
              "global_doc_freq": np.array(gdf, np.uint32), "term_freqs": None, "fieldnorm_ids": None}
        if f != 2:
            fd["term_freqs"] = np.array(tfs, np.uint32)
            fd["fieldnorm_ids"] = np.array([op.fieldnorm_to_id(n) for n in ix.doc_len[f][a:b]], np.uint8)
        fields.append(fd)
    alive = np.zeros((b - a + 31) // 32, np.uint32)
    for d in range(a, b):
        if ix.alive[d]:
            alive[(d - a) >> 5] |= np.uint32(1 << ((d - a) & 31))
    return nat.HostIndexDesc(b - a, fields, doc_id_base=a, global_n_docs=ND, alive_bitset=alive)


def term():
    t = word()
    r = rng.random()
    if r < 0.1:
        t = "text:" + t
    elif r < 0.2:
        t = "name:" + t
    if rng.random() < 0.15:
        t += rng.choice(["^2", "^0.5"])
    return t


def query():
    k = rng.random()
    n = rng.randint(1, 4)
    if k < 0.35:
        return " ".join(term() for _ in range(n))
    if k < 0.65:
        return " AND ".join(term() for _ in range(max(2, n)))
    if k < 0.8:
        return f"({term()} AND {term()}) OR ({term()} AND {term()})"
    if k < 0.9:
        return f"{term()} OR (+{term()} -{term()}) {term()}"
    return f"+{term()} {term()} -{term()}"


strings = [query() for _ in range(nq)]
filters = [rng.choice([[], [], ["/ns/n1"], ["/ns/n2", "/kind/k0/*"], ["*x*"], ["/nope"]]) for _ in strings]
pp = 50
pages = np.array([rng.choice([0, 0, 0, 1, 21]) for _ in strings], np.uint32)  # page 21 of 50: limit 1100, a deep page
want = []
for s_, fl, pg in zip(strings, filters, pages):
    try:
        want.append(op.search(ix, s_, fl, int(pg), pp)[0])
    except (op.Unsupported, op.ParseError) as e:
        want.append(e)

uid = nat.comm_unique_id()
errors: list[str] = []
answered = [0] * R


def rank_main(r: int) -> None:
    try:
        ctx = nat.Context(r)
        ds = Dataset(ctx)
        ds.adopt(shard_desc(r), terms)
        comm = nat.Comm(ctx, r, R, uid)
        q = QuerySet(strings, filters, 0, pp)
        q.pages[:] = pages
        h, n, st = ds.search_batch_sharded(comm, q)
        for i, s_ in enumerate(strings):
            w = want[i]
            if st[i] == nat.FG_ERR_UNSUPPORTED:
                continue
            if isinstance(w, Exception):
                if st[i] == 0:
                    errors.append(f"rank {r}: {s_!r} {filters[i]}: answered, the twin raises {w!r}")
                continue
            gs, gd = h[i, :n[i]]["score"].tolist(), h[i, :n[i]]["doc"].tolist()
            ws, wd = [x for _, x in w], [d for d, _ in w]
            ok = st[i] == 0 and len(gs) == len(ws) and all(abs(a - b) <= 1e-5 * max(abs(a), abs(b), 1e-30) for a, b in zip(gs, ws))
            if ok and gd != wd:  # documents may differ only inside score ties (or a tie cut by the end of the page)
                for j, (a, b) in enumerate(zip(gd, wd)):
                    if a != b:
                        tie = [x for x in range(len(ws)) if abs(ws[x] - ws[j]) <= 4e-5 * abs(ws[j])]
                        if not (a in [wd[x] for x in tie] or max(tie) == len(ws) - 1):
                            ok = False
            if not ok:
                errors.append(f"rank {r}: {s_!r} {filters[i]} page {pages[i]}: status {st[i]}, got {list(zip(gd, gs))[:4]} want {w[:4]}")
            else:
                answered[r] += 1
        comm.close()
        ds.close()
        ctx.close()
    except Exception as e:  # a rank that dies would leave the others waiting in a collective: report and exit hard
        sys.stderr.write(f"rank {r}: {type(e).__name__}: {e}\n")
        os._exit(2)


ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(R)]
for t in ts:
    t.start()
for t in ts:
    t.join()
if errors:
    sys.stderr.write("\n".join(errors[:12]) + "\n")
print(f"sharded fuzz x{R}: {answered} of {nq} requests answered and equal to the twin, {len(errors)} bad")
sys.exit(1 if errors or min(answered) < nq // 2 else 0)
