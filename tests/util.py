"""Shared helpers of the parity tests: a tiny planner for the synthetic "wN" query grammar, the
tolerance-aware top-k comparison and the GPU-vs-oracle batch check."""
from __future__ import annotations

import numpy as np

from fugu_b200 import _native as nat

REL_TOL = 1e-5  # BASELINE.json north_star: BM25 scores within 1e-5 relative in f32


from fugu_b200.synth import lower_queries as plan_queries  # noqa: E402,F401


def close(a: float, b: float, tol: float = REL_TOL) -> bool:
    return abs(a - b) <= tol * max(abs(a), abs(b), 1e-30)


def check_topk(g: np.ndarray, o: np.ndarray, k: int, tol: float = REL_TOL, ctx: str = "") -> None:
    """g, o: HIT_DT arrays of equal length (the valid hits). Scores must agree rank-wise within
    tol; doc ids must agree except inside runs of scores tied within tol (and such a run may be
    cut differently only when it touches the k-th position)."""
    assert len(g) == len(o), f"{ctx}: n_hits {len(g)} != {len(o)}"
    n = len(o)
    for i in range(n):
        assert close(float(g["score"][i]), float(o["score"][i]), tol), \
            f"{ctx}: rank {i}: score {g['score'][i]!r} vs oracle {o['score'][i]!r}"
    i = 0
    while i < n:
        j = i
        while j + 1 < n and close(float(o["score"][j + 1]), float(o["score"][j]), 4 * tol):
            j += 1
        gd, od = set(g["doc"][i:j + 1].tolist()), set(o["doc"][i:j + 1].tolist())
        if gd != od:
            assert j == n - 1 and n == k, f"{ctx}: ranks {i}..{j}: docs {sorted(gd)} vs oracle {sorted(od)}"
        i = j + 1


import os  # noqa: E402

FULL = bool(os.environ.get("FG_EMU_FULL"))
EMULATED = False  # set by tests/test_emulated_kernels.py while the library under test is tests/emu/libfugu_emu.so


class DevBuf:
    """A zeroed (or copied) int32 buffer in the memory the library's kernels run on: a torch CUDA
    tensor — or host memory while the kernel sources run under the SIMT emulation of tests/emu."""

    def __init__(self, shape, src: np.ndarray | None = None):
        if EMULATED:
            self.a = np.zeros(shape, np.int32) if src is None else np.ascontiguousarray(src, dtype=np.int32).copy()
            self.ptr = self.a.ctypes.data
        else:
            import torch

            dev = torch.device("cuda:0")
            self.t = torch.zeros(shape, dtype=torch.int32, device=dev) if src is None else \
                torch.from_numpy(np.ascontiguousarray(src, dtype=np.int32).copy()).to(dev)
            self.ptr = self.t.data_ptr()
            torch.cuda.synchronize()

    def numpy(self) -> np.ndarray:
        return self.a.copy() if EMULATED else self.t.cpu().numpy()


def gpu_search_device(index: nat.Index, batch: nat.HostBatch, want_bitmap: bool = False, flags: int = 0,
                      prep_flags: int = 0):
    """Split-phase path with device buffers provided by torch; returns numpy results (+ stats).
    The byte counters are defined on the block path: accounting runs lower the plan without columns."""
    if flags & nat.FG_EXEC_EXACT_ACCOUNTING:
        prep_flags |= nat.FG_PREP_NO_COLUMNS | nat.FG_PREP_LEGACY

    ks = batch.kmax
    nq = batch.n_queries
    d_hits, d_n, d_c = DevBuf((nq, ks, 2)), DevBuf(nq), DevBuf(nq)
    words = (index.n_docs + 31) // 32
    d_bm = DevBuf((nq, words)) if want_bitmap else None
    pb = index.prepare(batch, prep_flags)
    pb.execute(d_hits.ptr, d_n.ptr, d_c.ptr, None if d_bm is None else d_bm.ptr, k_stride=ks, flags=flags)
    st = pb.stats()
    index.ctx.synchronize()
    hits = d_hits.numpy().view(np.uint32).reshape(nq, ks, 2)
    out = np.zeros((nq, ks), nat.HIT_DT)
    out["score"] = hits[:, :, 0].view(np.float32)
    out["doc"] = hits[:, :, 1]
    res = (out, d_n.numpy().view(np.uint32), d_c.numpy().view(np.uint32),
           None if d_bm is None else d_bm.numpy().view(np.uint32), st)
    pb.close()
    return res


def check_batch_against_oracle(index: nat.Index, desc: nat.HostIndexDesc, batch: nat.HostBatch,
                               bitmaps: bool = True, threads: int = 4, legacy: bool = True) -> dict:
    """GPU (through the C ABI) vs oracle on the same descriptor + plan: matched doc-id sets and
    counts bit-exact, scores within REL_TOL, order identical modulo ties."""
    from oracle import orc

    if bitmaps:
        o_hits, o_n, o_c, o_bm = orc.search(desc, batch, threads=threads, want_bitmap=True)
    else:
        o_hits, o_n, o_c = orc.search(desc, batch, threads=threads)
        o_bm = None
    g_hits, g_n, g_c, g_bm, st = gpu_search_device(index, batch, want_bitmap=bitmaps)
    # also through the blocking host-buffer call (with match counts it runs on the windowed accumulator kernels)
    h_hits, h_n, h_c = index.search(batch)
    assert np.array_equal(h_n, g_n) and np.array_equal(h_c, g_c)
    for qi in range(batch.n_queries):  # (another engine, another summation order: equal within tolerance)
        n = int(g_n[qi])
        check_topk(h_hits[qi, :n], g_hits[qi, :n], int(batch.q["k"][qi]), ctx=f"host-vs-device query {qi}")
    # the TopDocs form (no match counts, no bitmap): the form in which the lead-driven kernels prune (MaxScore / block
    # maxima); it must answer exactly like the exhaustive forms
    t_hits, t_n, t_c = index.search(batch, want_counts=False)
    assert t_c is None and np.array_equal(t_n, o_n), f"TopDocs form: n_hits differ for queries {np.nonzero(t_n != o_n)[0][:10]}"
    for qi in range(batch.n_queries):
        n = int(o_n[qi])
        check_topk(t_hits[qi, :n], o_hits[qi, :n], int(batch.q["k"][qi]), ctx=f"TopDocs form (pruned), query {qi}")
    # two executions are bit-identical (a document's score is summed by one lane in leaf order: no float atomics)
    if not (EMULATED and not FULL):
        d1 = gpu_search_device(index, batch)
        d2 = gpu_search_device(index, batch, flags=nat.FG_EXEC_DETERMINISTIC)
        assert np.array_equal(d1[0]["doc"], d2[0]["doc"]) and np.array_equal(d1[0]["score"], d2[0]["score"])
        assert np.array_equal(d1[0]["doc"], g_hits["doc"]) and np.array_equal(d1[0]["score"], g_hits["score"])
    # the block path alone (terms with a dense tf column evaluated from their posting blocks)
    if index.info().n_columns:
        b_hits, b_n, b_c, b_bm, _ = gpu_search_device(index, batch, want_bitmap=bitmaps, prep_flags=nat.FG_PREP_NO_COLUMNS)
        assert np.array_equal(b_c, o_c), f"block path: match counts differ for queries {np.nonzero(b_c != o_c)[0][:10]}"
        assert np.array_equal(b_n, o_n)
        if bitmaps:
            assert np.array_equal(b_bm, o_bm), "block path: matched doc-id sets differ"
        for qi in range(batch.n_queries):
            n = int(o_n[qi])
            check_topk(b_hits[qi, :n], o_hits[qi, :n], int(batch.q["k"][qi]), ctx=f"block path, query {qi}")
    # the windowed accumulator kernels of round 1 (FG_PREP_LEGACY: the counting form of the host-buffer calls and the
    # exact-accounting pass run on them). `index.search(batch)` above already went through them with match counts;
    # here also with the matched doc-id sets. (Under emulation only in FG_EMU_FULL runs: CPU suite time.)
    if legacy and not (EMULATED and not FULL):
        l_hits, l_n, l_c, l_bm, _ = gpu_search_device(index, batch, want_bitmap=bitmaps, prep_flags=nat.FG_PREP_LEGACY)
        assert np.array_equal(l_c, o_c), f"legacy kernels: match counts differ for queries {np.nonzero(l_c != o_c)[0][:10]}"
        assert np.array_equal(l_n, o_n)
        if bitmaps:
            assert np.array_equal(l_bm, o_bm), "legacy kernels: matched doc-id sets differ"
        for qi in range(batch.n_queries):
            n = int(o_n[qi])
            check_topk(l_hits[qi, :n], o_hits[qi, :n], int(batch.q["k"][qi]), ctx=f"legacy kernels, query {qi}")
    bad = np.nonzero(g_c != o_c)[0]
    assert len(bad) == 0, f"match counts differ for queries {bad[:10]}: gpu {g_c[bad[:10]]} oracle {o_c[bad[:10]]}"
    assert np.array_equal(g_n, o_n), f"n_hits differ: {np.nonzero(g_n != o_n)[0][:10]}"
    if bitmaps:
        diff = np.nonzero((g_bm != o_bm).any(axis=1))[0]
        assert len(diff) == 0, f"matched doc-id sets differ for queries {diff[:10]}"
    for qi in range(batch.n_queries):
        n = int(o_n[qi])
        check_topk(g_hits[qi, :n], o_hits[qi, :n], int(batch.q["k"][qi]), ctx=f"query {qi}")
    return {"stats": st, "o_counts": o_c}


def batch_from_plans(plans) -> nat.HostBatch:
    """fgh_plan_t list -> flat fg_query_batch (what fgh_search_batch assembles internally)."""
    qs = []
    for p in plans:
        d = p.as_dict()
        qs.append({"k": d["k"], "clauses": d["clauses"]})
    return nat.HostBatch(qs)


def golden(name: str):
    import json
    import os

    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)))


def golden_dataset(ctx):
    """The Dataset (C++ host layer) loaded with the golden corpus, plus the python-twin index."""
    from fugu_b200.dataset import Dataset, ObjectRecord
    from oracle import oracle_py as op

    g = golden("search_cases.json")
    ds = Dataset(ctx)
    ix = op.PyIndex()
    recs = []
    for d in g["docs"]:
        recs.append(ObjectRecord(id=d["id"], text=d["text"], metadata={"name": d["name"]} if d["name"] else None, facets=d["facets"]))
        ix.upsert(d["id"], d["text"], d["name"], d["facets"])
    ds.upsert(recs, commit=False)
    for i in g["deletes"]:
        ds.delete(i, commit=False)
        ix.delete(i)
    return g, ds, ix


def host_desc_from_pyindex(ix) -> nat.HostIndexDesc:
    """CSR descriptor (text, name, facet) of a python-twin index, term ordinals = sorted terms.
    Returns (desc, [term lists])."""
    fields, terms = [], []
    for f in range(3):
        ts = sorted(ix.post[f])
        offs, docs, tfs = [0], [], []
        for t in ts:
            for d in sorted(ix.post[f][t]):
                docs.append(d)
                tfs.append(ix.post[f][t][d])
            offs.append(len(docs))
        from oracle import oracle_py as op

        fd = {"term_offsets": np.array(offs, np.uint64), "doc_ids": np.array(docs, np.uint32),
              "total_num_tokens": ix.total_tokens[f]}
        if f != 2:
            fd["term_freqs"] = np.array(tfs, np.uint32)
            fd["fieldnorm_ids"] = np.array([op.fieldnorm_to_id(n) for n in ix.doc_len[f]], np.uint8)
        else:
            fd["term_freqs"] = None
            fd["fieldnorm_ids"] = None
        fields.append(fd)
        terms.append(ts)
    alive = np.zeros((ix.n_docs + 31) // 32, np.uint32)
    for d, a in enumerate(ix.alive):
        if a:
            alive[d >> 5] |= np.uint32(1 << (d & 31))
    return nat.HostIndexDesc(ix.n_docs, fields, alive_bitset=alive), terms
