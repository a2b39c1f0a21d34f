"""ctypes binding of oracle/liboracle.so (the C++ restatement in oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY — the product (fugu_b200/) never imports this module.
It consumes the very same fg_index_desc / fg_query_batch structures the GPU library is given.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from fugu_b200 import _native as nat

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")
_lib = None


def _cpu_signature() -> str:
    """model + ISA flags of the host CPU: liboracle.so is built with -march=native, so a library built on another
    machine (the build container vs the GPU box) is rebuilt where it runs"""
    import hashlib

    sig = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith(("model name", "flags")):
                sig += line
                if line.startswith("flags"):
                    break
    except OSError:
        pass
    return hashlib.sha1(sig.encode()).hexdigest()[:16]


def build() -> None:
    src = os.path.join(_HERE, "oracle.cpp")
    mark = os.path.join(_HERE, ".built_for")
    sig = _cpu_signature()
    try:
        fresh = os.path.getmtime(LIB) >= os.path.getmtime(src) and open(mark).read().strip() == sig
    except OSError:
        fresh = False
    if fresh:
        return
    # -ffp-contract=off: tantivy is Rust, which never fuses a multiply and an add into an fma behind the programmer's back;
    # it also keeps the oracle's own evaluation forms (exhaustive, pruned) bit-identical to each other
    tmp = LIB + f".tmp{os.getpid()}"
    subprocess.check_call(["g++", "-O3", "-march=native", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", tmp, src])
    os.replace(tmp, LIB)
    open(mark, "w").write(sig)


def lib():
    global _lib
    if _lib is None:
        build()  # no-op when the library is current and was built for this CPU
        L = C.CDLL(LIB)
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
        L.orc_fieldnorm_to_id.argtypes = [u32]
        L.orc_fieldnorm_to_id.restype = C.c_uint8
        L.orc_id_to_fieldnorm.argtypes = [C.c_uint8]
        L.orc_id_to_fieldnorm.restype = u32
        L.orc_idf.argtypes = [u64, u64]
        L.orc_idf.restype = C.c_float
        L.orc_bm25_score.argtypes = [C.c_float, u64, u64, u64, C.c_uint8, u32]
        L.orc_bm25_score.restype = C.c_float
        L.orc_search_batch.argtypes = [C.POINTER(nat.IndexDesc), C.POINTER(nat.QueryBatch), u32, vp, vp, vp, vp, i32]
        L.orc_search_batch.restype = i32
        L.orc_algorithmic_bytes.argtypes = [C.POINTER(nat.IndexDesc), C.POINTER(nat.QueryBatch), vp, vp, i32]
        L.orc_algorithmic_bytes.restype = i32
        L.orc_search_union_of.argtypes = [C.POINTER(nat.IndexDesc), C.POINTER(nat.QueryBatch), u32, vp, C.POINTER(u32), C.POINTER(u32)]
        L.orc_search_union_of.restype = i32
        L.orc_search_union_of_filtered.argtypes = [C.POINTER(nat.IndexDesc), C.POINTER(nat.QueryBatch), u32, u32, vp, C.POINTER(u32), C.POINTER(u32)]
        L.orc_search_union_of_filtered.restype = i32
        L.orc_blockmax_build.argtypes = [C.POINTER(nat.IndexDesc), i32]
        L.orc_blockmax_build.restype = vp
        L.orc_blockmax_free.argtypes = [vp]
        L.orc_blockmax_free.restype = None
        L.orc_search_batch_pruned.argtypes = [C.POINTER(nat.IndexDesc), vp, C.POINTER(nat.QueryBatch), u32, vp, vp, i32]
        L.orc_search_batch_pruned.restype = i32
        _lib = L
    return _lib


class BlockMax:
    """Block-max metadata of an index (tantivy keeps it in its skip entries): built once per corpus, used by
    search_pruned."""

    def __init__(self, desc: nat.HostIndexDesc, threads: int = 1):
        self.desc = desc
        self.h = lib().orc_blockmax_build(C.byref(desc.desc), threads)

    def close(self):
        if self.h:
            lib().orc_blockmax_free(self.h)
            self.h = None


def search_pruned(bm: BlockMax, batch: nat.HostBatch, k_stride: int | None = None, threads: int = 1):
    """TopDocs form with block-max pruning where tantivy prunes (unions of plain term scorers, e.g. a single word over
    the two default fields); exhaustive scorers elsewhere. Returns (hits, n_hits)."""
    ks = k_stride or batch.kmax
    hits = np.zeros((batch.n_queries, ks), nat.HIT_DT)
    n = np.zeros(batch.n_queries, np.uint32)
    rc = lib().orc_search_batch_pruned(C.byref(bm.desc.desc), bm.h, C.byref(batch.batch), ks, hits.ctypes.data, n.ctypes.data, threads)
    if rc != 0:
        raise nat.FgError(rc, "oracle")
    return hits, n


def search(desc: nat.HostIndexDesc, batch: nat.HostBatch, k_stride: int | None = None, threads: int = 1,
           want_bitmap: bool = False):
    ks = k_stride or batch.kmax
    hits = np.zeros((batch.n_queries, ks), nat.HIT_DT)
    n = np.zeros(batch.n_queries, np.uint32)
    cnt = np.zeros(batch.n_queries, np.uint32)
    words = (desc.n_docs + 31) // 32
    bm = np.zeros((batch.n_queries, words), np.uint32) if want_bitmap else None
    rc = lib().orc_search_batch(C.byref(desc.desc), C.byref(batch.batch), ks, hits.ctypes.data, n.ctypes.data,
                                cnt.ctypes.data, None if bm is None else bm.ctypes.data, threads)
    if rc != 0:
        raise nat.FgError(rc, "oracle")
    return (hits, n, cnt, bm) if want_bitmap else (hits, n, cnt)


def algorithmic_bytes(desc: nat.HostIndexDesc, batch: nat.HostBatch, threads: int = 1):
    b = np.zeros(batch.n_queries, np.uint64)
    s = np.zeros(batch.n_queries, np.uint64)
    rc = lib().orc_algorithmic_bytes(C.byref(desc.desc), C.byref(batch.batch), b.ctypes.data, s.ctypes.data, threads)
    assert rc == 0
    return b, s


def search_union_of(desc: nat.HostIndexDesc, disjuncts: nat.HostBatch, k: int, n_filters: int = 0):
    """One query whose Should children are the batch's queries (BooleanQuery of boolean queries); the last n_filters of
    them are Must siblings of that union (facet filters). Returns (hits, count)."""
    hits = np.zeros(max(k, 1), nat.HIT_DT)
    n, cnt = C.c_uint32(), C.c_uint32()
    rc = lib().orc_search_union_of_filtered(C.byref(desc.desc), C.byref(disjuncts.batch), n_filters, k, hits.ctypes.data, C.byref(n), C.byref(cnt))
    if rc != 0:
        raise nat.FgError(rc, "oracle")
    return hits[:n.value], int(cnt.value)
