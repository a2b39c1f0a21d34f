"""ctypes binding of the C ABI in include/fugu_gpu.h (libfugu_gpu.so).

This is the same binding a reference-side maintainer would write in Rust (`extern "C"`), see
INTEGRATION.md. The library is loaded from the package directory (built in-tree by `make` /
`__graft_entry__.build()`); a missing library or a missing CUDA device is a hard error — there
is no CPU path behind these calls.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfugu_gpu.so")
# the host library (include/fugu_host.h: planner, tokenizer, dataset builder); contains no CUDA code and
# binds the device library on first use. None = libfugu_host.so next to LIB_PATH, or LIB_PATH itself when
# that directory has no separate host library (a monolithic build exports both ABIs).
HOST_LIB_PATH = None

FG_OK = 0
FG_ERR_INVALID = -1
FG_ERR_UNSUPPORTED = -2
FG_ERR_CUDA = -3
FG_ERR_OOM = -4
FG_ERR_NO_DEVICE = -5

FG_FIELD_HAS_FIELDNORMS = 1
FG_FIELD_HAS_FREQS = 2
FG_OCCUR_SHOULD, FG_OCCUR_MUST, FG_OCCUR_MUST_NOT = 0, 1, 2
FG_TERM_MISSING = 0xFFFFFFFF
FG_TERM_ALL = 0xFFFFFFFE
FG_EXEC_EXACT_ACCOUNTING = 1
FG_EXEC_DETERMINISTIC = 2
FG_EXEC_COUNTERS = 4
FG_EXEC_NO_PRUNE = 8
FG_PREP_NO_COLUMNS = 1
FG_PREP_LEGACY = 2
FG_PREP_PER_QUERY_STATUS = 4


class FgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fugu_gpu error {code}: {msg}")
        self.code = code


class FieldDesc(C.Structure):
    _fields_ = [
        ("flags", C.c_uint32),
        ("n_terms", C.c_uint32),
        ("total_num_tokens", C.c_uint64),
        ("fieldnorm_ids", C.c_void_p),
        ("term_offsets", C.c_void_p),
        ("doc_ids", C.c_void_p),
        ("term_freqs", C.c_void_p),
        ("global_doc_freq", C.c_void_p),
    ]


class IndexDesc(C.Structure):
    _fields_ = [
        ("n_docs", C.c_uint32),
        ("doc_id_base", C.c_uint32),
        ("global_n_docs", C.c_uint64),
        ("n_fields", C.c_uint32),
        ("reserved", C.c_uint32),
        ("fields", C.POINTER(FieldDesc)),
        ("alive_bitset", C.c_void_p),
    ]


class IndexInfo(C.Structure):
    _fields_ = [
        ("n_postings", C.c_uint64),
        ("n_blocks", C.c_uint64),
        ("packed_bytes", C.c_uint64),
        ("skip_bytes", C.c_uint64),
        ("device_bytes", C.c_uint64),
        ("n_docs", C.c_uint32),
        ("n_fields", C.c_uint32),
        ("column_bytes", C.c_uint64),
        ("n_columns", C.c_uint32),
        ("n_bitmaps", C.c_uint32),
        ("bitmap_bytes", C.c_uint64),
        ("appended_bytes_h2d", C.c_uint64),
    ]


class QueryBatch(C.Structure):
    _fields_ = [
        ("n_queries", C.c_uint32),
        ("n_clauses", C.c_uint32),
        ("n_leaves", C.c_uint32),
        ("reserved", C.c_uint32),
        ("queries", C.c_void_p),
        ("clauses", C.c_void_p),
        ("leaves", C.c_void_p),
    ]


class BatchStats(C.Structure):
    _fields_ = [
        ("bytes_blocks", C.c_uint64),
        ("bytes_redecode", C.c_uint64),
        ("scored_postings", C.c_uint64),
        ("n_work_items", C.c_uint64),
        ("n_launches", C.c_uint64),
        ("n_queries", C.c_uint64),
        ("sum_k", C.c_uint64),
        ("search_kernel_ms", C.c_float),
        ("merge_kernel_ms", C.c_float),
        ("colscan_chunks", C.c_uint64),
        ("colscan_chunks_skipped", C.c_uint64),
        ("bytes_meta", C.c_uint64),
        ("lead_blocks", C.c_uint64),
        ("lead_blocks_seen", C.c_uint64),
        ("plan_bytes", C.c_uint64),
    ]


# numpy mirrors of the plain-data ABI structs
LEAF_DT = np.dtype([("field", "<u4"), ("term_ord", "<u4"), ("boost", "<f4")])
CLAUSE_DT = np.dtype([("occur", "<u4"), ("leaf_begin", "<u4"), ("n_leaves", "<u4")])
QUERY_DT = np.dtype([("k", "<u4"), ("clause_begin", "<u4"), ("n_clauses", "<u4")])
HIT_DT = np.dtype([("score", "<f4"), ("doc", "<u4")])

# every symbol include/fugu_gpu.h declares (tests check the .so exports all of them)
ABI_SYMBOLS = [
    "fg_last_error", "fg_version", "fg_ctx_create", "fg_ctx_destroy", "fg_ctx_set_stream",
    "fg_ctx_synchronize", "fg_index_upload", "fg_index_release", "fg_index_with_alive", "fg_index_append", "fg_index_get_info",
    "fg_index_term_info", "fg_search_batch", "fg_search_union_of", "fg_search_union_of_filtered", "fg_batch_prepare", "fg_batch_prepare_ex", "fg_batch_release",
    "fg_batch_execute", "fg_batch_submit", "fg_batch_collect", "fg_batch_get_stats", "fg_merge_topk_device", "fg_fieldnorm_to_id",
    "fg_comm_unique_id", "fg_comm_create", "fg_comm_destroy", "fg_comm_allreduce_sum_u64", "fg_comm_allreduce_sum_u32",
    "fg_batch_execute_sharded", "fg_batch_query_status", "fg_batch_submit_sharded", "fg_comm_info", "fg_comm_allgather_bytes",
    "fg_id_to_fieldnorm", "fg_bm25_idf",
]

_lib = None


def lib() -> C.CDLL:
    """Load libfugu_gpu.so (hard error when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make` or `python -c 'import __graft_entry__ as g; g.build()'`. "
            "fugu_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u32, i32, u64 = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64
    L.fg_last_error.restype = C.c_char_p
    L.fg_version.restype = C.c_char_p
    L.fg_ctx_create.argtypes = [i32, C.POINTER(vp)]
    L.fg_ctx_destroy.argtypes = [vp]
    L.fg_ctx_destroy.restype = None
    L.fg_ctx_set_stream.argtypes = [vp, vp]
    L.fg_ctx_synchronize.argtypes = [vp]
    L.fg_index_upload.argtypes = [vp, C.POINTER(IndexDesc), C.POINTER(vp)]
    L.fg_index_release.argtypes = [vp]
    L.fg_index_release.restype = None
    L.fg_index_with_alive.argtypes = [vp, vp, C.POINTER(vp)]
    L.fg_index_append.argtypes = [vp, C.POINTER(IndexDesc), vp, C.POINTER(vp)]
    L.fg_index_get_info.argtypes = [vp, C.POINTER(IndexInfo)]
    L.fg_index_term_info.argtypes = [vp, u32, u32, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u64)]
    L.fg_search_batch.argtypes = [vp, C.POINTER(QueryBatch), u32, vp, vp, vp]
    L.fg_search_union_of.argtypes = [vp, C.POINTER(QueryBatch), u32, vp, C.POINTER(u32), C.POINTER(u32)]
    L.fg_search_union_of_filtered.argtypes = [vp, C.POINTER(QueryBatch), u32, u32, vp, C.POINTER(u32), C.POINTER(u32)]
    L.fg_batch_prepare.argtypes = [vp, C.POINTER(QueryBatch), C.POINTER(vp)]
    L.fg_batch_prepare_ex.argtypes = [vp, C.POINTER(QueryBatch), u32, C.POINTER(vp)]
    L.fg_batch_release.argtypes = [vp]
    L.fg_batch_query_status.argtypes = [vp, vp]
    L.fg_batch_release.restype = None
    L.fg_batch_execute.argtypes = [vp, u32, u32, vp, vp, vp, vp]
    L.fg_batch_submit.argtypes = [vp, u32, u32, i32]
    L.fg_batch_collect.argtypes = [vp, vp, vp, vp]
    L.fg_batch_get_stats.argtypes = [vp, C.POINTER(BatchStats)]
    L.fg_merge_topk_device.argtypes = [vp, vp, vp, u32, u32, u32, u32, vp, vp]
    L.fg_comm_unique_id.argtypes = [vp]
    L.fg_comm_create.argtypes = [vp, i32, i32, vp, C.POINTER(vp)]
    L.fg_comm_destroy.argtypes = [vp]
    L.fg_comm_destroy.restype = None
    L.fg_comm_allreduce_sum_u64.argtypes = [vp, vp, C.c_size_t]
    L.fg_comm_allreduce_sum_u32.argtypes = [vp, vp, C.c_size_t]
    L.fg_batch_execute_sharded.argtypes = [vp, vp, u32, u32, vp, vp]
    L.fg_batch_submit_sharded.argtypes = [vp, vp, u32, u32]
    L.fg_comm_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    L.fg_comm_allgather_bytes.argtypes = [vp, vp, C.c_size_t, vp]
    L.fg_fieldnorm_to_id.argtypes = [u32]
    L.fg_fieldnorm_to_id.restype = C.c_uint8
    L.fg_id_to_fieldnorm.argtypes = [C.c_uint8]
    L.fg_id_to_fieldnorm.restype = u32
    L.fg_bm25_idf.argtypes = [u64, u64]
    L.fg_bm25_idf.restype = C.c_float
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != FG_OK:
        raise FgError(rc, lib().fg_last_error().decode("utf-8", "replace"))


_host_lib = None


def host_lib() -> C.CDLL:
    """Load libfugu_host.so (no CUDA code is mapped by this call)."""
    global _host_lib
    if _host_lib is not None:
        return _host_lib
    path = HOST_LIB_PATH
    if path is None:
        path = os.path.join(os.path.dirname(LIB_PATH), "libfugu_host.so")
        if not os.path.exists(path):
            path = LIB_PATH
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `make` or `python -c 'import __graft_entry__ as g; g.build()'`.")
    _host_lib = C.CDLL(path)
    _host_lib.fgh_last_error.restype = C.c_char_p
    return _host_lib


def hcheck(rc: int) -> None:
    """status of an fgh_* call (include/fugu_host.h)"""
    if rc != FG_OK:
        raise FgError(rc, host_lib().fgh_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return None if a is None else a.ctypes.data


class HostIndexDesc:
    """Owns numpy arrays + the ctypes fg_index_desc that points at them (used for both the GPU
    upload and, in tests, the oracle, which takes the very same descriptor)."""

    def __init__(self, n_docs: int, fields: list[dict], doc_id_base: int = 0, global_n_docs: int = 0,
                 alive_bitset: np.ndarray | None = None):
        self.n_docs = int(n_docs)
        self.keep = []
        self.fields = fields
        arr = (FieldDesc * len(fields))()
        for i, f in enumerate(fields):
            offs = np.ascontiguousarray(f["term_offsets"], dtype=np.uint64)
            docs = np.ascontiguousarray(f["doc_ids"], dtype=np.uint32)
            tfs = None if f.get("term_freqs") is None else np.ascontiguousarray(f["term_freqs"], dtype=np.uint32)
            fn = None if f.get("fieldnorm_ids") is None else np.ascontiguousarray(f["fieldnorm_ids"], dtype=np.uint8)
            gdf = None if f.get("global_doc_freq") is None else np.ascontiguousarray(f["global_doc_freq"], dtype=np.uint32)
            self.keep += [offs, docs, tfs, fn, gdf]
            flags = (FG_FIELD_HAS_FIELDNORMS if fn is not None else 0) | (FG_FIELD_HAS_FREQS if tfs is not None else 0)
            arr[i].flags = flags
            arr[i].n_terms = len(offs) - 1
            arr[i].total_num_tokens = int(f["total_num_tokens"])
            arr[i].fieldnorm_ids = _ptr(fn)
            arr[i].term_offsets = _ptr(offs)
            arr[i].doc_ids = _ptr(docs)
            arr[i].term_freqs = _ptr(tfs)
            arr[i].global_doc_freq = _ptr(gdf)
            f["_offs"], f["_docs"], f["_tfs"], f["_fn"] = offs, docs, tfs, fn
        self.field_arr = arr
        self.alive = None if alive_bitset is None else np.ascontiguousarray(alive_bitset, dtype=np.uint32)
        d = IndexDesc()
        d.n_docs = self.n_docs
        d.doc_id_base = int(doc_id_base)
        d.global_n_docs = int(global_n_docs)
        d.n_fields = len(fields)
        d.fields = C.cast(arr, C.POINTER(FieldDesc))
        d.alive_bitset = _ptr(self.alive)
        self.desc = d


class HostBatch:
    """A query batch in the flat ABI form (numpy structured arrays + the ctypes view)."""

    def __init__(self, queries: list[dict]):
        """queries: [{"k": int, "clauses": [(occur, [(field, term_ord, boost), ...]), ...]}, ...]"""
        nq = len(queries)
        ncl = sum(len(q["clauses"]) for q in queries)
        nl = sum(len(ls) for q in queries for _, ls in q["clauses"])
        self.q = np.zeros(nq, QUERY_DT)
        self.c = np.zeros(max(ncl, 1), CLAUSE_DT)
        self.l = np.zeros(max(nl, 1), LEAF_DT)
        ci = li = 0
        for qi, q in enumerate(queries):
            self.q[qi] = (q["k"], ci, len(q["clauses"]))
            for occ, leaves in q["clauses"]:
                self.c[ci] = (occ, li, len(leaves))
                ci += 1
                for f, t, b in leaves:
                    self.l[li] = (f, t, b)
                    li += 1
        b = QueryBatch()
        b.n_queries, b.n_clauses, b.n_leaves = nq, ncl, nl
        b.queries, b.clauses, b.leaves = _ptr(self.q), _ptr(self.c), _ptr(self.l)
        self.batch = b
        self.n_queries = nq
        self.kmax = int(self.q["k"].max()) if nq else 1

    @classmethod
    def from_arrays(cls, q: np.ndarray, c: np.ndarray, l: np.ndarray) -> "HostBatch":
        self = cls.__new__(cls)
        self.q, self.c, self.l = (np.ascontiguousarray(q, QUERY_DT), np.ascontiguousarray(c, CLAUSE_DT),
                                  np.ascontiguousarray(l, LEAF_DT))
        b = QueryBatch()
        b.n_queries, b.n_clauses, b.n_leaves = len(self.q), len(self.c), len(self.l)
        b.queries, b.clauses, b.leaves = _ptr(self.q), _ptr(self.c), _ptr(self.l)
        self.batch = b
        self.n_queries = len(self.q)
        self.kmax = int(self.q["k"].max()) if len(self.q) else 1
        return self


class Context:
    def __init__(self, device: int = 0):
        self.h = C.c_void_p()
        check(lib().fg_ctx_create(device, C.byref(self.h)))

    def set_stream(self, stream_ptr: int | None) -> None:
        check(lib().fg_ctx_set_stream(self.h, C.c_void_p(stream_ptr or 0)))

    def synchronize(self) -> None:
        check(lib().fg_ctx_synchronize(self.h))

    def close(self) -> None:
        if self.h:
            lib().fg_ctx_destroy(self.h)
            self.h = C.c_void_p()


class Index:
    def __init__(self, ctx: Context, desc: HostIndexDesc):
        self.ctx = ctx
        self.h = C.c_void_p()
        check(lib().fg_index_upload(ctx.h, C.byref(desc.desc), C.byref(self.h)))
        self.n_docs = desc.n_docs

    def with_alive(self, alive_bitset: np.ndarray | None) -> "Index":
        """fg_index_with_alive: a snapshot sharing this one's device arrays with another alive bitset."""
        nx = Index.__new__(Index)
        nx.ctx, nx.n_docs, nx.h = self.ctx, self.n_docs, C.c_void_p()
        bits = None if alive_bitset is None else np.ascontiguousarray(alive_bitset, dtype=np.uint32)
        assert bits is None or len(bits) == (self.n_docs + 31) // 32
        check(lib().fg_index_with_alive(self.h, _ptr(bits), C.byref(nx.h)))
        return nx

    def append(self, segment: HostIndexDesc, alive_bitset: np.ndarray | None = None) -> "Index":
        """fg_index_append: this snapshot + one new segment (its docs get the ids after this snapshot's)."""
        nx = Index.__new__(Index)
        nx.ctx, nx.n_docs, nx.h = self.ctx, self.n_docs + segment.n_docs, C.c_void_p()
        bits = None if alive_bitset is None else np.ascontiguousarray(alive_bitset, dtype=np.uint32)
        assert bits is None or len(bits) == (nx.n_docs + 31) // 32
        check(lib().fg_index_append(self.h, C.byref(segment.desc), _ptr(bits), C.byref(nx.h)))
        return nx

    def info(self) -> IndexInfo:
        i = IndexInfo()
        check(lib().fg_index_get_info(self.h, C.byref(i)))
        return i

    def term_info(self, field: int, term: int) -> dict:
        a, b, c, d = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint64()
        check(lib().fg_index_term_info(self.h, field, term, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"local_df": a.value, "global_df": b.value, "n_blocks": c.value, "bytes": d.value}

    def search(self, batch: HostBatch, k_stride: int | None = None, want_counts: bool = True):
        """fg_search_batch: host buffers in, host buffers out (the reference-facing call)."""
        ks = k_stride or batch.kmax
        hits = np.zeros((batch.n_queries, ks), HIT_DT)
        n = np.zeros(batch.n_queries, np.uint32)
        cnt = np.zeros(batch.n_queries, np.uint32) if want_counts else None
        check(lib().fg_search_batch(self.h, C.byref(batch.batch), ks, _ptr(hits), _ptr(n), _ptr(cnt)))
        return hits, n, cnt

    def search_union_of(self, disjuncts: HostBatch, k: int, n_filters: int = 0):
        """fg_search_union_of[_filtered]: the batch's queries are the Should children (boolean queries themselves) of ONE
        query; the last n_filters of them are filter children. Returns (hits[n], match_count)."""
        hits = np.zeros(max(k, 1), HIT_DT)
        n, cnt = C.c_uint32(), C.c_uint32()
        if n_filters:
            check(lib().fg_search_union_of_filtered(self.h, C.byref(disjuncts.batch), n_filters, k, _ptr(hits), C.byref(n), C.byref(cnt)))
            return hits[:n.value], int(cnt.value)
        check(lib().fg_search_union_of(self.h, C.byref(disjuncts.batch), k, _ptr(hits), C.byref(n), C.byref(cnt)))
        return hits[:n.value], int(cnt.value)

    def prepare(self, batch: HostBatch, prep_flags: int = 0) -> "PreparedBatch":
        return PreparedBatch(self, batch, prep_flags)

    def close(self) -> None:
        if self.h:
            lib().fg_index_release(self.h)
            self.h = C.c_void_p()


class PreparedBatch:
    def __init__(self, index: Index, batch: HostBatch, prep_flags: int = 0):
        self.index = index
        self.h = C.c_void_p()
        self.n_queries = batch.n_queries
        self.kmax = batch.kmax
        check(lib().fg_batch_prepare_ex(index.h, C.byref(batch.batch), prep_flags, C.byref(self.h)))

    def execute(self, d_hits: int, d_n: int, d_count: int | None = None, d_bitmap: int | None = None,
                k_stride: int | None = None, flags: int = 0) -> None:
        check(lib().fg_batch_execute(self.h, flags, k_stride or self.kmax, C.c_void_p(d_hits), C.c_void_p(d_n),
                                     C.c_void_p(d_count or 0), C.c_void_p(d_bitmap or 0)))

    def execute_sharded(self, comm: "Comm", d_hits: int, d_n: int, k_stride: int | None = None, flags: int = 0) -> None:
        """fg_batch_execute_sharded: local top-k -> NCCL all-gather -> on-device merge; global result on every rank."""
        check(lib().fg_batch_execute_sharded(self.h, comm.h, flags, k_stride or self.kmax, C.c_void_p(d_hits), C.c_void_p(d_n)))

    def submit(self, k_stride: int | None = None, want_counts: bool = True, flags: int = 0) -> None:
        """fg_batch_submit: launch + queue the device->host copy of the results; returns at once."""
        self._sub = (k_stride or self.kmax, want_counts)
        check(lib().fg_batch_submit(self.h, flags, self._sub[0], 1 if want_counts else 0))

    def collect(self):
        """fg_batch_collect: wait for a submitted batch, return (hits, n_hits, match_count | None)."""
        ks, want_counts = self._sub
        hits = np.zeros((self.n_queries, ks), HIT_DT)
        n = np.zeros(self.n_queries, np.uint32)
        cnt = np.zeros(self.n_queries, np.uint32) if want_counts else None
        check(lib().fg_batch_collect(self.h, _ptr(hits), _ptr(n), _ptr(cnt)))
        return hits, n, cnt

    def query_status(self) -> np.ndarray:
        st = np.zeros(self.n_queries, np.int32)
        check(lib().fg_batch_query_status(self.h, st.ctypes.data))
        return st

    def stats(self) -> BatchStats:
        s = BatchStats()
        check(lib().fg_batch_get_stats(self.h, C.byref(s)))
        return s

    def close(self) -> None:
        if self.h:
            lib().fg_batch_release(self.h)
            self.h = C.c_void_p()


FG_COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(FG_COMM_ID_BYTES)
    check(lib().fg_comm_unique_id(buf))
    return buf.raw


class Comm:
    """fg_comm: this rank's end of the NCCL communicator the library exchanges per-shard top-k lists over."""

    def __init__(self, ctx: Context, rank: int, world: int, unique_id: bytes):
        assert len(unique_id) == FG_COMM_ID_BYTES
        self.h = C.c_void_p()
        self.rank, self.world = rank, world
        check(lib().fg_comm_create(ctx.h, rank, world, C.c_char_p(unique_id), C.byref(self.h)))

    def allreduce_sum(self, a: np.ndarray) -> np.ndarray:
        """in-place sum over all ranks of a uint32 / uint64 host array (collective)"""
        assert a.flags["C_CONTIGUOUS"] and a.dtype in (np.uint32, np.uint64)
        f = lib().fg_comm_allreduce_sum_u64 if a.dtype == np.uint64 else lib().fg_comm_allreduce_sum_u32
        check(f(self.h, a.ctypes.data, a.size))
        return a

    def close(self) -> None:
        if self.h:
            lib().fg_comm_destroy(self.h)
            self.h = C.c_void_p()


def merge_topk_device(ctx: Context, d_hits: int, d_n: int, n_ranks: int, n_queries: int, k: int, k_stride: int,
                      d_out_hits: int, d_out_n: int) -> None:
    check(lib().fg_merge_topk_device(ctx.h, C.c_void_p(d_hits), C.c_void_p(d_n), n_ranks, n_queries, k, k_stride,
                                     C.c_void_p(d_out_hits), C.c_void_p(d_out_n)))
