"""Python mirror of the reference's host interface for the query path, over the C ABI in
include/fugu_host.h (which itself mirrors the Rust code name by name).

    Dataset.search(query, filters, page, per_page)   <->  Dataset::search         src/db/search.rs:74-218
    Dataset.upsert(records) / ingest                  <->  DocumentOperations::upsert  src/db/document.rs:23-67
    ObjectRecord + facet path derivation              <->  src/object.rs:31-111, src/db/document.rs:277-312,
                                                           src/db/utils.rs:11-55
    perform_search (per_page clamp)                   <->  src/server/handlers/search.rs:350-402
    Dataset.list_facet / get_facets / get_facets_at / get_namespace_facets / get_available_namespaces /
    get_facet_tree / get_all_filter_paths             <->  src/db/facet.rs:33-270 (FacetCollector over AllQuery)

All searching happens on the GPU through libfugu_gpu.so; this module only marshals strings.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from . import _native as nat

FIELD_TEXT, FIELD_NAME, FIELD_FACET = 0, 1, 2
MAX_PLAN_CLAUSES, MAX_PLAN_LEAVES = 16, 64


class _Leaf(C.Structure):
    _fields_ = [("field", C.c_uint32), ("term_ord", C.c_uint32), ("boost", C.c_float)]


class _Clause(C.Structure):
    _fields_ = [("occur", C.c_uint32), ("leaf_begin", C.c_uint32), ("n_leaves", C.c_uint32)]


class Plan(C.Structure):
    _fields_ = [("k", C.c_uint32), ("offset", C.c_uint32), ("n_clauses", C.c_uint32), ("n_leaves", C.c_uint32),
                ("is_all", C.c_uint32), ("used_fallback", C.c_uint32),
                ("clauses", _Clause * MAX_PLAN_CLAUSES), ("leaves", _Leaf * MAX_PLAN_LEAVES), ("n_disjuncts", C.c_uint32)]

    def as_dict(self) -> dict:
        cl = []
        for i in range(self.n_clauses):
            c = self.clauses[i]
            cl.append((int(c.occur) & 0x7F, [(int(self.leaves[j].field), int(self.leaves[j].term_ord), float(self.leaves[j].boost))
                                             for j in range(c.leaf_begin, c.leaf_begin + c.n_leaves)]))
        d = {"k": int(self.k), "offset": int(self.offset), "is_all": bool(self.is_all),
             "used_fallback": bool(self.used_fallback), "clauses": cl}
        if self.n_disjuncts:  # nested query: which child of the top-level union each clause belongs to (1-based)
            d["disjunct_of_clause"] = [int(self.clauses[i].occur) >> 8 for i in range(self.n_clauses)]
            d["filter_child"] = any(int(self.clauses[i].occur) & 0x80 for i in range(self.n_clauses))  # the last child (facet filters)
        return d


HOST_SYMBOLS = [
    "fgh_last_error", "fgh_dataset_create", "fgh_dataset_destroy", "fgh_dataset_upsert", "fgh_dataset_delete", "fgh_dataset_commit", "fgh_dataset_commit_counts",
    "fgh_dataset_adopt", "fgh_dataset_num_docs", "fgh_dataset_index", "fgh_dataset_doc_id", "fgh_dataset_term_ord",
    "fgh_tokenize", "fgh_plan", "fgh_plan_batch", "fgh_search", "fgh_search_batch", "fgh_search_batch_sharded",
    "fgh_merge_shard_pages", "fgh_facet_children", "fgh_facet_counts",
    "fgh_batcher_create", "fgh_batcher_destroy", "fgh_batcher_search", "fgh_batcher_get_stats",
]
_bound = False


def _L():
    global _bound
    L = nat.host_lib()
    if not _bound:
        vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int32
        cpp = C.POINTER(C.c_char_p)
        L.fgh_dataset_create.argtypes = [vp, C.POINTER(vp)]
        L.fgh_dataset_destroy.argtypes = [vp]
        L.fgh_dataset_destroy.restype = None
        L.fgh_dataset_upsert.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p, cpp, u32]
        L.fgh_dataset_delete.argtypes = [vp, C.c_char_p]
        L.fgh_dataset_commit.argtypes = [vp]
        L.fgh_dataset_commit_counts.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.fgh_dataset_adopt.argtypes = [vp, C.POINTER(nat.IndexDesc), cpp, C.POINTER(C.c_uint64)]
        L.fgh_dataset_num_docs.argtypes = [vp]
        L.fgh_dataset_num_docs.restype = u32
        L.fgh_dataset_index.argtypes = [vp]
        L.fgh_dataset_index.restype = vp
        L.fgh_dataset_doc_id.argtypes = [vp, u32, C.c_char_p, u32]
        L.fgh_dataset_term_ord.argtypes = [vp, u32, C.c_char_p]
        L.fgh_dataset_term_ord.restype = u32
        L.fgh_tokenize.argtypes = [C.c_char_p, C.c_char_p, u32]
        L.fgh_plan.argtypes = [vp, C.c_char_p, cpp, u32, u32, u32, C.POINTER(Plan)]
        L.fgh_plan_batch.argtypes = [vp, u32, cpp, cpp, vp, vp, vp, vp, vp, u32, vp, u32, C.POINTER(u32), C.POINTER(u32), vp]
        L.fgh_search.argtypes = [vp, C.c_char_p, cpp, u32, u32, u32, vp, vp, vp]
        L.fgh_search_batch.argtypes = [vp, u32, cpp, cpp, vp, vp, vp, u32, vp, vp, vp, vp]
        L.fgh_search_batch_sharded.argtypes = [vp, vp, u32, cpp, cpp, vp, vp, vp, u32, vp, vp, vp]
        L.fgh_merge_shard_pages.argtypes = [vp, vp, u32, u32, u32, vp]
        L.fgh_merge_shard_pages.restype = u32
        for fn in (L.fgh_facet_children, L.fgh_facet_counts):
            fn.argtypes = [vp, C.c_char_p, u32, vp, u32, vp, u32, C.POINTER(u32), C.POINTER(u32)]
        L.fgh_batcher_create.argtypes = [vp, u32, u32, C.POINTER(vp)]
        L.fgh_batcher_destroy.argtypes = [vp]
        L.fgh_batcher_destroy.restype = None
        L.fgh_batcher_search.argtypes = [vp, C.c_char_p, cpp, u32, u32, u32, vp, C.POINTER(u32)]
        L.fgh_batcher_get_stats.argtypes = [vp, vp]
        _bound = True
    return L


def merge_shard_pages(lists: list[np.ndarray], page: int, per_page: int) -> np.ndarray:
    """fgh_merge_shard_pages: the page skip(page*per_page).take(per_page) of the merged order of per-shard result lists
    (each in TopDocs order, global doc ids)."""
    _L()
    ls = [np.ascontiguousarray(x, nat.HIT_DT) for x in lists]
    ptrs = (C.c_void_p * max(len(ls), 1))(*[x.ctypes.data for x in ls])
    lens = np.array([len(x) for x in ls], np.uint32)
    out = np.zeros(max(per_page, 1), nat.HIT_DT)
    n = _L().fgh_merge_shard_pages(ptrs, lens.ctypes.data, len(ls), page, per_page, out.ctypes.data)
    return out[:n]


def tokenize(text: str) -> list[str]:
    """tantivy "default" analyzer (SimpleTokenizer -> RemoveLongFilter(40) -> LowerCaser)."""
    raw = text.encode("utf-8")
    buf = C.create_string_buffer(2 * len(raw) + 64)
    n = _L().fgh_tokenize(raw, buf, len(buf))
    if n < 0:
        raise nat.FgError(nat.FG_ERR_INVALID, "tokenize buffer too small")
    out, pos = [], 0
    b = buf.raw
    for _ in range(n):
        e = b.index(b"\0", pos)
        out.append(b[pos:e].decode("utf-8"))
        pos = e + 1
    return out


@dataclass
class ObjectRecord:
    """src/object.rs:8-29 (fields that influence the docs index)."""
    id: str
    text: str
    metadata: dict | None = None
    facets: list[str] | None = None
    namespace: str | None = None
    organization: str | None = None
    conversation_id: str | None = None
    data_type: str | None = None

    def generate_namespace_facets(self) -> list[str]:
        """src/object.rs:81-111"""
        out = []
        if self.namespace is not None:
            out.append(f"/namespace/{self.namespace}")
            if self.organization is not None:
                out.append(f"/namespace/{self.namespace}/organization/{self.organization}")
            if self.conversation_id is not None:
                out.append(f"/namespace/{self.namespace}/conversation/{self.conversation_id}")
            if self.data_type is not None:
                out.append(f"/namespace/{self.namespace}/data/{self.data_type}")
        return out

    def validate(self) -> None:
        """ObjectRecord::validate, src/object.rs:31-78: raises ValueError with the reference's message. (Lengths are
        byte lengths, as Rust's `String::len`.)"""
        if not self.id:
            raise ValueError("Object ID cannot be empty")
        if len(self.id.encode()) > 256:
            raise ValueError("Object ID too long (max 256 characters)")
        if not self.text:
            raise ValueError("Object text cannot be empty")
        if len(self.text.encode()) > 10000:
            raise ValueError("Text too long (max 10000 characters)")
        if self.namespace is not None:
            if not self.namespace or "/" in self.namespace or " " in self.namespace:
                raise ValueError("Invalid namespace format")
            if len(self.namespace.encode()) > 128:
                raise ValueError("Namespace too long (max 128 characters)")
        if self.facets is not None:
            if len(self.facets) > 100:
                raise ValueError("Too many facets (max 100 per object)")
            for i, f in enumerate(self.facets):
                if not f:
                    raise ValueError(f"Facet at index {i} cannot be empty")
                if len(f.encode()) > 512:
                    raise ValueError(f"Facet at index {i} too long (max 512 characters)")

    def name(self) -> str | None:
        """metadata.name when it is a string (src/db/document.rs:131-139)."""
        if self.metadata and isinstance(self.metadata.get("name"), str):
            return self.metadata["name"]
        return None


def _metadata_facets(value: Any, prefix: list[str]) -> list[list[str]]:
    """create_metadata_facets, src/db/utils.rs:26-55"""
    out = []
    if isinstance(value, dict):
        for k, v in value.items():
            out += _metadata_facets(v, prefix + [k])
    elif isinstance(value, list):
        for it in value:
            out += _metadata_facets(it, list(prefix))
    elif isinstance(value, str) and value:
        out.append(prefix + [value])
    return out


def all_facet_paths(rec: ObjectRecord) -> list[str]:
    """get_all_facet_paths, src/db/document.rs:277-312 (including its use of only the FIRST path
    component of a metadata facet, `facet_path.first()`)."""
    out = []
    if rec.facets is not None:
        for p in rec.facets:
            out.append(p if p.startswith("/") else "/" + p)
    else:
        out += rec.generate_namespace_facets()
        if rec.metadata:
            for k, v in rec.metadata.items():
                for path in _metadata_facets(v, [k]):
                    first = path[0]
                    out.append(first if first.startswith("/") else f"/metadata/{first}")
    return out


@dataclass
class FuguSearchResult:
    """src/db/search.rs:20-27. `text` / `metadata` / `facets` are filled by Dataset.search from the host-side record
    table (hydration never touches the device); `doc` is the device's doc id."""
    id: str
    score: float
    doc: int = -1
    text: str = ""
    metadata: dict | None = None
    facets: list[str] | None = None

    def to_json(self) -> dict:
        """serde field order of FuguSearchResult"""
        return {"id": self.id, "score": self.score, "text": self.text, "metadata": self.metadata, "facets": self.facets}


@dataclass
class SearchResponse:
    """src/server/types.rs:146-152"""
    results: list[FuguSearchResult]
    total: int
    page: int
    per_page: int
    query: str


FACET_ENTRY_DT = np.dtype([("term_ord", "<u4"), ("depth", "<u4"), ("count", "<u8"), ("path_off", "<u4"), ("path_len", "<u4")])


@dataclass
class FacetNode:
    """src/db/facet.rs:16-22"""
    name: str
    path: str
    count: int
    children: dict = field(default_factory=dict)  # name -> FacetNode, kept sorted by name (BTreeMap)


@dataclass
class FacetTreeResponse:
    """src/db/facet.rs:25-30"""
    tree: dict
    max_depth: int
    total_facets: int


def build_facet_tree(all_facets: list[tuple[str, int]], max_depth: int | None) -> FacetTreeResponse:
    """The host half of get_facet_tree, src/db/facet.rs:121-203, quirks included: facets whose depth
    equals max_depth are collected (and counted in total_facets / max_depth) but not inserted, and a
    parent's final count is its own count PLUS the totals of its children."""
    tree: dict[str, FacetNode] = {}
    actual_max_depth = 0
    for path, count in all_facets:
        if path == "/":
            continue
        comps = [s for s in path.split("/") if s]
        depth = len(comps)
        actual_max_depth = max(actual_max_depth, depth)
        if max_depth is not None and depth >= max_depth:
            continue
        cur, cur_path = tree, ""
        for i, comp in enumerate(comps):
            cur_path += "/" + comp
            leaf = i == len(comps) - 1
            if comp not in cur:
                cur[comp] = FacetNode(comp, cur_path, count if leaf else 0)
            if leaf:
                cur[comp].count = count
            else:
                cur = cur[comp].children

    def update(node: FacetNode) -> int:
        if not node.children:
            return node.count
        total = node.count
        for ch in node.children.values():
            total += update(ch)
        node.count = total
        return total

    def sort_rec(m: dict) -> dict:
        return {k: FacetNode(v.name, v.path, v.count, sort_rec(v.children)) for k, v in sorted(m.items())}

    for n in tree.values():
        update(n)
    return FacetTreeResponse(sort_rec(tree), actual_max_depth, len(all_facets))


class QuerySet:
    """Request strings marshalled once into the char** arrays of the batched calls (so that a
    timed loop measures the library, not Python string encoding)."""

    def __init__(self, queries: list[str], filters: list[list[str]] | None = None, page: int = 0, per_page: int = 20):
        n = len(queries)
        self.n, self.page, self.per_page = n, page, per_page
        self.qarr = (C.c_char_p * max(n, 1))(*[q.encode() for q in queries])
        self.farr = self.foffs = None
        if filters is not None and any(filters):
            flat = [f.encode() for fl in filters for f in fl]
            self.farr = (C.c_char_p * max(len(flat), 1))(*flat)
            self.foffs = np.zeros(n + 1, np.uint32)
            self.foffs[1:] = np.cumsum([len(fl) for fl in filters])
        self.pages = np.full(n, page, np.uint32)
        self.pps = np.full(n, per_page, np.uint32)
        self.in_bytes = sum(len(q) + 1 for q in queries) + (0 if filters is None else sum(len(f) + 1 for fl in filters for f in fl))


class Dataset:
    """Mirror of `Dataset` (docs index only) whose search runs on the GPU."""

    def __init__(self, ctx: nat.Context | None):
        self.ctx = ctx
        self.h = C.c_void_p()
        nat.hcheck(_L().fgh_dataset_create(ctx.h if ctx is not None else None, C.byref(self.h)))
        self._keep = None
        # hit hydration (convert_doc_to_search_result, src/db/search.rs:534-590): id -> the stored fields of the
        # record, in host memory (row f1: a side table instead of a doc-store block decompress per hit)
        self._records: dict[str, ObjectRecord] = {}

    # ---- ingest -------------------------------------------------------------------------
    def upsert(self, records: list[ObjectRecord], commit: bool = True) -> None:
        L = _L()
        for r in records:
            self._records[r.id] = r
            facets = [f.encode() for f in all_facet_paths(r)]
            arr = (C.c_char_p * max(len(facets), 1))(*facets)
            nm = r.name()
            nat.hcheck(L.fgh_dataset_upsert(self.h, r.id.encode(), r.text.encode(), None if nm is None else nm.encode(),
                                           arr, len(facets)))
        if commit:
            self.commit()

    def delete(self, id_: str, commit: bool = True) -> None:
        self._records.pop(id_, None)
        nat.hcheck(_L().fgh_dataset_delete(self.h, id_.encode()))
        if commit:
            self.commit()

    def commit(self) -> None:
        nat.hcheck(_L().fgh_dataset_commit(self.h))

    def commit_counts(self) -> tuple[int, int]:
        """(snapshots built by a full upload, snapshots built by appending a segment)"""
        a, b = C.c_uint64(), C.c_uint64()
        nat.hcheck(_L().fgh_dataset_commit_counts(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def adopt(self, desc: nat.HostIndexDesc, terms: list[list[str] | None]) -> None:
        """Adopt a pre-built CSR (synthetic corpora) + per-field term dictionaries."""
        n = len(terms)
        blobs = [None if t is None else b"".join(x.encode() + b"\0" for x in t) for t in terms]
        bufs = [None if b is None else C.create_string_buffer(b, len(b)) for b in blobs]
        arr = (C.c_char_p * n)()
        for i, bf in enumerate(bufs):
            arr[i] = None if bf is None else C.cast(bf, C.c_char_p)
        sizes = (C.c_uint64 * n)(*[0 if b is None else len(b) for b in blobs])
        nat.hcheck(_L().fgh_dataset_adopt(self.h, C.byref(desc.desc), arr, sizes))
        self._keep = (desc, bufs)

    # ---- introspection ------------------------------------------------------------------
    @property
    def num_docs(self) -> int:
        return int(_L().fgh_dataset_num_docs(self.h))

    def doc_id(self, doc: int) -> str:
        buf = C.create_string_buffer(300)
        n = _L().fgh_dataset_doc_id(self.h, doc, buf, 300)
        return buf.value.decode() if n >= 0 else "unknown"

    def term_ord(self, field_: int, token: str) -> int:
        return int(_L().fgh_dataset_term_ord(self.h, field_, token.encode()))

    def plan(self, query: str, filters: list[str] | None = None, page: int = 0, per_page: int = 20) -> Plan:
        filters = filters or []
        arr = (C.c_char_p * max(len(filters), 1))(*[f.encode() for f in filters])
        p = Plan()
        nat.hcheck(_L().fgh_plan(self.h, query.encode(), arr, len(filters), page, per_page, C.byref(p)))
        return p

    def plan_batch(self, queries, filters: list[list[str]] | None = None, page: int = 0, per_page: int = 20):
        """fgh_plan_batch -> (HostBatch, status[n]): the flat fg_query_batch of n requests."""
        qs = queries if isinstance(queries, QuerySet) else QuerySet(queries, filters, page, per_page)
        n = qs.n
        q = np.zeros(n, nat.QUERY_DT)
        c = np.zeros(max(n * MAX_PLAN_CLAUSES, 1), nat.CLAUSE_DT)
        l = np.zeros(max(n * MAX_PLAN_LEAVES, 1), nat.LEAF_DT)
        nc, nl = C.c_uint32(), C.c_uint32()
        status = np.zeros(n, np.int32)
        nat.hcheck(_L().fgh_plan_batch(self.h, n, qs.qarr, qs.farr, None if qs.foffs is None else qs.foffs.ctypes.data,
                                      qs.pages.ctypes.data, qs.pps.ctypes.data, q.ctypes.data, c.ctypes.data, len(c),
                                      l.ctypes.data, len(l), C.byref(nc), C.byref(nl), status.ctypes.data))
        return nat.HostBatch.from_arrays(q, c[:nc.value].copy(), l[:nl.value].copy()), status

    def index(self) -> nat.Index:
        """Non-owning handle of the current device snapshot."""
        ix = nat.Index.__new__(nat.Index)
        ix.ctx = self.ctx
        ix.h = C.c_void_p(_L().fgh_dataset_index(self.h))
        ix.n_docs = self.num_docs
        ix.close = lambda: None
        return ix

    # ---- search -------------------------------------------------------------------------
    def search(self, query: str, filters: list[str] | None = None, page: int = 0, per_page: int = 20) -> list[FuguSearchResult]:
        """Dataset::search(query, filters, page, per_page) -> the requested page of hits."""
        filters = filters or []
        arr = (C.c_char_p * max(len(filters), 1))(*[f.encode() for f in filters])
        hits = np.zeros(max(per_page, 1), nat.HIT_DT)
        n = C.c_uint32()
        cnt = C.c_uint32()
        nat.hcheck(_L().fgh_search(self.h, query.encode(), arr, len(filters), page, per_page, hits.ctypes.data,
                                  C.byref(n), C.byref(cnt)))
        return [self._hydrate(self.doc_id(int(h["doc"])), float(h["score"]), int(h["doc"])) for h in hits[:n.value]]

    def _hydrate(self, id_: str, score: float, doc: int) -> FuguSearchResult:
        """convert_doc_to_search_result (src/db/search.rs:534-590) from the host-side record table: stored text,
        parsed metadata, the document's facet values (None when it has none)."""
        r = self._records.get(id_)
        if r is None:  # adopted (synthetic) corpora keep no records
            return FuguSearchResult(id_, score, doc)
        facets = all_facet_paths(r)
        return FuguSearchResult(id_, score, doc, text=r.text, metadata=r.metadata, facets=facets or None)

    def search_batch(self, queries, filters: list[list[str]] | None = None, page: int = 0, per_page: int = 20,
                     want_counts: bool = True):
        """Batched form: returns (hits[n, per_page], n_hits[n], match_count[n], status[n]).
        want_counts=False mirrors the reference exactly (TopDocs does not count matches): match_count is None."""
        qs = queries if isinstance(queries, QuerySet) else QuerySet(queries, filters, page, per_page)
        n, per_page = qs.n, qs.per_page
        hits = np.zeros((n, per_page), nat.HIT_DT)
        nh = np.zeros(n, np.uint32)
        cnt = np.zeros(n, np.uint32) if want_counts else None
        status = np.zeros(n, np.int32)
        nat.hcheck(_L().fgh_search_batch(self.h, n, qs.qarr, qs.farr, None if qs.foffs is None else qs.foffs.ctypes.data,
                                        qs.pages.ctypes.data, qs.pps.ctypes.data, per_page, hits.ctypes.data, nh.ctypes.data,
                                        None if cnt is None else cnt.ctypes.data, status.ctypes.data))
        return hits, nh, cnt, status

    def search_batch_sharded(self, comm, queries, filters: list[list[str]] | None = None, page: int = 0, per_page: int = 20):
        """Collective form for doc-id-range shards (one process per GPU): (hits[n, per_page], n_hits[n], status[n]) of
        the GLOBAL result on every rank."""
        qs = queries if isinstance(queries, QuerySet) else QuerySet(queries, filters, page, per_page)
        n, per_page = qs.n, qs.per_page
        hits = np.zeros((n, per_page), nat.HIT_DT)
        nh = np.zeros(n, np.uint32)
        status = np.zeros(n, np.int32)
        nat.hcheck(_L().fgh_search_batch_sharded(self.h, comm.h, n, qs.qarr, qs.farr, None if qs.foffs is None else qs.foffs.ctypes.data,
                                                qs.pages.ctypes.data, qs.pps.ctypes.data, per_page, hits.ctypes.data, nh.ctypes.data,
                                                status.ctypes.data))
        return hits, nh, status

    # ---- facets (src/db/facet.rs) -----------------------------------------------------------
    def _facets(self, fn, root: str, max_depth: int) -> list[tuple[str, int, int]]:
        n, nb = C.c_uint32(), C.c_uint32()
        r = root.encode()
        nat.hcheck(fn(self.h, r, max_depth, None, 0, None, 0, C.byref(n), C.byref(nb)))  # size query
        ents = np.zeros(max(n.value, 1), FACET_ENTRY_DT)
        buf = C.create_string_buffer(max(nb.value, 1))
        nat.hcheck(fn(self.h, r, max_depth, ents.ctypes.data, len(ents), buf, len(buf), C.byref(n), C.byref(nb)))
        raw = buf.raw
        return [(raw[int(e["path_off"]):int(e["path_off"]) + int(e["path_len"])].decode(), int(e["count"]), int(e["depth"]))
                for e in ents[:n.value]]

    def facet_children(self, root: str = "/", max_depth: int = 1) -> list[str]:
        """Facet dictionary entries below `root` (no device needed, no counts)."""
        return [p for p, _, _ in self._facets(_L().fgh_facet_children, root, max_depth)]

    def facet_counts(self, root: str = "/", max_depth: int = 1) -> list[tuple[str, int]]:
        """fgh_facet_counts: (facet path, alive docs under it) below root, counted on the device."""
        return [(p, c) for p, c, _ in self._facets(_L().fgh_facet_counts, root, max_depth)]

    def list_facet(self, from_level: str) -> list[tuple[str, int]]:
        """src/db/facet.rs:78-99: direct children of from_level with their doc counts."""
        return self.facet_counts(from_level, 1)

    def get_facets(self, namespace: str | None = None) -> list[tuple[str, int]]:
        """src/db/facet.rs:102-106"""
        return self.list_facet(namespace if namespace is not None else "/")

    def get_facets_at(self, prefix: str) -> list[tuple[str, int]]:
        """src/db/facet.rs:109-112"""
        return self.list_facet(prefix)

    def get_namespace_facets(self, namespace: str) -> list[tuple[str, int]]:
        """src/db/facet.rs:35-51"""
        return self.list_facet(f"/namespace/{namespace}")

    def get_available_namespaces(self) -> list[str]:
        """src/db/facet.rs:54-75"""
        out = []
        for p, _ in self.list_facet("/namespace"):
            if p.startswith("/namespace/") and "/" not in p[len("/namespace/"):]:
                out.append(p[len("/namespace/"):])
        return sorted(set(out))

    def get_facet_tree(self, max_depth: int | None = None) -> FacetTreeResponse:
        """src/db/facet.rs:115-203. The reference walks the tree with one FacetCollector search per
        node; here the whole walk is ONE device batch (one single-term count per facet)."""
        if max_depth is not None and max_depth <= 0:
            return build_facet_tree([], max_depth)  # collect_facets_recursive returns at once
        return build_facet_tree(self.facet_counts("/", max_depth or 0), max_depth)

    def get_all_filter_paths(self) -> dict[str, list[str]]:
        """src/db/facet.rs:236-270: parents that have leaf children -> the leaf names."""
        out: dict[str, list[str]] = {}

        def walk(node: FacetNode):
            if node.children:
                leaves = [name for name, ch in node.children.items() if not ch.children]
                if leaves:
                    out[node.path] = leaves
                for ch in node.children.values():
                    walk(ch)

        for n in self.get_facet_tree(None).tree.values():
            walk(n)
        return dict(sorted(out.items()))

    def close(self) -> None:
        if self.h:
            _L().fgh_dataset_destroy(self.h)
            self.h = C.c_void_p()


class Batcher:
    """Micro-batcher for single-query traffic (include/fugu_host.h fgh_batcher_*): `search`
    has Dataset.search's contract and may be called from any number of threads (one per in-flight HTTP request,
    src/server/handlers/search.rs:152); concurrent calls are answered together by one batched device call."""

    def __init__(self, ds: Dataset, max_batch: int = 4096, max_wait_us: int = 200):
        self.ds = ds
        self.h = C.c_void_p()
        nat.hcheck(_L().fgh_batcher_create(ds.h, max_batch, max_wait_us, C.byref(self.h)))

    def search_raw(self, query: str, filters: list[str] | None = None, page: int = 0, per_page: int = 20) -> np.ndarray:
        filters = filters or []
        arr = (C.c_char_p * max(len(filters), 1))(*[f.encode() for f in filters])
        hits = np.zeros(max(per_page, 1), nat.HIT_DT)
        n = C.c_uint32()
        nat.hcheck(_L().fgh_batcher_search(self.h, query.encode(), arr, len(filters), page, per_page, hits.ctypes.data, C.byref(n)))
        return hits[:n.value]

    def search(self, query: str, filters: list[str] | None = None, page: int = 0, per_page: int = 20) -> list[FuguSearchResult]:
        return [self.ds._hydrate(self.ds.doc_id(int(h["doc"])), float(h["score"]), int(h["doc"]))
                for h in self.search_raw(query, filters, page, per_page)]

    def stats(self) -> dict:
        st = np.zeros(4, np.uint64)
        nat.hcheck(_L().fgh_batcher_get_stats(self.h, st.ctypes.data))
        return {"n_requests": int(st[0]), "n_batches": int(st[1]), "max_batch_seen": int(st[2]), "wait_us_total": int(st[3])}

    def close(self) -> None:
        if self.h:
            _L().fgh_batcher_destroy(self.h)
            self.h = C.c_void_p()


def perform_search(datasets: dict[str, Dataset], namespace: str, query: str, filters: list[str], page: int,
                   per_page: int) -> SearchResponse:
    """src/server/handlers/search.rs:350-402: dataset lookup, per_page clamp to 20 when 0 or > 100."""
    ds = datasets.get(namespace)
    if ds is None:
        raise KeyError(f"Namespace '{namespace}' not found")
    if per_page == 0 or per_page > 100:
        per_page = 20
    results = ds.search(query, filters, page, per_page)
    return SearchResponse(results=results, total=len(results), page=page, per_page=per_page, query=query)
