"""Synthetic Zipfian corpus + query generators.

Bench/test INPUT infrastructure: produces the flat per-field CSR the reference-side loader would
read out of tantivy, and fugu-syntax query strings. Not part of the query hot path, not the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SYNTH_LIB = os.path.join(_HERE, "synth", "libfugu_synth.so")
_lib = None


def _L():
    global _lib
    if _lib is None:
        if not os.path.exists(SYNTH_LIB):
            raise ImportError(f"{SYNTH_LIB} missing: run `make`")
        L = C.CDLL(SYNTH_LIB)
        vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        L.fgs_corpus_create.argtypes = [u64, u32, C.c_double, C.c_int, u32]
        L.fgs_corpus_create.restype = vp
        L.fgs_corpus_destroy.argtypes = [vp]
        L.fgs_corpus_destroy.restype = None
        L.fgs_doc_len.argtypes = [vp, u64]
        L.fgs_doc_len.restype = u32
        L.fgs_doc_tokens.argtypes = [vp, u64, C.c_int, vp]
        L.fgs_doc_tokens.restype = u32
        L.fgs_facet_vocab.argtypes = [vp]
        L.fgs_facet_vocab.restype = u32
        L.fgs_doc_facets.argtypes = [vp, u64, vp]
        L.fgs_doc_facets.restype = u32
        L.fgs_facet_path.argtypes = [vp, u32, C.c_char_p, C.c_int]
        L.fgs_build_csr.argtypes = [vp, u64, u64, C.c_int]
        L.fgs_build_csr.restype = vp
        for name, rt in [("n_terms", u32), ("n_postings", u64), ("n_docs", u64), ("total_tokens", u64),
                         ("offsets", C.POINTER(u64)), ("docs", C.POINTER(u32)), ("tfs", C.POINTER(u32)),
                         ("doc_len", C.POINTER(u32))]:
            f = getattr(L, "fgs_csr_" + name)
            f.argtypes = [vp]
            f.restype = rt
        L.fgs_csr_free.argtypes = [vp]
        L.fgs_csr_free.restype = None
        L.fgs_doc_text.argtypes = [vp, u64, C.c_int, C.c_char_p, C.c_int]
        _lib = L
    return _lib


# seeds corpus 0xF0C0 + cfg, queries 0xBEEF + cfg
@dataclass(frozen=True)
class Config:
    cfg: int
    n_docs: int
    vocab: int
    n_queries: int
    k: int
    name_pct: int = 0
    n_ns: int = 0

    @property
    def corpus_seed(self) -> int:
        return 0xF0C0 + self.cfg

    @property
    def query_seed(self) -> int:
        return 0xBEEF + self.cfg


CONFIGS = {
    1: Config(1, 10_000, 50_000, 1_000, 10),
    2: Config(2, 1_000_000, 200_000, 5_000, 10, name_pct=10),
    3: Config(3, 10_000_000, 500_000, 2_000, 100),
    4: Config(4, 100_000_000, 1_000_000, 2_000, 10),
    5: Config(5, 10_000_000, 500_000, 10_000, 10, n_ns=64),
}

FIELD_TEXT, FIELD_NAME, FIELD_FACET = 0, 1, 2


class Corpus:
    def __init__(self, seed: int, vocab: int, zipf_s: float = 1.0, name_pct: int = 0, n_ns: int = 0):
        self.h = _L().fgs_corpus_create(seed, vocab, zipf_s, name_pct, n_ns)
        self.vocab, self.name_pct, self.n_ns = vocab, name_pct, n_ns

    @classmethod
    def for_config(cls, c: Config, vocab: int | None = None) -> "Corpus":
        return cls(c.corpus_seed, vocab or c.vocab, 1.0, c.name_pct, c.n_ns)

    def csr(self, d0: int, d1: int, field: int) -> dict:
        """Flat CSR of one field over docs [d0, d1) with local ids; arrays are numpy copies."""
        L = _L()
        p = L.fgs_build_csr(self.h, d0, d1, field)
        try:
            nt, npost, nd = L.fgs_csr_n_terms(p), L.fgs_csr_n_postings(p), L.fgs_csr_n_docs(p)
            offs = np.ctypeslib.as_array(L.fgs_csr_offsets(p), (nt + 1,)).copy()
            docs = np.ctypeslib.as_array(L.fgs_csr_docs(p), (max(npost, 1),))[:npost].copy()
            tfs = np.ctypeslib.as_array(L.fgs_csr_tfs(p), (max(npost, 1),))[:npost].copy()
            dl = np.ctypeslib.as_array(L.fgs_csr_doc_len(p), (max(nd, 1),))[:nd].copy()
            return {"term_offsets": offs, "doc_ids": docs, "term_freqs": tfs, "doc_len": dl,
                    "total_num_tokens": int(L.fgs_csr_total_tokens(p))}
        finally:
            L.fgs_csr_free(p)

    def doc_text(self, d: int, field: int = 0) -> str:
        buf = C.create_string_buffer(8192)
        n = _L().fgs_doc_text(self.h, d, field, buf, 8192)
        assert n >= 0
        return buf.value.decode()

    def facet_path(self, ord_: int) -> str:
        buf = C.create_string_buffer(256)
        _L().fgs_facet_path(self.h, ord_, buf, 256)
        return buf.value.decode()

    def facet_vocab(self) -> int:
        return _L().fgs_facet_vocab(self.h)

    def close(self):
        if self.h:
            _L().fgs_corpus_destroy(self.h)
            self.h = None


def fieldnorm_ids(doc_len: np.ndarray) -> np.ndarray:
    """tantivy fieldnorm id of each token count, vectorised with the product's table."""
    from . import _native as nat
    table = np.array([nat.lib().fg_id_to_fieldnorm(i) for i in range(256)], dtype=np.uint64)
    return (np.searchsorted(table, doc_len.astype(np.uint64), side="right") - 1).astype(np.uint8)


def build_fields(corpus: Corpus, d0: int, d1: int, with_name: bool | None = None, with_facets: bool | None = None) -> list[dict]:
    """Field list (text[, name][, facet]) for fg_index_upload over docs [d0, d1)."""
    with_name = corpus.name_pct > 0 if with_name is None else with_name
    with_facets = corpus.n_ns > 0 if with_facets is None else with_facets
    fields = []
    for fid in [FIELD_TEXT] + ([FIELD_NAME] if with_name else []):
        c = corpus.csr(d0, d1, fid)
        c["fieldnorm_ids"] = fieldnorm_ids(c["doc_len"])
        fields.append(c)
    if with_facets:
        if not with_name:  # keep field ids stable: 0 text, 1 name, 2 facet
            fields.append({"term_offsets": np.zeros(1, np.uint64), "doc_ids": np.zeros(0, np.uint32),
                           "term_freqs": np.zeros(0, np.uint32), "fieldnorm_ids": np.zeros(d1 - d0, np.uint8),
                           "total_num_tokens": 0})
        c = corpus.csr(d0, d1, FIELD_FACET)
        c["term_freqs"] = None  # facet postings are Basic (tf == 1), no fieldnorms
        c["fieldnorm_ids"] = None
        fields.append(c)
    return fields


# ---- query strings (counter-based, same stream everywhere) ------------------------------------
_M = (1 << 64) - 1


def _splitmix64(x: int) -> int:
    z = (x + 0x9E3779B97F4A7C15) & _M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M
    return z ^ (z >> 31)


def _rng(seed: int, q: int, i: int) -> int:
    return _splitmix64((_splitmix64(seed ^ ((q * 0xD6E8FEB86659FD93) & _M)) + i) & _M)


def _u01(x: int) -> float:
    return (x >> 11) / 9007199254740992.0


def gen_queries(c: Config, n: int | None = None, vocab: int | None = None) -> list[dict]:
    """Returns [{"query": str, "filters": [str], "k": int}] per the config table in SURVEY.md 8(d)."""
    n = c.n_queries if n is None else n
    V = vocab or c.vocab
    seed = c.query_seed
    out = []
    if c.cfg == 2:
        R = max(2, V // 10)
        w = np.arange(1, R + 1, dtype=np.float64) ** -0.8
        cdf = np.cumsum(w / w.sum())
    for q in range(n):
        filters: list[str] = []
        if c.cfg == 1:  # wA AND wB, ranks uniform in [10, 2000]
            lo, hi = 10, min(2000, V)
            a = lo + _rng(seed, q, 0) % (hi - lo + 1)
            b = a
            i = 1
            while b == a:
                b = lo + _rng(seed, q, i) % (hi - lo + 1)
                i += 1
            s = f"w{a} AND w{b}"
        elif c.cfg == 2:  # 1-4 terms, half AND-joined, half space-joined, Zipf(0.8) over [1, V/10]
            nt = 1 + _rng(seed, q, 0) % 4
            conj = _rng(seed, q, 1) % 2 == 0
            terms: list[int] = []
            i = 2
            while len(terms) < nt:
                r = int(np.searchsorted(cdf, _u01(_rng(seed, q, i)), side="right")) + 1
                r = min(r, R)
                i += 1
                if r not in terms:
                    terms.append(r)
            s = (" AND " if conj else " ").join(f"w{t}" for t in terms)
        elif c.cfg == 3:  # OR of 2-6 distinct terms from ranks [1, 64]
            nt = 2 + _rng(seed, q, 0) % 5
            terms = []
            i = 1
            while len(terms) < nt:
                r = 1 + _rng(seed, q, i) % min(64, V)
                i += 1
                if r not in terms:
                    terms.append(r)
            s = " ".join(f"w{t}" for t in terms)
        elif c.cfg == 4:  # 3-term AND from ranks [1, 5000]
            terms = []
            i = 0
            while len(terms) < 3:
                r = 1 + _rng(seed, q, i) % min(5000, V)
                i += 1
                if r not in terms:
                    terms.append(r)
            s = " AND ".join(f"w{t}" for t in terms)
        elif c.cfg == 5:  # 2-term OR text + 1-2 facet filters (Must group, OR inside)
            terms = []
            i = 0
            while len(terms) < 2:
                r = 1 + _rng(seed, q, i) % min(2000, V)
                i += 1
                if r not in terms:
                    terms.append(r)
            s = " ".join(f"w{t}" for t in terms)
            nf = 1 + _rng(seed, q, 100) % 2
            for j in range(nf):
                ns = _rng(seed, q, 101 + 2 * j) % max(c.n_ns, 1)
                kind = _rng(seed, q, 102 + 2 * j) % 3
                if kind == 0:
                    filters.append(f"/namespace/ns{ns:02d}")
                elif kind == 1:
                    filters.append(f"/namespace/ns{ns:02d}/organization/org{_rng(seed, q, 110 + j) % 16}")
                else:
                    filters.append(f"namespace/ns{ns:02d}/data/type{_rng(seed, q, 120 + j) % 4}")
        else:
            raise ValueError(c.cfg)
        out.append({"query": s, "filters": filters, "k": c.k})
    return out


def lower_queries(qs: list[dict], vocab: int, n_text_fields: int = 1, facet_lookup=None):
    """Lower synthetic fugu-syntax strings ("w3 AND w9", "w3 w9 w41") the way Dataset::search does
    (src/db/search.rs:108-151): each word -> Should group over the default fields [text, name];
    AND -> Must clauses, whitespace -> Should clauses; filters -> one Must clause OR-ing the facet
    terms. term ordinal of "wN" is N-1."""
    from . import _native as nat

    out = []
    for q in qs:
        s = q["query"]
        conj = " AND " in s
        words = [w for w in s.replace(" AND ", " ").split() if w]
        clauses = []
        for w in words:
            r = int(w[1:])
            t = r - 1 if 1 <= r <= vocab else nat.FG_TERM_MISSING
            leaves = [(f, t, 1.0) for f in range(n_text_fields)]
            clauses.append((nat.FG_OCCUR_MUST if conj else nat.FG_OCCUR_SHOULD, leaves))
        filters = q.get("filters") or []
        if filters:
            # Bool[Must(text_query), Must(facet group)]: Should words collapse into one Must clause
            if not conj:
                merged = [l for _, ls in clauses for l in ls]
                clauses = [(nat.FG_OCCUR_MUST, merged)]
            fl = []
            for f in filters:
                fid, t = facet_lookup(f)
                fl.append((fid, t, 1.0))
            clauses.append((nat.FG_OCCUR_MUST, fl))
        out.append({"k": q["k"], "clauses": clauses})
    return nat.HostBatch(out)
