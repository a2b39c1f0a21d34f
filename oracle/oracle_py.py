"""oracle_py.py — TEST INFRASTRUCTURE ONLY: the slow, independently written pure-Python twin of
oracle/oracle.cpp, plus an independent query parser/planner, for tiny corpora.

PARITY UNPINNED (see oracle/oracle.cpp): restates tantivy 0.24.1 semantics as fugu configures them
 and fugu's own planning code (/root/reference/src/db/search.rs:74-324,
594-610). Everything is exhaustive set arithmetic over python dicts with numpy.float32 scalars, so
that it shares no structure with the C++ scorers it cross-checks.
"""
from __future__ import annotations

import math
import re


import regex

import numpy as np

f32 = np.float32
K1, B = f32(1.2), f32(0.75)


# ---- A.3 fieldnorm code ----------------------------------------------------------------------
def _byte4_to_int(b: int) -> int:
    bits, shift = b & 0x07, b >> 3
    return bits if shift == 0 else (bits | 0x08) << (shift - 1)


FIELDNORM_TABLE = [b if b < 24 else 24 + _byte4_to_int(b - 24) for b in range(256)]


def fieldnorm_to_id(n: int) -> int:
    best = 0
    for i, v in enumerate(FIELDNORM_TABLE):
        if v <= n:
            best = i
    return best


# ---- A.1 analyzer ----------------------------------------------------------------------------
_ALNUM = regex.compile(r"[\p{Alphabetic}\p{N}]")


def tokenize(text: str) -> list[str]:
    out, cur = [], []
    for ch in text + " ":
        if ch.isascii() and ch.isalnum() or not ch.isascii() and _ALNUM.match(ch):  # char::is_alphanumeric = Alphabetic (incl. Other_Alphabetic marks) or N*
            cur.append(ch)
        else:
            if cur:
                tok = "".join(cur)
                if len(tok.encode("utf-8")) < 40:  # RemoveLongFilter::limit(40)
                    out.append("".join(c.lower() for c in tok))  # LowerCaser, char by char
                cur = []
    return out


def facet_ancestors(path: str) -> list[str]:
    if not path.startswith("/"):
        return []
    segs = [s for s in path[1:].split("/")] if len(path) > 1 else []
    return ["/" + "/".join(segs[:i]) for i in range(1, len(segs) + 1)]


# ---- index -------------------------------------------------------------------------------------
class PyIndex:
    """docs: [{"id", "text", "name"|None, "facets": [...]}]; upsert semantics of
    src/db/document.rs:23-67 (delete by id, append)."""

    def __init__(self):
        self.post = [{}, {}, {}]  # field -> term -> {doc: tf}
        self.doc_len = [[], [], []]
        self.total_tokens = [0, 0, 0]
        self.ids: list[str] = []
        self.alive: list[bool] = []
        self.by_id: dict[str, int] = {}

    def upsert(self, id_: str, text: str, name: str | None = None, facets: list[str] | None = None):
        if id_ in self.by_id:
            self.alive[self.by_id.pop(id_)] = False
        d = len(self.ids)
        self.ids.append(id_)
        self.alive.append(True)
        self.by_id[id_] = d
        for f, s in ((0, text), (1, name)):
            toks = tokenize(s) if s else []
            self.doc_len[f].append(len(toks))
            self.total_tokens[f] += len(toks)
            for t in toks:
                self.post[f].setdefault(t, {}).setdefault(d, 0)
                self.post[f][t][d] += 1
        keys = []
        for p in facets or []:
            keys += facet_ancestors(p if p.startswith("/") else "/" + p)
        self.doc_len[2].append(len(keys))
        self.total_tokens[2] += len(keys)
        for k in set(keys):
            self.post[2].setdefault(k, {})[d] = 1

    def delete(self, id_: str):
        if id_ in self.by_id:
            self.alive[self.by_id.pop(id_)] = False

    @property
    def n_docs(self) -> int:
        return len(self.ids)


# ---- A.4 BM25 ------------------------------------------------------------------------------------
def idf(df: int, n: int) -> np.float32:
    x = (f32(n - df) + f32(0.5)) / (f32(df) + f32(0.5))
    return f32(math.log(f32(1.0) + x)) if False else np.log(f32(1.0) + x, dtype=np.float32)


def term_scores(ix: PyIndex, field: int, term: str, boost: float = 1.0) -> dict[int, np.float32]:
    plist = ix.post[field].get(term)
    if not plist:
        return {}
    n = ix.n_docs  # includes deleted docs (A.4)
    avg = f32(ix.total_tokens[field]) / f32(n)
    weight = f32(boost) * (idf(len(plist), n) * (f32(1.0) + K1))
    out = {}
    for d, tf in plist.items():
        fn_id = fieldnorm_to_id(ix.doc_len[field][d]) if field != 2 else fieldnorm_to_id(1)
        norm = K1 * (f32(1.0) - B + B * f32(FIELDNORM_TABLE[fn_id]) / avg)
        t = f32(tf if field != 2 else 1)
        out[d] = weight * (t / (t + norm))
    return out


# ---- A.5 boolean trees: ("term", field, token, boost) | ("all", boost) | ("bool", [(occ, node)]) ----
def evaluate(ix: PyIndex, node) -> dict[int, np.float32]:
    kind = node[0]
    if kind == "term":
        return term_scores(ix, node[1], node[2], node[3])
    if kind == "all":
        return {d: f32(node[1]) for d in range(ix.n_docs)}
    must, should, mnot = [], [], []
    for occ, child in node[1]:
        {"must": must, "should": should, "not": mnot}[occ].append(evaluate(ix, child))
    if must:
        docs = set(must[0])
        for m in must[1:]:
            docs &= set(m)
        res = {}
        for d in docs:
            s = f32(0.0)
            for m in must:
                s = s + m[d]
            for sh in should:
                if d in sh:
                    s = s + sh[d]
            res[d] = s
    elif should:
        res = {}
        for sh in should:
            for d, v in sh.items():
                res[d] = res.get(d, f32(0.0)) + v
    else:
        res = {}
    for m in mnot:
        for d in m:
            res.pop(d, None)
    return res


def top_k(ix: PyIndex, scores: dict[int, np.float32], k: int) -> list[tuple[int, float]]:
    """A.6: alive docs only, score desc then doc asc."""
    items = [(d, s) for d, s in scores.items() if ix.alive[d]]
    items.sort(key=lambda t: (-float(t[1]), t[0]))
    return [(d, float(s)) for d, s in items[:k]]


# ---- A.2 query grammar (independent recursive-descent implementation) ------------------------------
class ParseError(Exception):
    pass


class Unsupported(Exception):
    pass


# (a boost is `^` + a float as nom's `double` reads it: sign, digits, fraction, exponent)
_TOKEN = re.compile(r'\s*(\(|\)|"[^"]*"|[^\s()":^]+|\^[-+]?(?:[0-9]+\.?[0-9]*|\.[0-9]+)(?:[eE][-+]?[0-9]+)?|:|")')


class _Tok(str):
    adj = False  # no blank space between the previous token and this one


def _lex(q: str) -> list[str]:
    toks, pos = [], 0
    q = q.strip()
    while pos < len(q):
        m = _TOKEN.match(q, pos)
        if not m:
            raise ParseError(f"cannot lex at {pos}")
        tok = _Tok(m.group(1))
        tok.adj = pos > 0 and m.start(1) == pos
        pos = m.end()
        # AND / OR / NOT are operators only when blank space, '(' or the end of the string follows (tantivy's grammar
        # matches "AND " / "OR "); `(w1 OR)` holds the word "or" (the analyzer lowercases it anyway)
        if tok in ("AND", "OR", "NOT") and pos < len(q) and not q[pos].isspace() and q[pos] != "(":
            adj = tok.adj
            tok = _Tok(tok.lower())
            tok.adj = adj
        toks.append(tok)
    return toks


def parse_query(q: str):
    """-> tree with leaves ("lit", field|None, text, boost) / ("all", boost) and ("bool", [(occ|None, node)])"""
    toks = _lex(q)
    pos = 0

    def peek():
        return toks[pos] if pos < len(toks) else None

    def take():
        nonlocal pos
        pos += 1
        return toks[pos - 1]

    def boost_of(node):
        if peek() and peek().startswith("^"):
            b = float(take()[1:])
            if node[0] == "lit":
                return ("lit", node[1], node[2], node[3] * b)
            if node[0] == "all":
                return ("all", node[1] * b)
            return ("boost", b, node)
        return node

    def leaf():
        t = peek()
        if t is None:
            raise ParseError("eof")
        if t == "(":
            take()
            n = seq(True)
            if peek() != ")":
                raise ParseError("missing )")
            take()
            return boost_of(n)
        if t in (")", ":", '"') or t.startswith("^"):
            raise ParseError(f"unexpected {t}")
        take()
        if t.startswith('"'):
            return boost_of(("lit", None, t[1:-1], 1.0))
        if any(c in t for c in "[]{}~"):
            raise Unsupported("range / fuzzy")
        if t == "*":
            return boost_of(("all", 1.0))
        if peek() == ":":
            take()
            if peek() is None or not getattr(peek(), "adj", True):
                raise ParseError("missing value after ':'")  # `field: value` is not field:value
            v = leaf()
            if v[0] != "lit":
                raise Unsupported("field group")
            return ("lit", t, v[2], v[3])
        return boost_of(("lit", None, t, 1.0))

    def occur_leaf():
        t = peek()
        occ = None
        if t == "NOT":
            take()
            occ = "not"
        elif t and len(t) > 1 and t[0] in "+-" and not t.startswith('"'):
            toks[pos] = t[1:]
            occ = "must" if t[0] == "+" else "not"
        elif t in ("+", "-") and pos + 1 < len(toks):
            take()
            occ = "must" if t == "+" else "not"
        return occ, leaf()

    def item():
        groups = [[occur_leaf()]]
        chain = False
        while peek() in ("AND", "OR"):
            op = take()
            if peek() is None:
                raise ParseError("dangling operator")
            chain = True
            if op == "AND":
                groups[-1].append(occur_leaf())
            else:
                groups.append([occur_leaf()])
        if not chain:
            return groups[0][0]

        def conj(g):
            if len(g) == 1:
                return g[0]
            return None, ("bool", [(o or "must", n) for o, n in g])

        if len(groups) == 1:
            return conj(groups[0])
        return None, ("bool", [((o or "should"), n) for o, n in map(conj, groups)])

    def seq(in_paren):
        kids = []
        while peek() is not None and not (in_paren and peek() == ")"):
            if peek() == ")":
                raise ParseError("unbalanced )")
            kids.append(item())
        if not kids:
            raise ParseError("empty")
        if len(kids) == 1 and kids[0][0] is None:
            return kids[0][1]
        return ("bool", kids)

    tree = seq(False)
    if pos != len(toks):
        raise ParseError("trailing input")
    return tree


def escape_query_string(q: str) -> str:
    """src/db/search.rs:603-610"""
    return "".join(c for c in q if c not in '()[]{}":+-!~*?\\^')


_FIELDS = {"text": 0, "name": 1}


def _resolve(node, boost=1.0):
    """user tree -> evaluable tree (default fields [text, name], default occur Should)."""
    if node[0] == "all":
        return ("all", node[1] * boost)
    if node[0] == "boost":
        return _resolve(node[2], boost * node[1])
    if node[0] == "lit":
        _, fld, text, b = node
        toks = tokenize(text)
        if len(toks) > 1:
            raise Unsupported("phrase")
        if fld is not None:
            if fld not in _FIELDS:
                raise ParseError(f"Field does not exist: {fld}")
            flds = [_FIELDS[fld]]
        else:
            flds = [0, 1]
        leaves = [("term", f, toks[0], b * boost) for f in flds] if toks else []
        if len(leaves) == 1:
            return leaves[0]
        return ("bool", [("should", l) for l in leaves])
    return ("bool", [((occ or "should"), _resolve(ch, boost)) for occ, ch in node[1]])


def normalize_facet_path(p: str) -> str:
    return p if p.startswith("/") else "/" + p


def plan_tree(query: str, filters: list[str]):
    """The planning half of Dataset::search (src/db/search.rs:90-151) -> evaluable tree."""
    non_wild = [f for f in filters if not (f.startswith("*") and f.endswith("*"))]
    if not query.strip():
        text = ("all", 1.0)
    else:
        try:
            text = _resolve(parse_query(query))
        except ParseError:
            text = _resolve(parse_query(escape_query_string(query)))
    if not non_wild:
        return text
    terms = []
    for f in non_wild:
        n = normalize_facet_path(f)
        if n.endswith("/*"):
            path = n[:-2]
        elif "=" in n:
            path = n.split("=", 1)[0]
        else:
            path = n
        if path.startswith("/"):
            anc = facet_ancestors(path)
            terms.append(("term", 2, anc[-1] if anc else "", 1.0))
    facet = ("bool", [("should", t) for t in terms]) if terms else ("all", 1.0)
    if not query.strip():
        return facet
    return ("bool", [("must", text), ("must", facet)])


def search(ix: PyIndex, query: str, filters: list[str] | None = None, page: int = 0, per_page: int = 20):
    """Dataset::search -> [(doc, score)] of the requested page, plus the match count."""
    tree = plan_tree(query, filters or [])
    scores = evaluate(ix, tree)
    limit = page * per_page + per_page
    hits = top_k(ix, scores, limit)
    n_match = sum(1 for d in scores if ix.alive[d])
    return hits[page * per_page:page * per_page + per_page], n_match


# ---- facet counting (src/db/facet.rs:33-233): FacetCollector over AllQuery ------------------------
def _facet_sort_key(path: str):
    """tantivy orders facets by their encoded form (segments joined by NUL)."""
    return path[1:].replace("/", "\0").encode("utf-8")


def facet_collect(ix: PyIndex, root: str) -> list[tuple[str, int]]:
    """FacetCollector::for_field("facet") + add_facet(root), searched with AllQuery: for every ALIVE
    doc, each direct child of `root` that one of the doc's facets equals or descends from is counted
    once. -> [(child path, count)] in facet order, zero counts omitted."""
    if not root.startswith("/"):
        raise ValueError("Facet::from panics on a path without a leading '/'")
    prefix = "/" if root == "/" else root + "/"  # "/a/" is a facet with an empty last segment: no children
    counts: dict[str, int] = {}
    for d in range(ix.n_docs):
        if not ix.alive[d]:
            continue
        under = set()
        for key, plist in ix.post[2].items():
            if d in plist and key.startswith(prefix) and "/" not in key[len(prefix):]:
                under.add(key)
        for c in under:
            counts[c] = counts.get(c, 0) + 1
    return sorted(counts.items(), key=lambda t: _facet_sort_key(t[0]))


def facet_collect_recursive(ix: PyIndex, path: str, depth: int, max_depth, out: list) -> None:
    """collect_facets_recursive, src/db/facet.rs:206-233"""
    if max_depth is not None and depth >= max_depth:
        return
    for facet, count in facet_collect(ix, path):
        out.append((facet, count))
        facet_collect_recursive(ix, facet, depth + 1, max_depth, out)


def facet_tree(ix: PyIndex, max_depth=None) -> dict:
    """get_facet_tree, src/db/facet.rs:115-203 -> {"tree": nested dicts, "max_depth", "total_facets"}."""
    allf: list = []
    facet_collect_recursive(ix, "/", 0, max_depth, allf)
    tree: dict = {}
    deepest = 0
    for path, count in allf:
        comps = [c for c in path.split("/") if c]
        deepest = max(deepest, len(comps))
        if max_depth is not None and len(comps) >= max_depth:
            continue
        level = tree
        for i, c in enumerate(comps):
            last = i == len(comps) - 1
            node = level.setdefault(c, {"name": c, "path": "/" + "/".join(comps[:i + 1]), "count": count if last else 0, "children": {}})
            if last:
                node["count"] = count
            level = node["children"]

    def roll(node) -> int:
        if node["children"]:
            node["count"] = node["count"] + sum(roll(ch) for ch in node["children"].values())
        return node["count"]

    for n in tree.values():
        roll(n)
    return {"tree": tree, "max_depth": deepest, "total_facets": len(allf)}


def facet_filter_paths(tree: dict) -> dict:
    """get_all_filter_paths, src/db/facet.rs:236-270: every node that has leaf children -> their names."""
    out: dict = {}
    stack = list(tree.values())
    while stack:
        n = stack.pop()
        kids = n["children"]
        names = sorted(k for k in kids if not kids[k]["children"])
        if names:
            out[n["path"]] = names
        stack.extend(kids.values())
    return out
