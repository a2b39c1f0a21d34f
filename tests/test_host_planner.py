"""CPU: the C++ planner (fgh_plan, mirror of the planning half of Dataset::search,
/root/reference/src/db/search.rs:90-160,221-324,594-610) — branch coverage + cross-check of its
flat plans against the independent python parser/evaluator on query strings."""
import numpy as np
import pytest

from fugu_b200 import _native as nat
from fugu_b200.dataset import Dataset, ObjectRecord, all_facet_paths, perform_search
from oracle import oracle_py as op
from oracle import orc
from tests.util import batch_from_plans, check_topk, host_desc_from_pyindex

S, M, N = nat.FG_OCCUR_SHOULD, nat.FG_OCCUR_MUST, nat.FG_OCCUR_MUST_NOT


@pytest.fixture(scope="module")
def small():
    ix = op.PyIndex()
    docs = [("a1", "red apple pie", "Apple", ["/namespace/x/organization/o1"]),
            ("a2", "green apple", None, ["/namespace/x"]),
            ("a3", "red cherry pie pie", "Cherry pie", ["/namespace/y/data/pdf"]),
            ("a4", "blue berry", None, ["namespace/y"]),
            ("a5", "red red red", "red", [])]
    for i, t, n, f in docs:
        ix.upsert(i, t, n, f)
    desc, terms = host_desc_from_pyindex(ix)
    ds = Dataset(None)
    ds.adopt(desc, terms)
    yield ix, desc, ds
    ds.close()


def test_plan_shapes(small):
    ix, desc, ds = small
    t = lambda f, w: ds.term_ord(f, w)
    # bare words: Should group per word over the default fields [text, name] (search.rs:108-112)
    p = ds.plan("red apple").as_dict()
    assert p["k"] == 20 and p["offset"] == 0 and not p["is_all"]
    assert p["clauses"] == [(S, [(0, t(0, "red"), 1.0), (1, t(1, "red"), 1.0)]),
                            (S, [(0, t(0, "apple"), 1.0), (1, t(1, "apple"), 1.0)])]
    # AND -> Must clauses; a word missing from a field's dictionary keeps a MISSING leaf
    p = ds.plan("red AND berry", page=2, per_page=5).as_dict()
    assert p["k"] == 15 and p["offset"] == 10  # limit = page*per_page + per_page (search.rs:154-160)
    assert [c[0] for c in p["clauses"]] == [M, M]
    assert p["clauses"][1][1] == [(0, t(0, "berry"), 1.0), (1, nat.FG_TERM_MISSING, 1.0)]
    # + / - / NOT / boost / field:
    p = ds.plan("+red pie -cherry").as_dict()
    assert [c[0] for c in p["clauses"]] == [M, S, N]
    p = ds.plan("red AND NOT cherry").as_dict()
    assert [c[0] for c in p["clauses"]] == [M, N]
    p = ds.plan("name:apple^2.5").as_dict()
    assert p["clauses"] == [(S, [(1, t(1, "apple"), 2.5)])]
    # case folding and punctuation go through the analyzer
    assert ds.plan("RED, Apple!").as_dict()["clauses"] == ds.plan("red apple").as_dict()["clauses"]


def test_plan_filters(small):
    ix, desc, ds = small
    f = lambda p: ds.term_ord(2, p)
    # filters OR-ed, Must-joined with the text query (search.rs:141-144); Should words collapse into one Must group
    p = ds.plan("red apple", ["/namespace/x", "namespace/y"]).as_dict()
    assert [c[0] for c in p["clauses"]] == [M, M]
    assert len(p["clauses"][0][1]) == 4
    assert p["clauses"][1][1] == [(2, f("/namespace/x"), 1.0), (2, f("/namespace/y"), 1.0)]
    # "/x/*" -> exact term of the prefix (search.rs:273-281); "a=b" -> path a only (:306-313)
    assert ds.plan("red", ["/namespace/x/*"]).as_dict()["clauses"][1][1] == [(2, f("/namespace/x"), 1.0)]
    assert ds.plan("red", ["/namespace/y=zzz"]).as_dict()["clauses"][1][1] == [(2, f("/namespace/y"), 1.0)]
    # *x* filters are dropped before the facet query is built (F8): plan identical to no filter
    assert ds.plan("red", ["*name*"]).as_dict() == ds.plan("red").as_dict()
    # empty query + filters -> the facet query alone; empty query alone -> AllQuery
    p = ds.plan("", ["/namespace/x"]).as_dict()
    assert p["clauses"] == [(S, [(2, f("/namespace/x"), 1.0)])] and not p["is_all"]
    assert ds.plan("   ").as_dict()["is_all"]
    assert ds.plan("*").as_dict()["is_all"]
    # unknown facet path: MISSING leaf in a Must clause (matches nothing)
    assert ds.plan("red", ["/nope"]).as_dict()["clauses"][1][1] == [(2, nat.FG_TERM_MISSING, 1.0)]


def test_plan_errors_and_fallback(small):
    ix, desc, ds = small
    # parse error -> strip special characters and retry (search.rs:120-125, 603-610)
    p = ds.plan("(red apple")
    assert p.used_fallback and p.as_dict()["clauses"] == ds.plan("red apple").as_dict()["clauses"]
    p = ds.plan("red nosuchfield:apple")
    assert p.used_fallback  # FieldDoesNotExist -> retry without ':' -> word "nosuchfieldapple"
    with pytest.raises(nat.FgError) as e:
        ds.plan("red AND")  # still an error after the fallback -> Err (HTTP 500 in the reference)
    assert e.value.code == nat.FG_ERR_INVALID
    with pytest.raises(nat.FgError) as e:
        ds.plan("red", per_page=0)  # TopDocs::with_limit(0) panics
    assert e.value.code == nat.FG_ERR_INVALID
    # (`red OR apple AND pie` = red OR (apple AND pie): a union of boolean queries, planned as disjuncts since round 2)
    assert ds.plan("red OR apple AND pie").as_dict().get("disjunct_of_clause") == [1, 2, 2]
    for q in ['"red apple"', "foo-bar", "apple~1", "[a TO b]", "red OR apple AND (pie OR (tart AND jam))"]:
        with pytest.raises(nat.FgError) as e:
            ds.plan(q)
        assert e.value.code == nat.FG_ERR_UNSUPPORTED, q
    # a quoted single token is a plain term query
    assert ds.plan('"apple"').as_dict()["clauses"] == ds.plan("apple").as_dict()["clauses"]


QUERIES = ["red", "red apple", "red AND pie", "red AND pie AND cherry", "+red pie", "red -pie", "red pie -cherry -apple",
           "(red apple) AND pie", "red AND (apple OR cherry)", "NOT red", "-red", "red^3 pie", "text:pie name:pie",
           "pie pie", "apple AND apple", "name:red AND text:red", "blue OR green OR cherry", "(red", "red:", "zzz", "zzz AND red",
           # a required group without a Must child: one of its Should children has to match (found by the differential fuzz:
           # lifted as optional clauses, `red AND (-pie pie)` matched documents although the group matches nothing)
           "red AND (-pie pie)", "red AND (-pie apple)", "+(apple cherry -green) red", "pie AND (apple cherry) AND (-blue red)",
           # `*` is a query string like any other: with filters it is Must(AllQuery) AND Must(facet query), +1.0 (times its boost) on
           # every hit (only a blank string drops the text query, search.rs:136); an operator word glued to ')' is a word
           "*", "*^2", "red AND *^3", "(red OR) AND pie", "pie: red"]
FILTERS = [[], ["/namespace/x"], ["/namespace/y", "/namespace/x/organization/o1"], ["/nope"], ["*x*"], ["/namespace/y/*"]]


@pytest.mark.parametrize("filters", FILTERS)
def test_flat_plans_equal_tree_semantics(small, filters):
    """Flattening rules: C++ plan evaluated by the C++ oracle == python tree evaluated exhaustively."""
    ix, desc, ds = small
    for q in QUERIES:
        want, n_match = op.search(ix, q, filters, 0, 10)
        plan = ds.plan(q, filters, 0, 10)
        if plan.is_all:
            continue
        hits, n, cnt = orc.search(desc, batch_from_plans([plan]))
        w = np.zeros(len(want), nat.HIT_DT)
        w["doc"] = [d for d, _ in want]
        w["score"] = [s for _, s in want]
        assert int(cnt[0]) == n_match, (q, filters)
        check_topk(hits[0, :n[0]], w, k=10, ctx=repr((q, filters)))


def test_object_record_facets_and_clamp():
    # get_all_facet_paths (document.rs:277-312): explicit facets win; else namespace facets + metadata (first component only)
    r = ObjectRecord(id="x", text="t", facets=["a/b", "/c"])
    assert all_facet_paths(r) == ["/a/b", "/c"]
    r = ObjectRecord(id="x", text="t", namespace="ns", organization="o", data_type="pdf",
                     metadata={"name": "Doc", "tags": ["u", "v"], "nested": {"k": "val"}, "n": 3})
    assert all_facet_paths(r) == ["/namespace/ns", "/namespace/ns/organization/o", "/namespace/ns/data/pdf",
                                  "/metadata/name", "/metadata/tags", "/metadata/tags", "/metadata/nested"]
    assert r.name() == "Doc"
    # perform_search clamps per_page to 20 when 0 or > 100 (handlers/search.rs:370-374)
    class Fake:
        def search(self, q, f, page, per_page):
            self.pp = per_page
            return []
    fk = Fake()
    assert perform_search({"ns": fk}, "ns", "q", [], 0, 0).per_page == 20 and fk.pp == 20
    assert perform_search({"ns": fk}, "ns", "q", [], 0, 101).per_page == 20
    assert perform_search({"ns": fk}, "ns", "q", [], 0, 100).per_page == 100
    with pytest.raises(KeyError):
        perform_search({}, "missing", "q", [], 0, 10)


def test_fast_path_equals_general_parser(small):
    """fgh_plan_batch's AST-free fast path (ASCII words, optional all-AND) == the general parser."""
    ix, desc, ds = small
    qs = ["red", "red apple", "RED  Apple", " red", "red AND pie", "red AND pie AND cherry", "red AND", "AND red",
          "red OR pie", "red NOT pie", "red AND pie cherry", "red pie AND cherry", "zzz", "zzz AND red", "a" * 45,
          "red-apple", "red_apple", "réd", "red\tpie", "red AND AND pie", "red and pie"]
    batch, status = ds.plan_batch(qs, None, 1, 7)
    for i, q in enumerate(qs):
        try:
            want = ds.plan(q, [], 1, 7).as_dict()
        except nat.FgError as e:
            assert status[i] == e.code, q
            continue
        if "disjunct_of_clause" in want:  # a nested query does not fit the flat batch fgh_plan_batch fills
            assert status[i] == nat.FG_ERR_UNSUPPORTED, q
            continue
        assert status[i] == 0, q
        got_k = int(batch.q["k"][i])
        cb, nc = int(batch.q["clause_begin"][i]), int(batch.q["n_clauses"][i])
        got = [(int(c["occur"]), [(int(l["field"]), int(l["term_ord"]), float(l["boost"]))
                                   for l in batch.l[c["leaf_begin"]:c["leaf_begin"] + c["n_leaves"]]]) for c in batch.c[cb:cb + nc]]
        assert got_k == want["k"] == 14 and got == want["clauses"], q


def test_nested_union_of_boolean_queries_plans_as_disjuncts():
    """`(a AND b) OR (c AND d)`, `a OR (b AND c)`: a union whose children are one-level boolean queries plans as
    disjuncts (fgh_plan_t::n_disjuncts; answered by fg_search_union_of). With facet filters the plan's last child is a
    FILTER child: Must(facet group) AND Must(any positive leaf of the children at boost 0). Deeper nesting or a Must
    sibling of a nested group stay FG_ERR_UNSUPPORTED."""
    from fugu_b200 import _native as nat
    from fugu_b200.dataset import Dataset, ObjectRecord

    ds = Dataset(None)
    ds.upsert([ObjectRecord(id="x", text="alpha beta gamma delta omega", metadata={"name": "alpha"})], commit=False)
    S, M, N = nat.FG_OCCUR_SHOULD, nat.FG_OCCUR_MUST, nat.FG_OCCUR_MUST_NOT
    p = ds.plan("(alpha AND beta) OR (gamma AND delta)").as_dict()
    assert p["disjunct_of_clause"] == [1, 1, 2, 2] and [c[0] for c in p["clauses"]] == [M, M, M, M]
    assert all(len(c[1]) == 2 for c in p["clauses"])  # every word over [text, name]
    p = ds.plan("omega OR (+alpha -beta) gamma").as_dict()
    # plain words of the top level form the first child; then the nested group
    assert p["disjunct_of_clause"] == [1, 1, 2, 2] and [c[0] for c in p["clauses"]] == [S, S, M, N]
    assert "disjunct_of_clause" not in ds.plan("alpha AND beta").as_dict()
    assert "disjunct_of_clause" not in ds.plan("alpha (beta gamma)").as_dict()
    ds.upsert([ObjectRecord(id="y", text="alpha", facets=["/namespace/x"])], commit=False)
    p = ds.plan("(alpha AND beta) OR gamma", ["namespace/x"]).as_dict()
    assert p["filter_child"] and p["disjunct_of_clause"] == [1, 2, 2, 3, 3]
    assert [c[0] for c in p["clauses"]] == [S, M, M, M, M]           # gamma | alpha AND beta | facet AND any-text
    assert [l[0] for l in p["clauses"][3][1]] == [2]                # the facet group (field 2), scored
    # the children's positive leaves that exist in the dictionaries (gamma, alpha, beta in text; alpha in name), boost 0
    assert len(p["clauses"][4][1]) == 4 and all(l[2] == 0.0 for l in p["clauses"][4][1])
    assert not ds.plan("(alpha AND beta) OR gamma").as_dict()["filter_child"]
    for q, f in (("alpha OR (beta AND (gamma OR (delta AND omega)))", []), ("+alpha (beta AND gamma)", []),
                 ("(alpha^-1 AND beta) OR gamma", ["namespace/x"])):
        with pytest.raises(nat.FgError) as e:
            ds.plan(q, f)
        assert e.value.code == nat.FG_ERR_UNSUPPORTED, q
    ds.close()
