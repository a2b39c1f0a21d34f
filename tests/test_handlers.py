"""The reference's search front-ends (src/server/handlers/search.rs:27-301,350-402) mirrored in
fugu_b200/handlers.py on top of the GPU Dataset: JSON shapes, defaults, namespace selection, per_page clamp,
"text" stripping, hydration (id, score, text, metadata, facets: src/db/search.rs:20-27,534-590) and
ObjectRecord::validate (src/object.rs:31-78). -m gpu; the same functions run under the SIMT emulation on CPU."""
import pytest

from fugu_b200 import _native as nat
from fugu_b200 import handlers as H
from fugu_b200.dataset import Dataset, ObjectRecord, all_facet_paths

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = nat.Context(0)
    yield c
    c.close()


def _state(ctx):
    main, other = Dataset(ctx), Dataset(ctx)
    recs = [ObjectRecord(id=f"doc{i}", text=f"alpha beta gamma{i % 3} filler{i}", metadata={"name": f"report {i}", "kind": "memo"},
                         namespace="acme", organization=f"org{i % 2}", data_type="note") for i in range(30)]
    recs.append(ObjectRecord(id="plain", text="alpha only here", facets=["/custom/path", "nolead/slash"]))
    main.upsert(recs, commit=True)
    other.upsert([ObjectRecord(id="o1", text="alpha in the other namespace")], commit=True)
    return H.AppState(datasets={"fugu_db": main, "second": other}, default_namespace="fugu_db"), main, other


def test_search_endpoint_shape_defaults_and_hydration(ctx):
    state, main, other = _state(ctx)
    code, out = H.search_endpoint(state, {"query": "alpha"})
    assert code == 200 and list(out) == ["status", "query", "filters", "page", "per_page", "total", "results"]
    assert out["status"] == "success" and out["page"] == 0 and out["per_page"] == 20 and out["filters"] == []
    assert out["total"] == len(out["results"]) == 20  # 31 docs match, default per_page 20, no clamp in this handler
    r = out["results"][0]
    assert list(r) == ["id", "score", "text", "metadata", "facets"] and isinstance(r["score"], float)
    byid = {x["id"]: x for x in out["results"]}
    assert byid["plain"]["text"] == "alpha only here" and byid["plain"]["metadata"] is None
    assert byid["plain"]["facets"] == ["/custom/path", "/nolead/slash"]  # explicit facets, normalised (document.rs:277-312)
    some = next(x for x in out["results"] if x["id"].startswith("doc"))
    assert some["metadata"]["kind"] == "memo" and "/namespace/acme" in some["facets"] and some["text"].startswith("alpha beta")
    # pagination object, filters (facet Must group), page 1
    code, out = H.search_endpoint(state, {"query": "alpha", "filters": ["/namespace/acme/organization/org1"], "page": {"page": 1, "per_page": 5}})
    assert code == 200 and out["page"] == 1 and out["per_page"] == 5 and out["total"] == 5
    assert all("/namespace/acme/organization/org1" in x["facets"] for x in out["results"])
    # per_page = 0 is an error here (TopDocs::with_limit(0) panics in the reference): HTTP 500 shape
    code, out = H.search_endpoint(state, {"query": "alpha", "page": {"per_page": 0}})
    assert code == 500 and out["status"] == "error" and out["error"].startswith("Search failed:")
    # POST /search only ever searches the default dataset (handlers/search.rs:169)
    assert all(x["id"] != "o1" for x in H.search_endpoint(state, {"query": "alpha", "page": {"per_page": 100}})[1]["results"])
    state.default_namespace = "missing"
    assert H.search_endpoint(state, {"query": "alpha"}) == (500, {"status": "error", "error": "Default dataset not found"})
    main.close(); other.close()


def test_query_json_post_namespace_text_flags_and_clamp(ctx):
    state, main, other = _state(ctx)
    code, out = H.query_json_post(state, {"query": "alpha"})
    assert code == 200 and list(out) == ["results", "total", "page", "per_page", "query", "includes_data_objects",
                                         "targeting_conversations_or_organizations"]
    assert out["per_page"] == 20 and out["total"] == 20 and out["query"] == "alpha"
    assert all("text" not in r for r in out["results"])  # stripped by default (:264-272)
    assert out["includes_data_objects"] is True and out["targeting_conversations_or_organizations"] is False
    # namespace from the body (:256-257)
    code, out = H.query_json_post(state, {"query": "alpha", "namespace": "second", "text": True})
    assert code == 200 and [r["id"] for r in out["results"]] == ["o1"] and out["results"][0]["text"].startswith("alpha in")
    # url flag wins, disagreement is reported (:221-236)
    code, out = H.query_json_post(state, {"query": "alpha", "text": True}, url_text=False)
    assert "developer_message" in out and all("text" not in r for r in out["results"])
    # perform_search clamp: per_page 0 or > 100 -> 20 (:370-374)
    for pp in (0, 101, 5000):
        code, out = H.query_json_post(state, {"query": "alpha", "page": {"per_page": pp}})
        assert code == 200 and out["per_page"] == 20
    code, out = H.query_json_post(state, {"query": "alpha", "page": {"per_page": 100}})
    assert out["per_page"] == 100 and out["total"] == 31
    # conversation / organization filters flip include_data's default (:241-248)
    code, out = H.query_json_post(state, {"query": "alpha", "filters": ["namespace/acme/organization/org0"]})
    assert out["targeting_conversations_or_organizations"] is True and out["includes_data_objects"] is False and out["total"] == 15
    code, out = H.query_json_post(state, {"query": "alpha", "namespace": "nope"})
    assert code == 500 and out == {"error": "Search failed: Namespace 'nope' not found"}
    main.close(); other.close()


def test_get_front_ends(ctx):
    state, main, other = _state(ctx)
    code, out = H.query_text_get(state, "alpha", limit=7)
    assert code == 200 and out["per_page"] == 7 and out["total"] == 7 and all("text" not in r for r in out["results"])
    code, out = H.query_text_get(state, "alpha", text=True, namespace="second")
    assert [r["id"] for r in out["results"]] == ["o1"] and "text" in out["results"][0]
    code, out = H.query_text_path(state, "alpha%20AND%20gamma1")
    assert code == 200 and out["query"] == "alpha AND gamma1" and out["total"] == 10 and out["per_page"] == 20
    assert H.query_text_path(state, "%ff%fe")[0] == 400
    main.close(); other.close()


def test_object_record_validate_messages(ctx):
    ok = ObjectRecord(id="a", text="t")
    ok.validate()
    for rec, msg in [(ObjectRecord(id="", text="t"), "Object ID cannot be empty"),
                     (ObjectRecord(id="x" * 257, text="t"), "Object ID too long (max 256 characters)"),
                     (ObjectRecord(id="a", text=""), "Object text cannot be empty"),
                     (ObjectRecord(id="a", text="y" * 10001), "Text too long (max 10000 characters)"),
                     (ObjectRecord(id="a", text="t", namespace="a b"), "Invalid namespace format"),
                     (ObjectRecord(id="a", text="t", namespace="n" * 129), "Namespace too long (max 128 characters)"),
                     (ObjectRecord(id="a", text="t", facets=["f"] * 101), "Too many facets (max 100 per object)"),
                     (ObjectRecord(id="a", text="t", facets=["ok", ""]), "Facet at index 1 cannot be empty"),
                     (ObjectRecord(id="a", text="t", facets=["z" * 513]), "Facet at index 0 too long (max 512 characters)")]:
        with pytest.raises(ValueError) as e:
            rec.validate()
        assert str(e.value) == msg


def test_micro_batcher_concurrent_single_query_requests(ctx):
    """Row f2: many threads, one query each (the HTTP API's shape, handlers/search.rs:152), answered through the
    micro-batcher: every caller gets exactly what a direct Dataset.search returns, requests are really batched,
    a request the device path does not take fails alone."""
    import threading

    from fugu_b200.dataset import Batcher

    state, main, other = _state(ctx)
    queries = [("alpha", [], 0, 20), ("beta AND gamma1", [], 0, 5), ("report", ["namespace/acme"], 0, 10), ("filler7 OR filler9", [], 0, 20),
               ("alpha", [], 1, 10), ("nosuchterm", [], 0, 20), ("alpha beta", ["/namespace/acme/organization/org1"], 0, 7)]
    want = [[(r.id, r.score) for r in main.search(q, f, p, pp)] for q, f, p, pp in queries]
    b = Batcher(main, max_batch=64, max_wait_us=20000)
    n_threads, got, errs = 24, {}, {}
    start = threading.Barrier(n_threads)

    def worker(t):
        q, f, p, pp = queries[t % len(queries)]
        start.wait()
        try:
            got[t] = [(r.id, r.score) for r in b.search(q, f, p, pp)]
            if t == 3:  # a phrase query is not evaluated on the device: this caller alone gets the error
                try:
                    b.search('"alpha beta"')
                    errs[t] = "phrase query was answered"
                except nat.FgError as e:
                    if e.code != nat.FG_ERR_UNSUPPORTED:
                        errs[t] = f"wrong code {e.code}"
        except Exception as e:  # noqa: BLE001
            errs[t] = repr(e)

    th = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    st = b.stats()
    b.close()
    assert not errs, errs
    for t in range(n_threads):
        assert got[t] == want[t % len(queries)], (t, queries[t % len(queries)])
    assert st["n_requests"] == n_threads + 1 and st["n_batches"] < st["n_requests"] and st["max_batch_seen"] >= 2, st
    main.close(); other.close()


def test_nested_boolean_queries_through_dataset_search(ctx):
    """`(a AND b) OR (c AND d)` and friends through Dataset.search / the batched call / the micro-batcher: planned as a
    union of boolean queries and answered by fg_search_union_of. Expected: a document matches when any child matches and
    scores the sum of the children it matches -- recomputed here from complete result lists of the children searched on
    their own (the children themselves are checked against the oracle elsewhere)."""
    from fugu_b200.dataset import Batcher, QuerySet

    state, main, other = _state(ctx)
    every = 1000  # more than the corpus: complete result lists

    def expected(children, page, per_page):
        tot = {}
        for ch in children:
            for r in main.search(ch, [], 0, every):
                tot[r.id] = tot.get(r.id, 0.0) + r.score
        docid = {r.id: r.doc for ch in children for r in main.search(ch, [], 0, every)}
        ranked = sorted(tot.items(), key=lambda kv: (-kv[1], docid[kv[0]]))
        return ranked[page * per_page:(page + 1) * per_page]

    cases = [("(alpha AND gamma1) OR (beta AND filler7)", ["alpha AND gamma1", "beta AND filler7"]),
             ("filler3 OR (alpha AND gamma2)", ["filler3", "alpha AND gamma2"]),
             ("(report AND gamma0) only filler9 OR (+alpha -beta)", ["only filler9", "report AND gamma0", "+alpha -beta"]),
             ("(nosuchterm AND alpha) OR (beta AND nosuchterm)", ["nosuchterm AND alpha", "beta AND nosuchterm"])]
    for q, children in cases:
        assert "disjunct_of_clause" in main.plan(q).as_dict(), q
        for page, pp in ((0, 20), (1, 7), (0, 100)):
            got = [(r.id, r.score) for r in main.search(q, [], page, pp)]
            want = expected(children, page, pp)
            assert [g[0] for g in got] == [w[0] for w in want], (q, page, pp)
            for (_, gs), (_, ws) in zip(got, want):
                assert abs(gs - ws) <= 1e-5 * max(abs(ws), 1e-30), (q, gs, ws)
    # with facet filters: Bool[Must(nested text query), Must(facet query)] (src/db/search.rs:140-144) -- the facet group must
    # hold for every hit and scores once (a filter child of the union). Expected from the Python twin's tree evaluation.
    from oracle import oracle_py as op

    ix = op.PyIndex()
    for i in range(30):
        r = ObjectRecord(id=f"doc{i}", text=f"alpha beta gamma{i % 3} filler{i}", metadata={"name": f"report {i}", "kind": "memo"},
                         namespace="acme", organization=f"org{i % 2}", data_type="note")
        ix.upsert(r.id, r.text, r.name(), all_facet_paths(r))
    plain = ObjectRecord(id="plain", text="alpha only here", facets=["/custom/path", "nolead/slash"])
    ix.upsert(plain.id, plain.text, plain.name(), all_facet_paths(plain))
    for q, _ in cases:
        for fl in (["/namespace/acme/organization/org1"], ["/namespace/acme/organization/org0", "/custom/path"], ["/nowhere"]):
            assert main.plan(q, fl).as_dict().get("filter_child"), (q, fl)
            for page, pp in ((0, 20), (1, 4)):
                got = main.search(q, fl, page, pp)
                want, n_match = op.search(ix, q, fl, page, pp)
                assert [r.id for r in got] == [ix.ids[d] for d, _ in want], (q, fl, page, pp)
                for r, (_, ws) in zip(got, want):
                    assert abs(r.score - ws) <= 1e-5 * max(abs(ws), 1e-30), (q, fl, r.score, ws)
            h, n, c, st = main.search_batch([q], [fl], 0, 20, want_counts=True)
            assert st[0] == 0 and int(c[0]) == op.search(ix, q, fl, 0, 20)[1], (q, fl)
    # in a batch next to ordinary queries, and through the micro-batcher
    qs = ["alpha", cases[0][0], "beta AND gamma1", cases[1][0]]
    hits, nh, cnt, status = main.search_batch(QuerySet(qs, None, 0, 20), want_counts=False)
    assert (status == 0).all()
    for i, q in enumerate(qs):
        one = main.search(q, [], 0, 20)
        assert nh[i] == len(one) and [int(h["doc"]) for h in hits[i, :nh[i]]] == [r.doc for r in one], q
    b = Batcher(main, max_batch=16, max_wait_us=1000)
    assert [r.id for r in b.search(cases[0][0], [], 0, 20)] == [r.id for r in main.search(cases[0][0], [], 0, 20)]
    b.close()
    main.close(); other.close()
