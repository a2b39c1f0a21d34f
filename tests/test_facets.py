"""Facet counting (/root/reference/src/db/facet.rs:33-270) and the delete-only
snapshot refresh (row f3). CPU: dictionary enumeration, facet order and the tree builder against the
python twin. GPU: counts (= match counts of single-term queries on the device) against the twin."""
import numpy as np
import pytest

from fugu_b200 import _native as nat
from fugu_b200.dataset import Dataset, FacetNode, ObjectRecord, build_facet_tree
from oracle import oracle_py as op
from tests.util import check_topk, golden, golden_dataset

TRICKY = [
    ("t1", "alpha beta", ["/a/b", "/a-b/c", "/a/b/c d", "/namespace/ns1/organization/o-1"]),
    ("t2", "alpha", ["/a", "/a b", "/zeta", "namespace/ns1/data/pdf"]),
    ("t3", "beta", ["/a/b/c d/e", "/Z", "/namespace/ns2"]),
    ("t4", "gamma", []),
]


def _tricky(ctx=None):
    ds, ix = Dataset(ctx), op.PyIndex()
    ds.upsert([ObjectRecord(id=i, text=t, facets=f) for i, t, f in TRICKY], commit=False)
    for i, t, f in TRICKY:
        ix.upsert(i, t, None, f)
    return ds, ix


def _node_dict(n: FacetNode) -> dict:
    return {"name": n.name, "path": n.path, "count": n.count, "children": {k: _node_dict(v) for k, v in n.children.items()}}


def test_enumeration_matches_twin_order_and_depth():
    ds, ix = _tricky()
    keys = sorted(ix.post[2], key=op._facet_sort_key)
    assert ds.facet_children("/", 0) == keys  # facet order: '/' sorts below every other byte
    assert ds.facet_children("/", 0).index("/a/b") < ds.facet_children("/", 0).index("/a b") < ds.facet_children("/", 0).index("/a-b")
    for root in ["/", "/a", "/a/b", "/namespace", "/namespace/ns1", "/nosuch", "/a/"]:
        want = [p for p, _ in op.facet_collect(ix, root)]  # nothing deleted: every child has a doc
        assert ds.facet_children(root, 1) == want, root
    assert ds.facet_children("/a", 2) == ["/a/b", "/a/b/c d"]
    with pytest.raises(nat.FgError) as e:
        ds.facet_children("a/b")
    assert e.value.code == nat.FG_ERR_INVALID
    ds.close()


def test_counts_need_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ds, _ = _tricky()
    with pytest.raises(nat.FgError) as e:
        ds.facet_counts("/")
    assert e.value.code == nat.FG_ERR_NO_DEVICE  # no CPU fallback: df from the dictionary is never reported as a count
    ds.close()


@pytest.mark.parametrize("max_depth", [None, 0, 1, 2, 3, 5])
def test_tree_builder_matches_twin(max_depth):
    g = golden("search_cases.json")
    ix = op.PyIndex()
    for d in g["docs"]:
        ix.upsert(d["id"], d["text"], d["name"], d["facets"])
    for i, t, f in TRICKY:
        ix.upsert(i, t, None, f)
    for i in g["deletes"]:
        ix.delete(i)
    flat: list = []
    op.facet_collect_recursive(ix, "/", 0, max_depth, flat)
    want = op.facet_tree(ix, max_depth)
    got = build_facet_tree(flat, max_depth)
    assert got.total_facets == want["total_facets"] and got.max_depth == want["max_depth"]
    assert {k: _node_dict(v) for k, v in got.tree.items()} == want["tree"]
    if max_depth is None:  # quirk of update_parent_counts: a parent counts its own docs plus its children's totals
        ns = got.tree["namespace"]
        assert ns.count == dict(op.facet_collect(ix, "/"))["/namespace"] + sum(c.count for c in ns.children.values())


def test_golden_facet_fixture():
    """tests/golden/facet_cases.json (written by make_golden.py from the twin; part of the replay kit: the
    "tree" entries are what GET /facets/tree of a real fugu should answer for the replay corpus)."""
    g = golden("facet_cases.json")
    flat = [(p, c) for p, c in g["walk"]]
    for md in (None, 2, 3):
        want = g["tree"][str(md)]
        sub = [(p, c) for p, c in flat if md is None or p.count("/") <= md]  # what the walk collects down to max_depth
        got = build_facet_tree(sub, md)
        assert got.total_facets == want["total_facets"] and got.max_depth == want["max_depth"], md
        assert {k: _node_dict(v) for k, v in got.tree.items()} == want["tree"], md
    assert [p for p, _ in g["collect"]["/namespace"]] == ["/namespace/ns0", "/namespace/ns1", "/namespace/ns2", "/namespace/ns3"]
    assert g["collect"]["/nosuch"] == []


# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ctx():
    c = nat.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
def test_facet_counts_golden_corpus_with_deletes(ctx):
    g, ds, ix = golden_dataset(ctx)
    ds.commit()
    gf = golden("facet_cases.json")
    for root in ["/", "/namespace", "/namespace/ns0", "/namespace/ns1/organization", "/namespace/ns3/data", "/nosuch"]:
        assert ds.list_facet(root) == op.facet_collect(ix, root) == [tuple(x) for x in gf["collect"][root]], root
    assert ds.facet_counts("/", 0) == [tuple(x) for x in gf["walk"]]
    assert ds.get_available_namespaces() == sorted(p[len("/namespace/"):] for p, _ in op.facet_collect(ix, "/namespace"))
    assert ds.get_namespace_facets("ns2") == op.facet_collect(ix, "/namespace/ns2")
    assert ds.get_facets() == op.facet_collect(ix, "/")
    for md in [None, 1, 2, 3, 4]:
        want = op.facet_tree(ix, md)
        got = ds.get_facet_tree(md)
        assert got.total_facets == want["total_facets"] and got.max_depth == want["max_depth"], md
        assert {k: _node_dict(v) for k, v in got.tree.items()} == want["tree"], md
    paths = ds.get_all_filter_paths()
    assert paths == op.facet_filter_paths(op.facet_tree(ix, None)["tree"]) and "/namespace/ns1/organization" in paths
    ds.close()


def _check_search(ds, ix, q, fl, k=10, tag=""):
    want, n_match = op.search(ix, q, fl, 0, k)
    hits, nh, cnt, st = ds.search_batch([q], [fl], 0, k)
    assert int(st[0]) == 0 and int(cnt[0]) == n_match and int(nh[0]) == len(want), (tag, q, fl)
    o = np.zeros(len(want), nat.HIT_DT)
    o["doc"] = [d for d, _ in want]
    o["score"] = [s for _, s in want]
    check_topk(hits[0, :len(want)], o, k, ctx=f"{tag} {q!r} {fl}")


QUERIES = [("fox", []), ("alpha beta", []), ("alpha AND beta", []), ("", ["/namespace/ns1"]), ("", []),
           ("fox -alpha", ["/namespace/ns0", "/namespace/ns2"]), ("the a of", [])]


@pytest.mark.gpu
def test_delete_only_commit_refreshes_alive_bitset(ctx):
    """Deletes + commit take the fg_index_with_alive route (shared posting arrays, new alive bitset):
    searches and facet counts equal the twin's and a whole re-upload's; statistics keep counting the
    deleted docs (the twin does the same: tantivy's do until a merge)."""
    g, ds, ix = golden_dataset(ctx)
    ds.commit()
    before = ds.index().info()
    victims = [d["id"] for d in g["docs"][5:25:3]] + ["t-none"]
    for v in victims:
        ds.delete(v, commit=False)
        ix.delete(v)
    ds.commit()  # only deletes since the last commit: refresh
    after = ds.index().info()
    assert after.n_postings == before.n_postings and after.n_blocks == before.n_blocks
    ds.commit()  # nothing pending: no-op
    _, full, _ = golden_dataset(ctx)
    for v in victims:
        full.delete(v, commit=False)
    full.commit()  # the same state through a whole upload (first commit of this dataset)
    for q, fl in QUERIES:
        _check_search(ds, ix, q, fl, tag="refresh")
        _check_search(full, ix, q, fl, tag="whole upload")
    flat: list = []
    op.facet_collect_recursive(ix, "/", 0, None, flat)
    assert ds.facet_counts("/", 0) == flat
    assert full.facet_counts("/", 0) == flat
    # a second refresh on top of a refreshed snapshot, then an append (whole rebuild) on top of that
    ds.delete(g["docs"][0]["id"], commit=True)
    ix.delete(g["docs"][0]["id"])
    assert ds.facet_counts("/namespace", 1) == op.facet_collect(ix, "/namespace")
    _check_search(ds, ix, "fox", [], tag="second refresh")
    ds.upsert([ObjectRecord(id="fresh1", text="fox fox alpha", facets=["/namespace/ns9/data/new"])])
    ix.upsert("fresh1", "fox fox alpha", None, ["/namespace/ns9/data/new"])
    assert ds.facet_counts("/namespace", 1) == op.facet_collect(ix, "/namespace")
    for q, fl in QUERIES:
        _check_search(ds, ix, q, fl, tag="append after refresh")
    ds.close()
    full.close()


@pytest.mark.gpu
def test_with_alive_shares_arrays_and_outlives_its_base(ctx):
    from fugu_b200 import synth
    from tests.util import check_batch_against_oracle, plan_queries

    cfg = synth.Config(cfg=2, n_docs=30_000, vocab=4_000, n_queries=64, k=10, name_pct=10)
    corpus = synth.Corpus.for_config(cfg)
    fields = synth.build_fields(corpus, 0, cfg.n_docs)
    base = nat.Index(ctx, nat.HostIndexDesc(cfg.n_docs, fields))
    rng = np.random.default_rng(5)
    words = (cfg.n_docs + 31) // 32
    alive = (rng.integers(0, 2**32, words, dtype=np.uint64) | rng.integers(0, 2**32, words, dtype=np.uint64)).astype(np.uint32)  # ~75 % alive
    derived = base.with_alive(alive)
    again = derived.with_alive(None)  # all alive again, derived from a derived snapshot
    base.close()  # the shared arrays must survive their first owner
    batch = plan_queries(synth.gen_queries(cfg), vocab=cfg.vocab, n_text_fields=2)
    check_batch_against_oracle(derived, nat.HostIndexDesc(cfg.n_docs, fields, alive_bitset=alive), batch)
    derived.close()
    check_batch_against_oracle(again, nat.HostIndexDesc(cfg.n_docs, fields), batch)
    again.close()


@pytest.mark.gpu
def test_facet_counts_large_synthetic_with_facet_columns(ctx):
    """Config-5 corpus (src/object.rs:81-111 namespace / organization / data facets with ancestor
    terms): some facets are frequent enough for dense tf columns, most are not; the count must equal
    the alive docs of each posting list either way."""
    from fugu_b200 import synth

    ds, want, paths = _synthetic_facets(ctx)
    assert ds.index().info().n_columns > 0
    got = ds.facet_counts("/", 0)
    assert dict(got) == want and len(got) == len(want)
    assert [p for p, _ in got] == sorted(want, key=op._facet_sort_key)
    top = ds.list_facet("/")
    assert [p for p, _ in top] == sorted({"/" + p.split("/")[1] for p in want}, key=op._facet_sort_key)
    for p, c in top:
        assert c == want[p]
    ds.close()


def _synthetic_facets(ctx):
    from fugu_b200 import synth

    cfg = synth.Config(cfg=5, n_docs=40_000, vocab=4_000, n_queries=8, k=10, n_ns=64)
    corpus = synth.Corpus.for_config(cfg)
    fields = synth.build_fields(corpus, 0, cfg.n_docs)
    assert len(fields) == 3
    rng = np.random.default_rng(11)
    alive_docs = rng.random(cfg.n_docs) < 0.9
    bits = np.zeros((cfg.n_docs + 31) // 32, np.uint32)
    idx = np.nonzero(alive_docs)[0]
    np.bitwise_or.at(bits, idx >> 5, np.uint32(1) << (idx & 31).astype(np.uint32))
    desc = nat.HostIndexDesc(cfg.n_docs, fields, alive_bitset=bits)
    ds = Dataset(ctx)
    paths = [corpus.facet_path(i) for i in range(corpus.facet_vocab())]
    ds.adopt(desc, [[f"w{i + 1}" for i in range(cfg.vocab)], [], paths])
    offs, docs = fields[2]["term_offsets"], fields[2]["doc_ids"]
    want = {}
    for t, p in enumerate(paths):
        c = int(alive_docs[docs[int(offs[t]):int(offs[t + 1])]].sum())
        if c:
            want[p] = c
    return ds, want, paths


def test_synthetic_facet_dictionary_enumerates_on_cpu():
    ds, want, paths = _synthetic_facets(None)
    assert ds.facet_children("/", 0) == sorted(paths, key=op._facet_sort_key)
    assert set(want) <= set(paths) and len(want) > 64
    ds.close()


@pytest.mark.gpu
def test_uncommitted_documents_are_invisible(ctx):
    """A searcher only sees committed segments (src/db/document.rs:65, core.rs:86): a document upserted
    but not committed changes no result, and a term that only it contains is an empty scorer, not an
    error (its dictionary ordinal is beyond the snapshot's); after the commit both are visible."""
    ds, ix = Dataset(ctx), op.PyIndex()
    base = [("a", "alpha beta", ["/namespace/n1"]), ("b", "beta gamma", ["/namespace/n1"]), ("c", "alpha alpha", ["/namespace/n2"])]
    ds.upsert([ObjectRecord(id=i, text=t, facets=f) for i, t, f in base])
    for i, t, f in base:
        ix.upsert(i, t, None, f)
    ds.upsert([ObjectRecord(id="z", text="zeta alpha", facets=["/namespace/n9"])], commit=False)
    assert ds.search("zeta") == [] and ds.search("zeta", ["/namespace/n9"]) == []
    for q, fl in [("alpha", []), ("alpha zeta", []), ("beta", ["/namespace/n1"]), ("", [])]:
        want, n_match = op.search(ix, q, fl, 0, 10)  # the twin has not seen "z"
        got = ds.search(q, fl, 0, 10)
        assert [r.doc for r in got] == [d for d, _ in want], (q, fl)
    assert ds.facet_counts("/namespace", 1) == op.facet_collect(ix, "/namespace")
    assert ds.plan("zeta").as_dict()["clauses"][0][1][0][1] == nat.FG_TERM_MISSING
    ds.commit()
    ix.upsert("z", "zeta alpha", None, ["/namespace/n9"])
    assert [r.id for r in ds.search("zeta")] == ["z"]
    _check_search(ds, ix, "alpha", [], tag="after commit")
    assert ds.facet_counts("/namespace", 1) == op.facet_collect(ix, "/namespace")
    ds.close()
