"""Generates tests/golden/*.json from the pure-Python twin (oracle/oracle_py.py) and from values
derived by hand from the published formulas. The reference itself cannot run
here (no Rust toolchain, tantivy un-vendored), so these are the known-answer vectors that pin the
C++ oracle and the GPU path; they double as the replay kit for a future `cargo` box:
corpus.jsonl bodies are POST /ingest payloads, queries are POST /search bodies
(/root/reference/src/server/types.rs:58-68,83-85).

    python tests/golden/make_golden.py        # rewrites the fixtures in place
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle_py as op  # noqa: E402

WORDS = ["alpha", "beta", "gamma", "delta", "epsilon", "zeta", "eta", "theta", "iota", "kappa", "lambda", "mu",
         "fox", "dog", "cat", "quick", "lazy", "brown", "the", "a", "and", "of", "report", "invoice", "q3"]


def corpus(seed=7, n=60):
    rng = random.Random(seed)
    docs = []
    for i in range(n):
        ln = rng.randint(3, 40)
        # skewed word choice so that a few words are in most docs
        text = " ".join(WORDS[min(int(rng.expovariate(0.25)), len(WORDS) - 1)] for _ in range(ln))
        name = " ".join(rng.choice(WORDS) for _ in range(rng.randint(1, 3))) if rng.random() < 0.3 else None
        ns = f"ns{rng.randint(0, 3)}"
        facets = [f"/namespace/{ns}", f"/namespace/{ns}/organization/org{rng.randint(0, 2)}"]
        if rng.random() < 0.5:
            facets.append(f"/namespace/{ns}/data/type{rng.randint(0, 1)}")
        docs.append({"id": f"d{i:09d}", "text": text, "name": name, "facets": facets})
    # two upserts of existing ids (delete + append) and one plain delete
    docs.append({"id": "d000000003", "text": "fox fox fox quick report", "name": "fox", "facets": ["/namespace/ns0"]})
    docs.append({"id": "d000000010", "text": "alpha beta", "name": None, "facets": ["/namespace/ns1"]})
    return docs, ["d000000020"]


QUERIES = [
    ("fox", []), ("alpha beta", []), ("alpha AND beta", []), ("alpha AND beta AND gamma", []),
    ("+alpha beta -gamma", []), ("alpha OR fox", []), ("NOT alpha", []), ("alpha -beta", []),
    ("fox^2 dog", []), ("text:fox", []), ("name:fox", []), ("(alpha beta) AND gamma", []),
    ("alpha AND (beta OR gamma)", []), ("nosuchword", []), ("alpha nosuchword", []), ("alpha AND nosuchword", []),
    ("", []), ("*", []), ("fox", ["/namespace/ns0"]), ("alpha beta", ["/namespace/ns1", "namespace/ns2"]),
    ("alpha AND beta", ["/namespace/ns0/organization/org1"]), ("", ["/namespace/ns3"]),
    ("fox", ["/namespace/ns0/*"]), ("fox", ["/namespace/ns1=whatever"]), ("fox", ["*ns*"]),
    ("alpha", ["/does/not/exist"]), ("fox dog:", []), ("(alpha", []), ("alpha AND", []),
    ("The QUICK, brown fox!", []), ("alpha beta gamma delta epsilon zeta", []), ("a the and of", []),
]


def main():
    docs, deletes = corpus()
    ix = op.PyIndex()
    for d in docs:
        ix.upsert(d["id"], d["text"], d["name"], d["facets"])
    for i in deletes:
        ix.delete(i)
    cases = []
    for q, fl in QUERIES:
        for page, per_page in ((0, 10), (1, 3)):
            try:
                hits, n = op.search(ix, q, fl, page, per_page)
                cases.append({"query": q, "filters": fl, "page": page, "per_page": per_page, "match_count": n,
                              "hits": [[ix.ids[d], s] for d, s in hits]})
            except op.Unsupported as e:
                cases.append({"query": q, "filters": fl, "page": page, "per_page": per_page, "error": "unsupported"})
            except op.ParseError:
                cases.append({"query": q, "filters": fl, "page": page, "per_page": per_page, "error": "invalid"})
    json.dump({"docs": docs, "deletes": deletes, "cases": cases}, open(os.path.join(HERE, "search_cases.json"), "w"), indent=1)

    # formula-level known answers (hand-derivable)
    fn = {"table_spot": {"0": 0, "23": 23, "24": 24, "40": 40, "41": 42, "47": 54, "48": 56, "49": 60, "255": 2013265944},
          "to_id": {"0": 0, "1": 1, "40": 40, "41": 40, "42": 41, "43": 41, "60": 49, "400": op.fieldnorm_to_id(400),
                    "10000": op.fieldnorm_to_id(10000), "4294967295": 255}}
    bm = []
    for df, n, tot, fnid, tf, boost in [(1, 3, 16, 4, 1, 1.0), (2, 3, 16, 4, 1, 1.0), (5, 1000, 60000, 49, 3, 1.0),
                                        (974000, 1000000, 68000000, 52, 7, 1.0), (1000, 1000, 60000, 40, 1, 1.0),
                                        (17, 100000, 6800000, 60, 2, 2.5), (3, 64, 384, 1, 1, 1.0)]:
        f32 = op.f32
        avg = f32(tot) / f32(n)
        w = f32(boost) * (op.idf(df, n) * (f32(1) + op.K1))
        norm = op.K1 * (f32(1) - op.B + op.B * f32(op.FIELDNORM_TABLE[fnid]) / avg)
        bm.append({"df": df, "n_docs": n, "total_tokens": tot, "fieldnorm_id": fnid, "tf": tf, "boost": boost,
                   "idf": float(op.idf(df, n)), "score": float(w * (f32(tf) / (f32(tf) + norm)))})
    tok = [[s, op.tokenize(s)] for s in ["Hello, World!", "foo-bar_baz", "x" * 39 + " " + "y" * 40, "ÀÉÎõü straße İstanbul",
                                         "d000000042 w17", "tab\tnew\nline", "日本語 テキスト 123", "a.b.c@d.com", "", "   ",
                                         # Other_Alphabetic marks (Rust's char::is_alphabetic counts them): Devanagari vowel signs stay, the
                                         # virama splits; a combining acute splits; Arabic harakat stay; a Thai tone mark splits; Nl / No digits
                                         "हिन्दी भाषा", "cafe\u0301 naïve", "كِتَاب جديد", "น้ำ ภาษาไทย", "Ⅻ ½ x²"]]
    json.dump({"fieldnorm": fn, "bm25": bm, "tokenizer": tok}, open(os.path.join(HERE, "formulas.json"), "w"), indent=1, ensure_ascii=False)
    # replay kit for a real fugu (POST /ingest body + POST /search bodies)
    with open(os.path.join(HERE, "replay_ingest.json"), "w") as f:
        json.dump({"data": [{"id": d["id"], "text": d["text"], "metadata": ({"name": d["name"]} if d["name"] else None),
                             "facets": d["facets"]} for d in docs]}, f)
    with open(os.path.join(HERE, "replay_queries.jsonl"), "w") as f:
        for c in cases:
            f.write(json.dumps({"query": c["query"], "filters": c["filters"], "page": {"page": c["page"], "per_page": c["per_page"]}}) + "\n")
    # facet counting (src/db/facet.rs): what FacetCollector over AllQuery reports for a few roots, the
    # recursive walk of get_facet_tree and the response of GET /facets/tree for three depths
    roots = ["/", "/namespace", "/namespace/ns0", "/namespace/ns1/organization", "/namespace/ns3/data", "/nosuch"]
    walk = []
    op.facet_collect_recursive(ix, "/", 0, None, walk)
    json.dump({"collect": {r: op.facet_collect(ix, r) for r in roots}, "walk": walk,
               "tree": {str(md): op.facet_tree(ix, md) for md in (None, 2, 3)}},
              open(os.path.join(HERE, "facet_cases.json"), "w"), indent=1)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
