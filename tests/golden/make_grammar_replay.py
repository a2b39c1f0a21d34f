"""Generates the grammar replay kit: a 400-document corpus (text, metadata.name, facets; some documents deleted) and
360 query strings from the differential fuzz's generator -- bare words, field prefixes, boosts, + / -, AND / OR /
AND NOT, parenthesised groups two levels deep, `*`, stray punctuation and operator words that send QueryParser into
fugu's escape-and-retry fallback (/root/reference/src/db/search.rs:118-127, 603-610) -- with facet filters and pages:

  grammar_replay/ingest.json      one `POST /ingest` body: {"data": [ObjectRecord, ...]} (types.rs:83-85)
  grammar_replay/deletes.json     ids to DELETE afterwards (DocumentOperations::delete_document, document.rs:70-97)
  grammar_replay/cases.jsonl      per case: {"body": <POST /search body>, "hits": [[id, score], ...]} or {"body", "error": ..}
                                  from the Python twin (oracle/oracle_py.py), which evaluates the parse tree directly

C1's replay kit pins the arithmetic; this one pins the GRAMMAR and the planner: which strings parse, which fall back,
how groups, occurs and boosts nest, what `*` and facet filters add. The restatements' readings of tantivy-query-grammar
0.24 that it would confirm or refute are listed in DESIGN.md section 6. Requests the twin refuses as outside the device
path's scope (phrases, ranges, fuzzy terms) are left out. Run from the repo root: python tests/golden/make_grammar_replay.py"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as op  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "grammar_replay")
V = 40


def build(rng):
    def word():
        return f"w{min(int(rng.paretovariate(0.9)), V)}"

    ix, docs = op.PyIndex(), []
    for i in range(400):
        text = " ".join(word() for _ in range(rng.randint(3, 40)))
        name = " ".join(word() for _ in range(rng.randint(1, 4))) if rng.random() < 0.3 else None
        facets = [f"/ns/n{i % 5}"] + ([f"/kind/k{i % 3}/sub{i % 2}"] if i % 4 == 0 else [])
        docs.append({"id": f"d{i}", "text": text, "metadata": {"name": name} if name else None, "facets": facets})
        ix.upsert(f"d{i}", text, name, facets)
    deletes = [f"d{i}" for i in range(0, 400, 17)]
    for i in deletes:
        ix.delete(i)

    def term():
        t = word()
        if rng.random() < 0.05:
            t = rng.choice(["*", "W3", "w2.", "zzz", "w1,", "(w2", "w3)", "w4:", "w5^", "AND", "OR", "NOT"])
        r = rng.random()
        if r < 0.1:
            t = "text:" + t
        elif r < 0.2:
            t = "name:" + t
        if rng.random() < 0.15:
            t += rng.choice(["^2", "^0.5", "^3.5"])
        return t

    def group(depth):
        parts = []
        for _ in range(rng.randint(1, 4)):
            if depth < 2 and rng.random() < 0.3:
                p = "(" + group(depth + 1) + ")" + rng.choice(["", "", "^2", "^0.5"])
            else:
                p = term()
            r = rng.random()
            if r < 0.12:
                p = "+" + p
            elif r < 0.2:
                p = "-" + p
            parts.append(p)
        j = rng.random()
        if j < 0.3:
            return " AND ".join(parts)
        if j < 0.45:
            return " OR ".join(parts)
        if j < 0.5 and len(parts) > 1:
            return parts[0] + " AND NOT " + " ".join(parts[1:])
        return " ".join(parts)

    cases = []
    while len(cases) < 360:
        q = group(0)
        if rng.random() < 0.03:
            q = rng.choice(["", "  ", "*", "* w1", "w1 AND *", "*^2"])
        fl = rng.choice([[], [], [], ["/ns/n1"], ["/ns/n2", "/kind/k0/*"], ["*x*"], ["/nope"], ["ns/n3"], ["kind=k1"]])
        page, pp = rng.choice([(0, 10), (0, 20), (1, 5), (0, 100), (3, 7), (11, 100)])
        body = {"query": q, "filters": fl, "page": {"page": page, "per_page": pp}}
        try:
            hits, _ = op.search(ix, q, fl, page, pp)
        except op.Unsupported:
            continue
        except op.ParseError as e:
            cases.append({"body": body, "error": f"parse error after the escaped retry ({e}): the reference answers HTTP 500"})
            continue
        cases.append({"body": body, "hits": [[ix.ids[d], s] for d, s in hits]})
    return docs, deletes, cases


def main():
    docs, deletes, cases = build(random.Random(20261019))
    os.makedirs(OUT, exist_ok=True)
    json.dump({"data": docs}, open(os.path.join(OUT, "ingest.json"), "w"), separators=(",", ":"))
    json.dump(deletes, open(os.path.join(OUT, "deletes.json"), "w"))
    with open(os.path.join(OUT, "cases.jsonl"), "w") as f:
        for c in cases:
            f.write(json.dumps(c, ensure_ascii=False) + "\n")
    print("wrote", OUT, {fn: os.path.getsize(os.path.join(OUT, fn)) for fn in sorted(os.listdir(OUT))}, "cases", len(cases),
          "errors", sum("error" in c for c in cases))


if __name__ == "__main__":
    main()
