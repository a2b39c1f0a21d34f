"""Generates the C1 replay kit (BASELINE.json configs[0]: 10 k synthetic docs, 1 k two-term AND queries, top-10):

  c1_replay/ingest.jsonl.gz   one `POST /ingest` body per line: {"data": [ObjectRecord, ...]} (500 records each;
                              /root/reference/src/server/types.rs:83-85, handlers/ingest.rs:14)
  c1_replay/queries.jsonl     one `POST /search` body per line: {"query": "w17 AND w203", "page": {"page": 0, "per_page": 10}}
                              (types.rs:58-68, handlers/search.rs:152)
  c1_replay/expected.jsonl    per query: {"hits": [[id, score], ...], "total_matches": n} from oracle/oracle.cpp

Anyone with `cargo build --release` can start a real fugu 0.1.0, POST the ingest bodies in order, POST the queries and
diff ids (exact, except inside score ties) and scores (1e-5 relative): that pins the oracle -- and through it the
CUDA path -- to tantivy 0.24.1. Ids are single lowercase alphanumeric tokens ("d000000042") so that upsert's
delete-by-id term matches (SURVEY.md A.1 pitfall). Run from the repo root: python tests/golden/make_c1_replay.py"""
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fugu_b200 import _native as nat  # noqa: E402
from fugu_b200 import synth  # noqa: E402
from oracle import orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "c1_replay")


def main():
    cfg = synth.CONFIGS[1]
    corpus = synth.Corpus.for_config(cfg)
    os.makedirs(OUT, exist_ok=True)
    with gzip.GzipFile(os.path.join(OUT, "ingest.jsonl.gz"), "wb", mtime=0) as f:
        for b0 in range(0, cfg.n_docs, 500):
            recs = [{"id": f"d{d:09d}", "text": corpus.doc_text(d)} for d in range(b0, min(b0 + 500, cfg.n_docs))]
            f.write((json.dumps({"data": recs}, separators=(",", ":")) + "\n").encode())
    qs = synth.gen_queries(cfg)
    with open(os.path.join(OUT, "queries.jsonl"), "w") as f:
        for q in qs:
            f.write(json.dumps({"query": q["query"], "page": {"page": 0, "per_page": cfg.k}}) + "\n")
    fields = synth.build_fields(corpus, 0, cfg.n_docs)
    desc = nat.HostIndexDesc(cfg.n_docs, fields)
    batch = synth.lower_queries(qs, vocab=cfg.vocab, n_text_fields=1)
    hits, n, cnt = orc.search(desc, batch, threads=4)
    with open(os.path.join(OUT, "expected.jsonl"), "w") as f:
        for qi in range(len(qs)):
            hl = [[f"d{int(hits['doc'][qi, r]):09d}", float(hits["score"][qi, r])] for r in range(int(n[qi]))]
            f.write(json.dumps({"hits": hl, "total_matches": int(cnt[qi])}) + "\n")
    print("wrote", OUT, {fn: os.path.getsize(os.path.join(OUT, fn)) for fn in sorted(os.listdir(OUT))})


if __name__ == "__main__":
    main()
