"""CPU: the synthetic corpus generator is counter-based, so any doc range
reproduces the same documents; the CSR it emits equals the CSR obtained by tokenising its text."""
import numpy as np

from fugu_b200 import synth
from fugu_b200.dataset import tokenize


def test_shards_reproduce_the_whole():
    cfg = synth.Config(cfg=2, n_docs=3000, vocab=800, n_queries=10, k=10, name_pct=10, n_ns=4)
    c = synth.Corpus.for_config(cfg)
    for field in (0, 1, 2):
        whole = c.csr(0, 3000, field)
        parts = [c.csr(0, 1100, field), c.csr(1100, 3000, field)]
        assert whole["total_num_tokens"] == sum(p["total_num_tokens"] for p in parts)
        nt = len(whole["term_offsets"]) - 1
        for t in range(0, nt, max(1, nt // 97)):
            a = whole["doc_ids"][whole["term_offsets"][t]:whole["term_offsets"][t + 1]]
            b = np.concatenate([p["doc_ids"][p["term_offsets"][t]:p["term_offsets"][t + 1]] + base
                                for p, base in zip(parts, (0, 1100))])
            assert np.array_equal(a, b)
            assert np.all(np.diff(a.astype(np.int64)) > 0)
    c.close()


def test_text_round_trips_through_the_analyzer():
    cfg = synth.Config(cfg=2, n_docs=200, vocab=500, n_queries=10, k=10, name_pct=50)
    c = synth.Corpus.for_config(cfg)
    for field in (0, 1):
        csr = c.csr(0, 200, field)
        counts = {}
        for d in range(200):
            toks = tokenize(c.doc_text(d, field))
            assert len(toks) == csr["doc_len"][d]
            for t in toks:
                assert t[0] == "w"
                counts[(int(t[1:]) - 1, d)] = counts.get((int(t[1:]) - 1, d), 0) + 1
        got = {}
        for t in range(len(csr["term_offsets"]) - 1):
            for i in range(csr["term_offsets"][t], csr["term_offsets"][t + 1]):
                got[(t, int(csr["doc_ids"][i]))] = int(csr["term_freqs"][i])
        assert got == counts
    lens = [c.csr(0, 200, 0)["doc_len"][d] for d in range(200)]
    assert 8 <= min(lens) and max(lens) <= 400
    c.close()


def test_query_generators_follow_the_config_table():
    for cid, chk in ((1, lambda q: q.count(" AND ") == 1), (3, lambda q: " AND " not in q and 2 <= len(q.split()) <= 6),
                     (4, lambda q: q.count(" AND ") == 2)):
        qs = synth.gen_queries(synth.CONFIGS[cid], n=50)
        assert all(chk(q["query"]) for q in qs) and qs == synth.gen_queries(synth.CONFIGS[cid], n=50)
    qs = synth.gen_queries(synth.CONFIGS[2], n=400)
    n_terms = [len(q["query"].replace(" AND ", " ").split()) for q in qs]
    assert set(n_terms) == {1, 2, 3, 4}
    assert all(q["k"] == 10 for q in qs)
    q5 = synth.gen_queries(synth.CONFIGS[5], n=20)
    assert all(1 <= len(q["filters"]) <= 2 for q in q5)
