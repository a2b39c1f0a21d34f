"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the oracle."""
import numpy as np
import pytest

from fugu_b200 import _native as nat
from fugu_b200 import synth
from tests.util import check_batch_against_oracle, plan_queries

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = nat.Context(0)
    yield c
    c.close()


def _setup(ctx, cfg, n_docs=None, vocab=None):
    n_docs = n_docs or cfg.n_docs
    vocab = vocab or cfg.vocab
    corpus = synth.Corpus.for_config(cfg, vocab=vocab)
    fields = synth.build_fields(corpus, 0, n_docs)
    desc = nat.HostIndexDesc(n_docs, fields)
    return corpus, desc, nat.Index(ctx, desc)


def test_config1_two_term_and(ctx):
    """BASELINE config 0: 10k docs, 1k two-term AND queries, top-10."""
    cfg = synth.CONFIGS[1]
    corpus, desc, index = _setup(ctx, cfg)
    batch = plan_queries(synth.gen_queries(cfg), vocab=cfg.vocab, n_text_fields=1)
    check_batch_against_oracle(index, desc, batch)
    index.close()


def test_config2_small_mixed(ctx):
    """Config 2 shape at 50k docs: mixed 1-4 term AND/OR, two-field expansion (text + name)."""
    cfg = synth.Config(cfg=2, n_docs=50_000, vocab=20_000, n_queries=500, k=10, name_pct=10)
    corpus, desc, index = _setup(ctx, cfg)
    batch = plan_queries(synth.gen_queries(cfg), vocab=cfg.vocab, n_text_fields=2)
    check_batch_against_oracle(index, desc, batch)
    index.close()


def test_config3_small_or_top100(ctx):
    """Config 3 shape at 100k docs: stop-word-heavy OR of 2-6 terms, top-100 (KS=4 queue)."""
    cfg = synth.Config(cfg=3, n_docs=100_000, vocab=20_000, n_queries=100, k=100)
    corpus, desc, index = _setup(ctx, cfg)
    batch = plan_queries(synth.gen_queries(cfg), vocab=cfg.vocab, n_text_fields=1)
    check_batch_against_oracle(index, desc, batch)
    index.close()
