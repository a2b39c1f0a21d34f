"""Dev tool (no GPU needed): runs a C2-shaped sample through the kernel sources under the SIMT emulation
(tests/emu) with the column scan's sparse-hit gate on and off, checks both against the oracle and prints
the chunk counters (profiles/r01d_cpu_only_session.md quotes them). FGEMU_PROFILE=1 adds host time per launch.

    python tools/emu_gating_counters.py [n_docs=100000] [n_queries=300]
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from fugu_b200 import _native as nat  # noqa: E402

subprocess.check_call(["make", "-s", "-j4", "-C", os.path.join(ROOT, "tests", "emu")])
nat.LIB_PATH = os.path.join(ROOT, "tests", "emu", "libfugu_emu.so")
from tests import util  # noqa: E402

util.EMULATED = True
from fugu_b200 import synth  # noqa: E402
from oracle import orc  # noqa: E402
from tests.util import check_topk, plan_queries  # noqa: E402

nd = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 300
cfg = synth.Config(cfg=2, n_docs=nd, vocab=nd // 5, n_queries=nq, k=10, name_pct=10)
corpus = synth.Corpus.for_config(cfg)
fields = synth.build_fields(corpus, 0, cfg.n_docs)
desc = nat.HostIndexDesc(cfg.n_docs, fields)
ctx = nat.Context(0)
index = nat.Index(ctx, desc)
print("tf columns:", index.info().n_columns)
batch = plan_queries(synth.gen_queries(cfg), vocab=cfg.vocab, n_text_fields=2)
oh, on, _ = orc.search(desc, batch, threads=4)
for flags, name in [(nat.FG_EXEC_COUNTERS, "gate on "), (nat.FG_EXEC_COUNTERS | nat.FG_EXEC_NO_PRUNE, "gate off")]:
    ks = batch.kmax
    d_hits, d_n = util.DevBuf((nq, ks, 2)), util.DevBuf(nq)
    pb = index.prepare(batch)
    t = time.time()
    pb.execute(d_hits.ptr, d_n.ptr, None, None, k_stride=ks, flags=flags)
    st = pb.stats()
    dt = time.time() - t
    raw = d_hits.numpy().view(np.uint32).reshape(nq, ks, 2)
    out = np.zeros((nq, ks), nat.HIT_DT)
    out["score"], out["doc"] = raw[:, :, 0].view(np.float32), raw[:, :, 1]
    n = d_n.numpy().view(np.uint32)
    assert (n == on).all()
    for qi in range(nq):
        check_topk(out[qi, :on[qi]], oh[qi, :on[qi]], 10, ctx=f"{name} query {qi}")
    print(f"{name}: emulation {dt:.1f} s, windowed chunks {st.colscan_chunks}, skipped {st.colscan_chunks_skipped}, work items {st.n_work_items}")
