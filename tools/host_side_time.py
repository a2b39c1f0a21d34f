"""Dev tool (no GPU needed): host time of the request path for the bench's 5000-query C2 batch — the C++
planner (fgh_plan_batch) and the plan lowering (fg_batch_prepare) — against the emulated library, whose
host code is the product's (only the device calls are stand-ins). FG_HOST_THREADS=n sets the pool size.

    python tools/host_side_time.py [n_docs=1000000]
"""
import ctypes as C
import os
import subprocess
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from fugu_b200 import _native as nat  # noqa: E402

subprocess.check_call(["make", "-s", "-j4", "-C", os.path.join(ROOT, "tests", "emu")])
nat.LIB_PATH = os.path.join(ROOT, "tests", "emu", "libfugu_emu.so")
os.environ["FG_TIMING"] = "1"
import bench  # noqa: E402
from fugu_b200 import dataset as dsm  # noqa: E402
from fugu_b200.dataset import Dataset, QuerySet  # noqa: E402

a = types.SimpleNamespace(cfg=2, docs=int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, vocab=200_000, queries=5000)
cfg, corpus, fields, queries, d0, d1 = bench.build_workload(a, 0, 1)
desc = nat.HostIndexDesc(cfg.n_docs, fields)
ctx = nat.Context(0)
ds = Dataset(ctx)
t = time.perf_counter()
ds.adopt(desc, bench.term_lists(corpus, cfg, len(fields)))
print(f"fg_index_upload (host packing, emulated copies): {time.perf_counter() - t:.2f} s")
index = ds.index()
qs = QuerySet([q["query"] for q in queries], [q["filters"] for q in queries], 0, cfg.k)
n = qs.n
q = np.zeros(n, nat.QUERY_DT)
c = np.zeros(n * 16, nat.CLAUSE_DT)
lv = np.zeros(n * 64, nat.LEAF_DT)
nc, nl = C.c_uint32(), C.c_uint32()
status = np.zeros(n, np.int32)
L = dsm._L()
for rep in range(5):
    t = time.perf_counter()
    nat.check(L.fgh_plan_batch(ds.h, n, qs.qarr, qs.farr, None, qs.pages.ctypes.data, qs.pps.ctypes.data, q.ctypes.data, c.ctypes.data,
                               len(c), lv.ctypes.data, len(lv), C.byref(nc), C.byref(nl), status.ctypes.data))
    t1 = time.perf_counter()
    batch = nat.HostBatch.from_arrays(q, c[:nc.value], lv[:nl.value])
    t2 = time.perf_counter()
    pb = index.prepare(batch)
    t3 = time.perf_counter()
    print(f"plan {1e3 * (t1 - t):.2f} ms   lower + upload {1e3 * (t3 - t2):.2f} ms")
    pb.close()
