"""TEST INFRASTRUCTURE ONLY: runs bench.py's main() and __graft_entry__.smoke() on a box without a GPU,
against tests/emu/libfugu_emu.so, to check the SCRIPTS' logic (argument handling, accounting, the JSON
line's keys) before they meet a B200. torch's CUDA entry points are replaced by CPU stand-ins for this
process only; every number printed by such a run is meaningless."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from fugu_b200 import _native as nat  # noqa: E402
from tests import util  # noqa: E402

nat.LIB_PATH = os.path.join(ROOT, "tests", "emu", "libfugu_emu.so")
util.EMULATED = True


class _Stream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass


class _Event:
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


_real_device = torch.device
torch.device = lambda *a, **k: _real_device("cpu")
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.Stream = _Stream
torch.cuda.set_stream = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.Event = _Event

if __name__ == "__main__":
    what = sys.argv[1]
    sys.argv = [sys.argv[0]] + sys.argv[2:]
    if what == "bench":
        import bench

        bench.main()
    elif what == "smoke":
        import __graft_entry__ as g

        g.smoke()
    else:
        raise SystemExit("usage: run_entrypoints_emulated.py bench|smoke [args]")
