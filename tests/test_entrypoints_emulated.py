"""CPU: bench.py's main() and __graft_entry__.smoke() executed end to end against the emulated library
(tests/emu/run_entrypoints_emulated.py), so that a typo in the scripts is found here and not by the
round-end run on a B200. Only the scripts' logic and the shape of the JSON line are checked: the numbers
come from an emulator and mean nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "emu", "run_entrypoints_emulated.py")


def _run(*args, timeout=900):
    subprocess.check_call(["make", "-s", "-j4", "-C", os.path.join(ROOT, "tests", "emu")])
    return subprocess.run([sys.executable, DRIVER, *args], cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def test_bench_main_emits_one_complete_json_line():
    r = _run("bench", "--docs", "20000", "--vocab", "4000", "--queries", "64", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["metric"] == "queries_per_sec" and d["n_gpus"] == 1 and d["steps"] == 2 and d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    rf = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "touched", "exhaustive"):
        assert key in rf, key
    assert rf["bound"] == "hbm" and 0 < rf["algorithmic_bytes_per_launch"] <= rf["exhaustive"]["algorithmic_bytes_per_launch"]
    assert rf["touched"]["lead_blocks_tested"] >= rf["touched"]["lead_blocks_decoded"] > 0
    assert d["parity"]["checked"] > 0 and d["parity"]["failed"] == 0, d["parity"]
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] > 0


def test_bench_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--docs", "30000", "--vocab", "6000",
                        "--queries", "96", "--steps", "2", "--warmup", "1"], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_two_ranks_control_flow():
    """torchrun, 2 ranks, doc-id-range shards: the multi-rank control flow of bench.py (global statistics,
    agreed number of untimed hold steps, barriers, max-over-ranks timing, all-gather + merge per step) with
    the host-memory exchange route. The N = 8 hang of round 1 was a rank-local step count in this flow."""
    subprocess.check_call(["make", "-s", "-j4", "-C", os.path.join(ROOT, "tests", "emu")])
    env = dict(os.environ, FG_BENCH_EXCHANGE="gloo", FG_BENCH_DEADLINE="600")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29617", DRIVER, "bench", "--gpus", "2", "--docs", "20000", "--vocab", "4000", "--queries", "48",
                        "--steps", "2", "--warmup", "3", "--no-cpu-baseline"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["scaling"] == "strong" and d["value"] > 0 and "x2" in d["config"]["sharding"]
    assert d["gpu_launches"] > 0 and d["e2e"]["value"] > 0


def test_smoke_entry_point():
    r = _run("smoke")
    assert r.returncode == 0 and "smoke OK" in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]
