#!/usr/bin/env python
"""bench.py — fugu query hot path on B200: queries/s + algorithmic posting GB/s vs the HBM roofline.

Workload (BASELINE.json configs[1], SURVEY.md 8(d) "C2"): 1M-doc synthetic Zipfian corpus
(vocab 200k, 10% of docs carry metadata.name), 5000 mixed 1-4-term AND/OR queries, top-10.
A "step" = one pass of the hot path (posting decode -> AND/OR -> BM25 -> top-k, + per-query merge)
over the whole 5000-query batch.

  value    : queries/s with the index AND the lowered query batch resident in HBM, CUDA-event time
  e2e      : queries/s through the reference-facing blocking call fgh_search_batch with HOST buffers
             (query strings -> planner -> plan lowering -> H2D of the plan -> kernels -> D2H of hits, all
             inside the timed region)
  roofline : algorithmic posting bytes (stated on the posting-block layout) of one step's
             concurrent group of search kernels / its CUDA-event duration vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference : the CPU oracle (restatement of tantivy 0.24.1 semantics — the
             real reference cannot be built here: no Rust toolchain) on the box's host cores.

N > 1 (torchrun): documents are sharded by doc-id range (strong scaling: fixed corpus), every rank
evaluates the whole batch on its shard, local top-k -> all_gather (NCCL) -> on-device merge.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--docs", type=int, default=None, help="default: the configuration's size (SURVEY.md 8(d))")
    ap.add_argument("--vocab", type=int, default=None)
    ap.add_argument("--queries", type=int, default=None)
    ap.add_argument("--cfg", type=int, default=2, help="BASELINE.json configuration 1..5 (2 = the metric's 1-GPU configuration)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    a = ap.parse_args()
    from fugu_b200 import synth

    base = synth.CONFIGS[a.cfg]
    a.docs = a.docs or base.n_docs
    a.vocab = a.vocab or base.vocab
    a.queries = a.queries or base.n_queries
    return a


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines: list[str] = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(args, rank: int, world: int):
    """Corpus shard of this rank (doc-id range), the query strings and the lowered plan."""
    from fugu_b200 import _native as nat
    from fugu_b200 import synth

    base = synth.CONFIGS[args.cfg]
    cfg = synth.Config(cfg=base.cfg, n_docs=args.docs, vocab=args.vocab, n_queries=args.queries, k=base.k,
                       name_pct=base.name_pct, n_ns=base.n_ns)
    corpus = synth.Corpus.for_config(cfg)
    d0, d1 = cfg.n_docs * rank // world, cfg.n_docs * (rank + 1) // world
    fields = synth.build_fields(corpus, d0, d1)
    queries = synth.gen_queries(cfg)
    flt = os.environ.get("FG_BENCH_FILTER")  # dev: time one query shape of the mix ("and", "or", "single"); never set by the driver
    if flt:
        def shape(q):
            s = q["query"]
            return "and" if " AND " in s else ("single" if " " not in s.strip() else "or")
        queries = [q for q in queries if shape(q) == flt]
        cfg = synth.Config(cfg=cfg.cfg, n_docs=cfg.n_docs, vocab=cfg.vocab, n_queries=len(queries), k=cfg.k, name_pct=cfg.name_pct, n_ns=cfg.n_ns)
    return cfg, corpus, fields, queries, d0, d1


def term_lists(corpus, cfg, n_fields):
    """Term dictionaries of the synthetic corpus (term ordinal r-1 <-> "w<r>"; facet paths)."""
    words = [f"w{i + 1}" for i in range(cfg.vocab)]
    terms = [words]
    if n_fields >= 2:
        terms.append(words if cfg.name_pct > 0 else [])
    if n_fields >= 3:
        terms.append([corpus.facet_path(i) for i in range(corpus.facet_vocab())])
    return terms


def ncu_traffic(cfg_id: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the search kernel of one step, from the committed
    `ncu --set full` capture of THIS round's build (profiles/r02_traffic.json, written by tools/ncu_summary.py
    from the capture named in it), or None when no capture of this configuration is committed."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        return json.load(open(p))[f"C{cfg_id}"]["dram_bytes_per_step"]
    except Exception:
        return None


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def compare_topk(gs, gd, os_, od, k, tol=1e-5):
    """None when the hit lists agree (scores rank-wise within tol, docs equal except inside runs of scores tied
    within tolerance, a run cut at rank k may differ), else a description of the first difference."""
    if len(gs) != len(os_):
        return f"n_hits {len(gs)} != {len(os_)}"
    n = len(os_)
    for i in range(n):
        a, b = float(gs[i]), float(os_[i])
        if abs(a - b) > tol * max(abs(a), abs(b), 1e-30):
            return f"rank {i}: score {a!r} vs {b!r}"
    i = 0
    while i < n:
        j = i
        while j + 1 < n and abs(float(os_[j + 1]) - float(os_[j])) <= 4 * tol * max(abs(float(os_[j])), 1e-30):
            j += 1
        if set(gd[i:j + 1].tolist()) != set(od[i:j + 1].tolist()) and not (j == n - 1 and n == k):
            return f"ranks {i}..{j}: docs {sorted(gd[i:j + 1].tolist())} vs {sorted(od[i:j + 1].tolist())}"
        i = j + 1
    return None


def check_apart_sharded(ds, comm, queries, batch, pdesc, rank):
    """N > 1: the requests the fused exchange does not take -- deep pages (limit above 1024) and nested boolean queries --
    through the collective host call (every shard answers them alone, pages merged on the host). COLLECTIVE: every rank
    calls; rank 0 checks: complete deep lists against the oracle's ranking of the unsharded corpus, the deep page as
    rows of that list, a nested query (A) OR (B) against the oracle's union of A's and B's scorers."""
    from fugu_b200 import _native as nat
    from fugu_b200.dataset import QuerySet

    nd, limit, pp = 8, 3100, 100
    base = [q["query"] for q in queries[:80]]
    nested = [f"({base[2 * i]}) OR ({base[2 * i + 1]})" for i in range(4)]
    whole = QuerySet(base[:nd], None, 0, limit)                 # complete first 3100 of 8 queries
    w_hits, w_n, w_st = ds.search_batch_sharded(comm, whole)
    mixed = QuerySet(base + nested, None, 0, pp)                # ordinary pages + deep pages + nested, one request
    mixed.pages[:nd] = limit // pp - 1
    m_hits, m_n, m_st = ds.search_batch_sharded(comm, mixed)
    if rank != 0:
        return None
    from oracle import orc  # checker only

    failed, first = 0, None

    def bad(msg):
        nonlocal failed, first
        failed += 1
        first = first or msg

    if not ((w_st == 0).all() and (m_st == 0).all()):
        bad(f"status {w_st.tolist()} {m_st[m_st != 0].tolist()}")
    q = batch.q[:nd].copy()
    q["k"] = limit
    o_hits, o_n, _ = orc.search(pdesc, nat.HostBatch.from_arrays(q, batch.c, batch.l), threads=host_cores())
    for i in range(nd):
        msg = compare_topk(w_hits[i, :w_n[i]]["score"], w_hits[i, :w_n[i]]["doc"], o_hits["score"][i, :o_n[i]], o_hits["doc"][i, :o_n[i]], limit)
        if msg:
            bad(f"deep list {i} ({base[i]!r}): {msg}")
        want = w_hits[i, limit - pp:w_n[i]]
        if m_n[i] != len(want) or not np.array_equal(m_hits[i, :m_n[i]], want):
            bad(f"deep page {i} ({base[i]!r}) is not rows [{limit - pp}, {limit}) of the complete list")
    for j, s_ in enumerate(nested):
        children = nat.HostBatch.from_arrays(batch.q[[2 * j, 2 * j + 1]].copy(), batch.c, batch.l)
        u_hits, _ = orc.search_union_of(pdesc, children, pp)   # the oracle's union of the two children's scorers
        qi = len(base) + j
        msg = compare_topk(m_hits[qi, :m_n[qi]]["score"], m_hits[qi, :m_n[qi]]["doc"], u_hits["score"], u_hits["doc"], pp)
        if msg:
            bad(f"nested {s_!r}: {msg}")
    return {"checked": 2 * nd + len(nested), "failed": failed, "first_failure": first,
            "what": "fgh_search_batch_sharded: deep lists vs oracle, deep pages as rows of them, nested (A) OR (B) vs the oracle's union of the children"}


def run_reference(args):
    """--impl reference: the reference's CPU path on the box's host cores. The real reference
    (Rust + tantivy 0.24.1) cannot be built here, so this is the oracle port (oracle/oracle.cpp)."""
    rank, _, world = env_rank()
    if rank != 0:
        return
    from fugu_b200 import _native as nat
    from fugu_b200 import synth
    from oracle import orc

    cfg, corpus, fields, queries, d0, d1 = build_workload(args, 0, 1)
    from fugu_b200.dataset import Dataset, QuerySet

    desc = nat.HostIndexDesc(cfg.n_docs, fields)
    n_fields = len(fields)
    cores = host_cores()
    pds = Dataset(None)  # planning only (no device): the same C++ planner both arms use
    pds.adopt(desc, term_lists(corpus, cfg, n_fields))
    # bounded sample: each step = the whole batch when it is small enough, else a prefix
    nq = min(len(queries), 5000)
    qset = QuerySet([q["query"] for q in queries[:nq]], [q["filters"] for q in queries[:nq]], 0, cfg.k)
    # block-max metadata of the index (tantivy keeps it in its skip entries: index-build time, not query time)
    bmx = orc.BlockMax(desc, cores)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        batch, _ = pds.plan_batch(qset)  # query strings -> plan, inside the timed region as on the GPU e2e arm
        # TopDocs form: block-max pruning where tantivy prunes (unions of plain term scorers), exhaustive scorers elsewhere
        orc.search_pruned(bmx, batch, threads=cores)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    tot = sum(times)
    qps = nq * len(times) / tot
    line = {
        "impl": "reference", "metric": "queries_per_sec", "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, args, 1),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": f"{nq} of {len(queries)} queries per step, oracle C++ -O3 -march=native (restatement of tantivy 0.24.1: "
                                   f"buffered-union / intersection DAAT; block-max pruning for unions of plain term scorers, where tantivy "
                                   f"runs block-max WAND), {cores} threads one query per thread, planning included; cpu: {cpu_model()}"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


EXCHANGE_NOTE: dict = {}  # filled by main(): how the per-shard top-k lists were exchanged
QUERY_SHAPES = {1: "two-term AND queries", 2: "mixed 1-4-term AND/OR queries", 3: "OR queries of 2-6 of the 64 most frequent terms",
                4: "3-term AND queries", 5: "2-term OR queries with 1-2 facet filters"}


def workload_config(cfg, args, world):
    return {"workload": f"C{cfg.cfg}: {cfg.n_docs}-doc synthetic Zipfian corpus (vocab {cfg.vocab}, {cfg.name_pct}% docs with name), "
                        f"{cfg.n_queries} {QUERY_SHAPES.get(cfg.cfg, 'queries')}, top-{cfg.k}",
            "n_docs": cfg.n_docs, "n_queries": cfg.n_queries, "k": cfg.k,
            "sharding": f"doc-id range x{world}" if world > 1 else "single shard",
            "exchange": EXCHANGE_NOTE.get("mode", "none"),
            "l2": "L2 flushed (256 MiB memset) before every timed step" if not args.no_flush else "warm L2 (no flush)"}


class Exchange:
    """The multi-GPU exchange step of the path: all-gather of the per-shard top-k lists,
    plus the small control-plane reductions of the bench (global statistics, barriers, max-over-ranks
    timing). Default route: everything on NCCL with device tensors (`all_gather_into_tensor` over
    NVLink) - the configuration verified on B200s. `FG_BENCH_EXCHANGE=gloo` is an explicit opt-in for a
    box whose NCCL cannot initialise: the lists then travel D2H -> gloo -> H2D and the JSON line says so
    (`config.exchange`)."""

    def __init__(self, dist, torch, dev, mode: str = "nccl"):
        self.dist, self.torch, self.dev, self.mode = dist, torch, dev, mode
        self.world = dist.get_world_size()

    def barrier(self):
        self.dist.barrier()

    def allreduce_cpu(self, arr: np.ndarray, op: str = "sum") -> np.ndarray:
        t = self.torch.from_numpy(np.ascontiguousarray(arr).copy())
        if self.mode == "nccl":
            t = t.to(self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def all_gather(self, out, inp):
        """out [world, ...] <- inp [...] from every rank (device tensors)."""
        if self.mode == "nccl":
            self.dist.all_gather_into_tensor(out, inp)
            return
        h = inp.cpu()
        parts = [self.torch.empty_like(h) for _ in range(self.world)]
        self.dist.all_gather(parts, h)
        out.copy_(self.torch.stack(parts).to(out.device, non_blocking=False))


def arm_watchdog(seconds: float):
    """A hung collective or kernel must not eat the caller's whole time limit: after `seconds` dump every
    thread's Python stack to stderr and exit non-zero (FG_BENCH_DEADLINE overrides; 0 disables)."""
    import faulthandler

    seconds = float(os.environ.get("FG_BENCH_DEADLINE", seconds))
    if seconds <= 0:
        return
    faulthandler.enable()
    faulthandler.dump_traceback_later(max(30.0, seconds - 20.0), repeat=False, file=sys.stderr)

    def _kill():
        time.sleep(seconds)
        sys.stderr.write(f"bench.py: no result after {seconds:.0f} s (rank {os.environ.get('RANK', '0')}), giving up\n")
        sys.stderr.flush()
        os._exit(124)

    threading.Thread(target=_kill, daemon=True).start()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    rank, local_rank, world = env_rank()
    arm_watchdog(420 if world == 1 else 300)
    # stdout carries exactly ONE JSON line: anything native libraries print there (NCCL's version banner)
    # is sent to stderr; the line itself is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        # leave the host cores to all ranks (planner / lowering pools and the upload's packing threads)
        os.environ.setdefault("FG_HOST_THREADS", str(max(2, host_cores() // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world))))))
    import torch

    from fugu_b200 import _native as nat
    from fugu_b200 import synth

    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the product has no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout (one JSON line)
        import torch.distributed as dist_mod

        dist = dist_mod
        xmode = os.environ.get("FG_BENCH_EXCHANGE", "nccl")
        if xmode == "nccl":
            dist.init_process_group("nccl", device_id=dev)
        else:
            os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
            dist.init_process_group("gloo")
    xch = Exchange(dist, torch, dev, mode=xmode) if dist else None
    EXCHANGE_NOTE["mode"] = ("inside the library: fg_batch_execute_sharded = local top-k -> ncclAllGather (one fused launch) -> on-device merge"
                             if xch.mode == "nccl" else
                             "host-memory (gloo) all-gather + on-device merge (FG_BENCH_EXCHANGE=gloo)") if xch else "none"

    ctx = nat.Context(local_rank)
    comm = None
    if world > 1 and xch.mode == "nccl":
        # the library's own communicator (fg_comm): rank 0's id travels over the launcher's process group once
        ids = [nat.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = nat.Comm(ctx, rank, world, ids[0])
    cfg, corpus, fields, queries, d0, d1 = build_workload(args, rank, world)
    n_local = d1 - d0
    if world > 1:
        # global statistics (tantivy computes N, df and total_num_tokens over all segments, A.4)
        for f in fields:
            if comm:
                f["global_doc_freq"] = comm.allreduce_sum(np.diff(f["term_offsets"]).astype(np.uint32))
                f["total_num_tokens"] = int(comm.allreduce_sum(np.array([f["total_num_tokens"]], np.uint64))[0])
            else:
                f["global_doc_freq"] = xch.allreduce_cpu(np.diff(f["term_offsets"]).astype(np.int64)).astype(np.uint32)
                f["total_num_tokens"] = int(xch.allreduce_cpu(np.array([f["total_num_tokens"]], np.int64))[0])
    desc = nat.HostIndexDesc(n_local, fields, doc_id_base=d0, global_n_docs=cfg.n_docs)
    # a real (non-default) stream: the legacy default stream's handle is 0, which fg_ctx_set_stream
    # reads as "use the context's own stream"; events below are recorded on this same stream
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    from fugu_b200.dataset import Dataset, QuerySet

    n_fields = len(fields)
    t0 = time.perf_counter()
    ds = Dataset(ctx)
    ds.adopt(desc, term_lists(corpus, cfg, n_fields))  # fg_index_upload + term dictionaries for the planner
    index = ds.index()
    upload_s = time.perf_counter() - t0
    info = index.info()
    qset = QuerySet([q["query"] for q in queries], [q["filters"] for q in queries], 0, cfg.k)
    batch, pstatus = ds.plan_batch(qset)
    assert (pstatus == 0).all(), "planner rejected a benchmark query"
    nq, k = batch.n_queries, batch.kmax
    pb = index.prepare(batch)

    d_hits = torch.zeros((nq, k, 2), dtype=torch.int32, device=dev)
    d_n = torch.zeros(nq, dtype=torch.int32, device=dev)
    d_c = torch.zeros(nq, dtype=torch.int32, device=dev)
    if world > 1:
        g_hits = torch.zeros((world, nq, k, 2), dtype=torch.int32, device=dev)
        g_n = torch.zeros((world, nq), dtype=torch.int32, device=dev)
        f_hits = torch.zeros((nq, k, 2), dtype=torch.int32, device=dev)
        f_n = torch.zeros(nq, dtype=torch.int32, device=dev)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    counts = bool(os.environ.get("FG_BENCH_COUNTS"))  # TopDocs::with_limit does not count matches: off by default

    def step():
        if comm:  # one call: local top-k -> NCCL all-gather -> merge, all inside the library
            pb.execute_sharded(comm, f_hits.data_ptr(), f_n.data_ptr(), k_stride=k)
            return
        pb.execute(d_hits.data_ptr(), d_n.data_ptr(), d_c.data_ptr() if counts else None, None, k_stride=k)
        if world > 1:
            xch.all_gather(g_hits, d_hits)
            xch.all_gather(g_n, d_n)
            nat.merge_topk_device(ctx, g_hits.data_ptr(), g_n.data_ptr(), world, nq, k, k, f_hits.data_ptr(), f_n.data_ptr())

    # exact algorithmic-byte accounting pass (untimed). SURVEY.md 8(d) states the algorithmic bytes on
    # exhaustive evaluation of the posting-block layout: this pass lowers the plan for the windowed kernels
    # WITHOUT dense tf columns (FG_PREP_LEGACY | FG_PREP_NO_COLUMNS: every leaf decoded from its blocks).
    pb_blocks = index.prepare(batch, nat.FG_PREP_NO_COLUMNS | nat.FG_PREP_LEGACY)
    pb_blocks.execute(d_hits.data_ptr(), d_n.data_ptr(), d_c.data_ptr(), None, k_stride=k, flags=nat.FG_EXEC_EXACT_ACCOUNTING)
    st = pb_blocks.stats()
    algo_bytes = st.bytes_blocks + st.scored_postings + 8 * st.sum_k
    pb_blocks.close()
    # bytes the timed configuration really touches ("if the implementation prunes, report touched
    # bytes (kernel counters) and use min(algorithmic, touched)"): the same batch, counters on, untimed. Counted
    # by the kernel: payload + skip entry of every block decoded (lead or lookup), block-max words and skip
    # entries read while skipping / galloping, 1-byte gathers (fieldnorm ids, tf-column bytes), 8 B per hit.
    pb.execute(d_hits.data_ptr(), d_n.data_ptr(), d_c.data_ptr() if counts else None, None, k_stride=k, flags=nat.FG_EXEC_COUNTERS)
    st_touched = pb.stats()
    touched_bytes = st_touched.bytes_blocks + st_touched.bytes_meta + st_touched.scored_postings + 8 * st_touched.sum_k
    touched_algo_bytes = float(min(algo_bytes, touched_bytes))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    t_w = time.perf_counter()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    # keep the GPU under the same load for ~1 s so that nvidia-smi (200 ms period) sees the clocks the
    # timed steps run at; these extra steps are untimed warm-up. Every step contains collectives when
    # N > 1, so ALL RANKS MUST RUN THE SAME NUMBER OF THEM: the count is agreed on (max over ranks), never
    # decided by each rank's own clock (a rank-local `while time < deadline` loop deadlocks as soon as two
    # ranks disagree by one step).
    t_step = max((time.perf_counter() - t_w) / max(args.warmup, 1), 1e-4)
    n_hold = int(min(2000, max(1, round(1.0 / t_step)))) if args.warmup else 100
    if os.environ.get("FG_BENCH_HOLD"):  # dev (profiling runs under ncu): fewer untimed hold steps
        n_hold = int(os.environ["FG_BENCH_HOLD"])
    if dist:
        n_hold = int(xch.allreduce_cpu(np.array([n_hold], np.int64), "max")[0])
    for _ in range(n_hold):
        step()
    torch.cuda.synchronize()
    if dist:
        xch.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms = []
    torch.cuda.synchronize()
    for i in range(args.steps):
        if not args.no_flush:
            flush_buf.fill_(i & 0xFF)
        ev[i][0].record(stream)
        step()
        ev[i][1].record(stream)
        s2 = pb.stats()  # synchronises; per-launch search-kernel time from the library's own CUDA events
        kern_ms.append(s2.search_kernel_ms)
    torch.cuda.synchronize()
    if dist:
        xch.barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    if dist:
        total_ms = float(xch.allreduce_cpu(np.array([total_ms], np.float64), "max")[0])
        algo_total = float(xch.allreduce_cpu(np.array([float(touched_algo_bytes)], np.float64))[0])
    else:
        algo_total = float(touched_algo_bytes)
    st_timed = pb.stats()
    ms_per_step = total_ms / args.steps
    qps = nq / (ms_per_step * 1e-3)
    step_ms = [a.elapsed_time(b) for a, b in ev]

    # ---- the same batch evaluated exhaustively (every posting of every leaf visited; results identical): the pass on
    # which algorithmic bytes == touched bytes. It runs on the engine the library uses whenever every match must be
    # visited (match counts): the windowed accumulator kernels (FG_PREP_LEGACY, dense tf columns on). Timed the same
    # way, reported beside the pruned figures. ----
    exh_ms = []
    pb_exh = index.prepare(batch, nat.FG_PREP_LEGACY)
    for i in range(2 + min(args.steps, 5)):
        if not args.no_flush:
            flush_buf.fill_(i & 0xFF)
        pb_exh.execute(d_hits.data_ptr(), d_n.data_ptr(), None, None, k_stride=k, flags=nat.FG_EXEC_NO_PRUNE)
        s2 = pb_exh.stats()
        if i >= 2:
            exh_ms.append(s2.search_kernel_ms)
    pb_exh.close()
    step()  # leave the buffers holding the result of the timed (pruned) configuration
    torch.cuda.synchronize()

    # ---- parity of the benchmarked configuration at the benchmarked size: the hits the timed path produced
    # vs the oracle on the same batch (first PARITY_Q queries; scores within 1e-5 relative, same docs except
    # inside ties) ----
    parity = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import orc  # checker only

        pq = min(nq, int(os.environ.get("FG_BENCH_PARITY_Q", "1000")))
        if world > 1:
            full_fields = synth.build_fields(corpus, 0, cfg.n_docs)
            pdesc = nat.HostIndexDesc(cfg.n_docs, full_fields)
            got_hits, got_n = f_hits.cpu().numpy(), f_n.cpu().numpy().view(np.uint32)
        else:
            pdesc = desc
            got_hits, got_n = d_hits.cpu().numpy(), d_n.cpu().numpy().view(np.uint32)
        sub = nat.HostBatch.from_arrays(batch.q[:pq], batch.c, batch.l)
        o_hits, o_n, _ = orc.search(pdesc, sub, threads=host_cores())
        raw = got_hits.view(np.uint32).reshape(nq, k, 2)
        failed, first = 0, None
        for qi in range(pq):
            msg = compare_topk(raw[qi, :, 0].view(np.float32)[:got_n[qi]], raw[qi, :, 1][:got_n[qi]],
                               o_hits["score"][qi, :o_n[qi]], o_hits["doc"][qi, :o_n[qi]], int(batch.q["k"][qi]))
            if msg:
                failed += 1
                first = first or f"query {qi} ({queries[qi]['query']!r}): {msg}"
        parity = {"checked": pq, "failed": failed, "against": "oracle/oracle.cpp on the same batch and corpus",
                  "tolerance": "scores 1e-5 relative, doc ids equal except inside ties", "first_failure": first}

    # ---- e2e through the blocking host-buffer ABI call ----
    e2e_times = []
    for it in range(2 + min(args.steps, 10)):
        if dist:
            xch.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if comm:
            # fgh_search_batch_sharded (collective): strings (host) -> planning shared by the ranks -> lowering for this
            # shard -> H2D -> kernels -> NCCL all-gather -> merge -> D2H (host), pipelined in chunks like the 1-GPU call
            e_hits, e_n, e_status = ds.search_batch_sharded(comm, qset)
        else:
            h_hits, h_n, h_c, _ = ds.search_batch(qset, want_counts=counts)  # fgh_search_batch: strings -> plan -> H2D -> kernels -> D2H
            if world > 1:
                xch.all_gather(g_hits, torch.from_numpy(h_hits.view(np.int32).reshape(nq, k, 2)).to(dev))
                xch.all_gather(g_n, torch.from_numpy(h_n.view(np.int32)).to(dev))
                nat.merge_topk_device(ctx, g_hits.data_ptr(), g_n.data_ptr(), world, nq, k, k, f_hits.data_ptr(), f_n.data_ptr())
                f_hits.cpu(); f_n.cpu()
        dt = time.perf_counter() - t0
        if it >= 2:
            e2e_times.append(dt)
    e2e_s = float(np.mean(e2e_times))
    if comm and rank == 0:  # the end-to-end call must return what the timed device path produced
        ref_h, ref_n = f_hits.cpu().numpy().view(np.uint32).reshape(nq, k, 2), f_n.cpu().numpy().view(np.uint32)
        assert (e_status == 0).all() and np.array_equal(e_n, ref_n), "fgh_search_batch_sharded: hit counts differ from the device path"
        for qi in range(nq):
            assert np.array_equal(e_hits[qi, :e_n[qi]]["doc"], ref_h[qi, :ref_n[qi], 1]), f"fgh_search_batch_sharded: query {qi} differs"
    apart_check = None
    if comm and len(queries) >= 80 and not any(q.get("filters") for q in queries[:80]) and not os.environ.get("FG_BENCH_NO_APART"):
        if rank == 0 and args.no_cpu_baseline:  # (otherwise the parity check above has built the unsharded corpus)
            pdesc = nat.HostIndexDesc(cfg.n_docs, synth.build_fields(corpus, 0, cfg.n_docs))
        apart_check = check_apart_sharded(ds, comm, queries, batch, pdesc if rank == 0 else None, rank)
    if os.environ.get("FG_TIMING"):
        print("e2e_times ms", [round(x * 1e3, 2) for x in e2e_times], file=sys.stderr)
    if dist:
        e2e_s = float(xch.allreduce_cpu(np.array([e2e_s], np.float64), "max")[0])
    lowered_bytes = st_touched.plan_bytes  # the lowered plan as uploaded: LQuery + LLeaf arrays + work-item records
    out_bytes = nq * k * 8 + nq * 8

    def shutdown():
        if comm:
            comm.close()
        if dist:
            dist.destroy_process_group()

    if rank != 0:
        shutdown()
        return

    peak, peak_src = peak_hbm()
    kms = float(np.mean(kern_ms)) if kern_ms else ms_per_step
    achieved = touched_algo_bytes / (kms * 1e-3) / 1e9  # this rank's launch; min(algorithmic, touched)
    exh_kms = float(np.mean(exh_ms)) if exh_ms else kms
    line = {
        "metric": "queries_per_sec", "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, args, world),
        "posting_gbs": algo_total / (ms_per_step * 1e-3) / 1e9,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(cfg.cfg), "peak_source": peak_src,
                     "kernel": "lead_kernel (fg_lead.cu): one persistent launch per step walks the lead leaves block by block "
                               "(bit-unpack + warp prefix sum), looks candidates up in the other leaves (tf-column byte or "
                               "skip-table gallop + block decode), BM25, per-warp top-k; + lead_init_kernel, lead_merge_kernel",
                     "kernel_ms": kms, "kernel_ms_min": float(np.min(kern_ms)), "kernel_ms_median": float(np.median(kern_ms)),
                     "algorithmic_bytes_per_launch": int(touched_algo_bytes),
                     "bytes_per_query": touched_algo_bytes / nq,
                     "numerator": "min(algorithmic, touched) per SURVEY.md 8(d): the timed configuration prunes (MaxScore + "
                                  "block-max), so the numerator is the bytes the kernel really read (device counters of an untimed "
                                  "pass of the same batch)",
                     "touched": {"block_bytes": int(st_touched.bytes_blocks), "meta_bytes": int(st_touched.bytes_meta),
                                 "gather_bytes": int(st_touched.scored_postings), "hit_bytes": int(8 * st_touched.sum_k),
                                 "lead_blocks_decoded": int(st_touched.lead_blocks), "lead_blocks_tested": int(st_touched.lead_blocks_seen)},
                     "exhaustive_equivalent_gbs": algo_bytes / (kms * 1e-3) / 1e9,
                     "exhaustive": {"algorithmic_bytes_per_launch": int(algo_bytes), "kernel_ms": exh_kms,
                                    "achieved": algo_bytes / (exh_kms * 1e-3) / 1e9, "frac": algo_bytes / (exh_kms * 1e-3) / 1e9 / peak,
                                    "note": "the same batch with every posting of every leaf visited (results identical), on the "
                                            "engine the library uses for match counts (windowed accumulator kernels, FG_PREP_LEGACY): "
                                            "algorithmic bytes of SURVEY.md 8(d) (exact-accounting pass on the block layout) over this "
                                            "pass's own CUDA-event kernel time; exhaustive_equivalent_gbs = the same bytes over the "
                                            "TIMED (pruned) kernel time, i.e. what an exhaustive evaluator would need to sustain to "
                                            "answer as fast"},
                     "note": "index snapshot is %.0f MB in HBM incl. %.0f MB of tf columns (L2 is 126 MB); L2 is flushed before "
                             "each timed step" % (info.device_bytes / 1e6, info.column_bytes / 1e6)},
        "e2e": {"value": nq / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(lowered_bytes),
                "d2h_bytes_per_step": int(out_bytes), "ms_per_step": e2e_s * 1e3,
                "what": ("fgh_search_batch_sharded (collective): query strings (host) -> planning shared by the ranks (all-gather of the "
                         "plans) -> lowering for this shard -> H2D -> kernels -> NCCL all-gather + merge -> D2H (host), two pipelined chunks"
                         if comm else
                         "fgh_search_batch: query strings (host) -> C++ planner -> plan lowering -> H2D plan -> kernels -> D2H hits "
                         "(host); requests of 3072 queries and more are pipelined in two chunks (35 / 65 %%); match counts %s" % ("on" if counts else "off (TopDocs does not count)"))},
        "gpu_launches": int(st_timed.n_launches + (1 if world > 1 else 0)) * args.steps,
        "parity": parity,
        "parity_deep_and_nested_sharded": apart_check,
        "ms_per_step_min": float(np.min(step_ms)), "ms_per_step_median": float(np.median(step_ms)),
        "clocks": clocks,
        "index": {"postings": int(info.n_postings), "blocks": int(info.n_blocks), "packed_bytes": int(info.packed_bytes),
                  "device_bytes": int(info.device_bytes), "upload_s": upload_s, "work_items": int(st_touched.n_work_items),
                  "tf_columns": int(info.n_columns), "tf_column_bytes": int(info.column_bytes)},
    }
    if not args.no_cpu_baseline:
        from oracle import orc  # cpu_baseline leg: the one place bench.py may run the oracle (as the measured CPU arm)

        odesc = pdesc  # the whole corpus at every N (at N > 1 rank 0 rebuilt it for the parity check): comparable across N
        cores = host_cores()
        ns = min(nq, 5000)
        sub, _ = ds.plan_batch(QuerySet([q["query"] for q in queries[:ns]], [q["filters"] for q in queries[:ns]], 0, cfg.k))
        best = None
        t_end = time.perf_counter() + 20
        bmx = orc.BlockMax(odesc, cores)  # index-time metadata (tantivy: block maxima in the skip entries)
        for _ in range(3):
            t0 = time.perf_counter()
            orc.search_pruned(bmx, sub, threads=cores)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            if time.perf_counter() > t_end:
                break
        t0 = time.perf_counter()
        sub1, _ = ds.plan_batch(QuerySet([q["query"] for q in queries[:max(1, ns // 10)]], None, 0, cfg.k))
        orc.search_pruned(bmx, sub1, threads=1)
        dt1 = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": ns / best, "unit": "queries/s", "cores": cores, "kind": "port",
                                "single_thread_qps": max(1, ns // 10) / dt1,
                                "sample": f"first {ns} queries of the batch on the whole corpus, best of 3, oracle C++ -O3 -march=native "
                                          f"(restatement of tantivy 0.24.1 semantics: DAAT scorers; block-max pruning for unions of plain term "
                                          f"scorers, where tantivy runs block-max WAND), {cores} threads; cpu: {cpu_model()}"}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    shutdown()
    if apart_check and apart_check["failed"]:
        sys.stderr.write(f"bench.py: PARITY FAILURE of deep pages / nested queries across shards: {apart_check}\n")
        sys.exit(3)
    if parity and parity["failed"]:
        sys.stderr.write(f"bench.py: PARITY FAILURE on the benchmarked batch: {parity}\n")
        sys.exit(3)


if __name__ == "__main__":
    main()
