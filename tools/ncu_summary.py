"""Dev tool: per-kernel summary of an `ncu --set full` report (one bench step) -> text table, and the DRAM traffic of
the step's kernels as an entry of profiles/r02_traffic.json (bench.py reads it for roofline.traffic).
usage: ncu_summary.py report.ncu-rep out_prefix [config-key e.g. C2]"""
import csv, json, os, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
key = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
H, U = rows[0], rows[1]
cols = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_allocated", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
idx = [H.index(c) for c in cols if c in H]
lines, dram = [], 0.0
scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
for r in rows[2:]:
    if len(r) != len(H):
        continue
    lines.append("kernel: " + r[H.index("Kernel Name")])
    for i in idx[1:]:
        lines.append(f"    {H[i]:90s} {r[i]:>16s} {U[i]}")
    if "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum" in H and "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum" in H:
        s_, q_ = float(r[H.index("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")]), float(r[H.index("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")])
        if q_:
            lines.append(f"    {'sectors per global-load request (32 B sectors; 4 = a fully coalesced 32-bit warp load)':90s} {s_ / q_:16.2f}")
    for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = H.index(c)
        v = float(r[i])
        if v == v:  # (a kernel whose metric passes were cut short reports nan)
            dram += v * scale.get(U[i], 1.0)
n_launch = len([r for r in rows[2:] if len(r) == len(H)])
open(out + ".txt", "w").write("\n".join(lines) + f"\n\nDRAM bytes (read + write) over the {n_launch} captured launches: {dram:.0f}\n")
if key:
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_traffic.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[key] = {"dram_bytes_per_step": int(dram), "source": os.path.basename(rep), "launches": n_launch}
    json.dump(d, open(path, "w"), indent=1)
print(open(out + ".txt").read())
