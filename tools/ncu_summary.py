"""Dev tool: per-kernel summary of an `ncu --set full` report (one bench step) -> text table + the
DRAM traffic of the step's search kernels as JSON (bench.py reads profiles/r01_traffic.json).
usage: ncu_summary.py report.ncu-rep out_prefix"""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
H, U = rows[0], rows[1]
cols = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_allocated", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
idx = [H.index(c) for c in cols if c in H]
lines, dram = [], 0.0
for r in rows[2:]:
    if len(r) != len(H):
        continue
    lines.append("kernel: " + r[H.index("Kernel Name")])
    for i in idx[1:]:
        lines.append(f"    {H[i]:90s} {r[i]:>16s} {U[i]}")
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
    for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = H.index(c)
        dram += float(r[i]) * scale.get(U[i], 1.0)
open(out + ".txt", "w").write("\n".join(lines) + f"\n\nDRAM bytes (read + write) over the captured launches: {dram:.0f}\n")
json.dump({"dram_bytes_per_step": int(dram), "source": rep.split("/")[-1], "launches": len([r for r in rows[2:] if len(r) == len(H)])},
          open(out + ".json", "w"))
print(open(out + ".txt").read())
