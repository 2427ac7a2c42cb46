"""Print the metrics we track from an .ncu-rep (raw page) — used to write profiles/*.md summaries."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_global_red.sum",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "sm__cycles_active.max", "sm__cycles_active.min"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=====", d.get("Kernel Name", "")[:90])
    for k in KEYS:
        if k in d:
            print(f"  {k:72s} {d[k]:>18s} {units[hdr.index(k)]}")
    items = []
    for k in hdr:
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            try:
                items.append((float(d[k].replace(",", "")), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("  stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(items, reverse=True)[:8]))
    pipes = []
    for k in hdr:
        if k.startswith("sm__inst_executed_pipe_") and k.endswith(".avg.pct_of_peak_sustained_active"):
            try:
                pipes.append((float(d[k]), k[len("sm__inst_executed_pipe_"):].split(".")[0]))
            except ValueError:
                pass
    print("  pipes (% of peak):", ", ".join(f"{n}={v:.1f}" for v, n in sorted(pipes, reverse=True)[:8]))
