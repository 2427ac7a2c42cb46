"""profiles/r2_view_ncu.md table + profiles/traffic.json from gpurun_out/r2_view.ncu-rep (tools/gpu_r2_ncu.sh).
    python tools/view_ncu_table.py gpurun_out/r2_view.ncu-rep <lib digest>"""
import csv
import json
import subprocess
import sys
from collections import OrderedDict

rep, digest = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
NAMES = [("onesweep", "sort_onesweep"), ("finalize_sorted", "tile_ranges"), ("tile_order", "tile_order"),
         ("composite_fwd", "composite_fwd"), ("composite_bwd", "composite_bwd"), ("preprocess_bwd", "preprocess_bwd"),
         ("preprocess_fwd", "preprocess_fwd"), ("tile_scan", "tile_scan"), ("emit_keys", "emit_keys"),
         ("radix_histogram", "sort_histogram")]


def f(d, k):
    try:
        return float(d[k].replace(",", ""))
    except (KeyError, ValueError):
        return float("nan")


agg = OrderedDict()
for r in rows[2:]:
    d = dict(zip(hdr, r))
    kn = d.get("Kernel Name", "")
    stage = next((s for key, s in NAMES if key in kn), None)
    if stage is None:
        continue
    a = agg.setdefault(stage, [])
    a.append(d)
traffic = {}
print("| kernel | launches captured | us/launch (ncu) | regs | issue-active % | warps active % | lanes/instr | warp instr (M) | DRAM read + write (MB) |")
print("|---|---|---|---|---|---|---|---|---|")
tot = 0.0
per_view = {"sort_onesweep": 5}
for stage, ds in agg.items():
    n = len(ds)
    m = lambda k: sum(f(d, k) for d in ds) / n
    us = m("gpu__time_duration.sum")
    rd, wr = m("dram__bytes_read.sum"), m("dram__bytes_write.sum")
    # the raw page reports bytes in the unit of the units row; normalise through the units row
    units = dict(zip(hdr, rows[1]))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd *= scale.get(units["dram__bytes_read.sum"], 1.0)
    wr *= scale.get(units["dram__bytes_write.sum"], 1.0)
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units["gpu__time_duration.sum"], 1.0)
    us *= tscale
    traffic[stage] = int(rd + wr)
    tot += us * per_view.get(stage, 1)
    print(f"| {stage} | {n} | {us:.1f} | {int(m('launch__registers_per_thread'))} | "
          f"{m('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | {m('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{m('smsp__thread_inst_executed_per_inst_executed.ratio'):.1f} | {m('smsp__inst_executed.sum') / 1e6:.2f} | "
          f"{rd / 1e6:.1f} + {wr / 1e6:.1f} |")
print(f"\nSum of the launches of one view under ncu (5 sort passes): {tot:.0f} us.")
json.dump({"lib_digest": digest,
           "source": "gpurun_out/r2_view.ncu-rep (ncu --set full --clock-control none, one fused cfg3 view, tools/gpu_r2_ncu.sh, "
                     "summarised by tools/view_ncu_table.py): dram__bytes_read.sum + dram__bytes_write.sum per launch",
           "cfg3": traffic}, open("profiles/traffic.json", "w"), indent=1)
