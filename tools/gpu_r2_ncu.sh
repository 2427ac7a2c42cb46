#!/bin/bash
# round-2 ncu evidence: --set full of every kernel of one fused cfg3 view (eager launches of what a graph replay contains)
mkdir -p gpurun_out
K='preprocess|tile_scan|emit_keys|radix_histogram|onesweep|finalize|tile_order|composite'
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 34 -c 17 -f -o gpurun_out/r2_view python tools/prof_pass.py cfg3 4 fused 2>&1 | tail -2
ls -la gpurun_out/*.ncu-rep; cat hair-gs_b200/lib/libhairgs_rast.stamp
