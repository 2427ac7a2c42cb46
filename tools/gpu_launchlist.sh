#!/bin/bash
# launch list (gpu__time_duration.sum) of the fused cfg3 training view, eager launches (what a CUDA-graph replay contains)
mkdir -p gpurun_out
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"composite|preprocess|onesweep|finalize|emit_keys|radix_histogram|tile_scan|tile_order" -s 28 -c 28 --csv --log-file gpurun_out/launches_fused_half.csv python tools/prof_pass.py cfg3 4 fused > gpurun_out/ncu_fused.log 2>&1
wc -l gpurun_out/launches_fused_half.csv
