#!/bin/bash
# launch list (gpu__time_duration.sum) of the fused cfg3 training view, eager launches (what a CUDA-graph replay contains)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -k "graphed" 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hgs" -s 30 -c 45 --csv --log-file gpurun_out/launches_fused_half.csv python tools/prof_pass.py cfg3 4 fused > gpurun_out/ncu_fused.log 2>&1
wc -l gpurun_out/launches_fused_half.csv
