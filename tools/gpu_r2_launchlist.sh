#!/bin/bash
# ncu launch list (gpu__time_duration.sum) of the headline command, all launches; summarised by tools/launchlist_summary.py
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
wc -l gpurun_out/r2_launches_bench.csv; gzip -f -9 gpurun_out/r2_launches_bench.csv; ls -la gpurun_out/r2_launches_bench.csv.gz
