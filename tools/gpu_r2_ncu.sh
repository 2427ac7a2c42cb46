#!/bin/bash
# round-2 ncu evidence: (1) --set full of every kernel of one fused cfg3 view; (2) launch list of a short bench.py run
mkdir -p gpurun_out
K='preprocess|tile_scan|emit_keys|radix_histogram|onesweep|finalize|tile_order|composite'
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 34 -c 17 -f -o gpurun_out/r2_view python tools/prof_pass.py cfg3 4 fused 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
wc -l gpurun_out/r2_launches_bench.csv; ls -la gpurun_out/*.ncu-rep
