#!/bin/bash
# ncu launch list (gpu__time_duration.sum) of the end-to-end batch-graph step bench.py's headline e2e times; summarised by
# tools/launchlist_summary.py.  The capture/warm-up launches come first; the last step is what the summary's --tail takes.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_batch.csv python tools/prof_pass.py cfg3 2 batch > gpurun_out/r2_ncu_batch.log 2>&1
wc -l gpurun_out/r2_launches_batch.csv; tail -3 gpurun_out/r2_ncu_batch.log; gzip -f -9 gpurun_out/r2_launches_batch.csv
