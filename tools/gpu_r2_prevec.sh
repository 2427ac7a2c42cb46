#!/bin/bash
# A/B of the 128-bit [P,3] row loads of preprocess_fwd (generic entry, cfg2: 300 k Gaussians, SH degree 3)
mkdir -p gpurun_out
echo "== parity (generic entry)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "forward_and_backward or golden or tiny or cfg2 or python_switches or autograd_surface" 2>&1 | tail -3
M=gpu__time_duration.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed.sum,dram__bytes_read.sum
for v in 1 0; do
echo "== HGS_PRE_VEC=$v"
HGS_PRE_VEC=$v timeout 600 ncu --metrics $M --clock-control none -k regex:"preprocess_fwd" -s 2 -c 2 --csv python tools/prof_pass.py cfg2 6 2>/dev/null | python -c "
import csv, sys
for r in csv.reader(sys.stdin):
    if len(r) > 10 and r[0].isdigit():
        print(r[0], r[4][:40], r[-3], r[-1])
"
done
