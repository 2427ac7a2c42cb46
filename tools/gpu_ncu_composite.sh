#!/bin/bash
# ncu --set full of the compositors (cfg3 fused strand step) in both block modes; reports land in gpurun_out/.
mkdir -p gpurun_out
for mode in 4x4 8x4; do
  HGS_COMPOSITE_BLOCKS=$mode timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"composite_(fwd|bwd)" -s 4 -c 2 -f -o gpurun_out/composite_$mode python tools/prof_pass.py cfg3 3 fused 2>&1 | tail -3
done
ls -la gpurun_out/*.ncu-rep
