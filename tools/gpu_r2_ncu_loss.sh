#!/bin/bash
# ncu --set full of the image-loss kernels inside the end-to-end batch step
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ssim|pointwise|unpack" -s 24 -c 3 -f -o gpurun_out/r2_loss python tools/prof_pass.py cfg3 2 batch 2>&1 | tail -2
ls -la gpurun_out/r2_loss.ncu-rep
