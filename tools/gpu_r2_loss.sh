#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "loss or graphed or fused" 2>&1 | tail -4
timeout 300 python tools/loss_kernels_bench.py
timeout 300 python tools/loss_kernels_bench.py
