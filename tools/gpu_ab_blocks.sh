#!/bin/bash
# One GPU-box session: parity suite (both compositor block shapes), A/B bench of 8x4 vs 4x4, sanitizer on the new kernels.
# Usage (from the repo root): gpurun --timeout 1500 -- 'bash tools/gpu_ab_blocks.sh'
mkdir -p gpurun_out
set -o pipefail
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for r in 1 2; do
  for mode in 8x4 4x4; do
    echo "== bench $mode run $r"
    HGS_COMPOSITE_BLOCKS=$mode timeout 600 python bench.py --steps 48 --warmup 8 --no-cpu-baseline > gpurun_out/bench_${mode}_$r.json 2> gpurun_out/bench_${mode}_$r.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${mode}_$r.json").read().strip().splitlines()[-1])
    print("$mode", "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"],
          {k: round(v["ms_per_launch"], 4) for k, v in d["stages"].items() if k.startswith("composite")})
except Exception as e:
    print("bench $mode failed:", e)
PY
  done
done
echo "== memcheck (new kernels)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "block_shapes or multichannel" 2>&1 | tail -8 | tee gpurun_out/memcheck_blocks.log
echo "== racecheck (new kernels)"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "block_shapes" 2>&1 | tail -8 | tee gpurun_out/racecheck_blocks.log
