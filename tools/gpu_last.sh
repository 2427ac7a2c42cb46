#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_parity.py -x -q -k "graphed" 2>&1 | tail -3
timeout 150 python bench.py --no-cpu-baseline > gpurun_out/bench_vps8.json 2> gpurun_out/bench_vps8.err; tail -2 gpurun_out/bench_vps8.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_vps8.json").read().strip().splitlines()[-1])
print("value", d["value"], d.get("value_eager"), "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("value_eager"), d["e2e"].get("value_incl_optimizer"), d["e2e"].get("graph"), d["gpu_launches"], d["config"].get("views_per_step"))
PY
timeout 60 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
