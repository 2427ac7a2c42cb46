#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
timeout 600 python bench.py --steps 48 --warmup 8 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err
tail -5 gpurun_out/bench4.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench4.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", json.dumps(d["e2e"], indent=1))
PY
