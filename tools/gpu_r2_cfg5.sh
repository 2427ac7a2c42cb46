#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload cfg5 --no-cpu-baseline > gpurun_out/r2_final_ours_cfg5.json 2> gpurun_out/r2_final_ours_cfg5.err; tail -3 gpurun_out/r2_final_ours_cfg5.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_final_ours_cfg5.json").read().strip().splitlines()[-1])
print("cfg5 value", d.get("value"), "ms", d.get("ms_per_step"), "e2e", json.dumps(d.get("e2e"))[:900])
PY
