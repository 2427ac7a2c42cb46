#!/bin/bash
# quick perf iteration on one B200: selected tests + our bench arm without the CPU legs
mkdir -p gpurun_out
echo "== pytest (selected: $1)"; timeout 900 python -m pytest tests -m gpu -q -k "$1" 2>&1 | tail -12
echo "== bench ours"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $2 > gpurun_out/r2_quick.json 2> gpurun_out/r2_quick.err; tail -5 gpurun_out/r2_quick.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_quick.json").read().strip().splitlines()[-1])
print("value", d["value"], "eager", d.get("value_eager"), "ms", d["ms_per_step"], d.get("ms_per_step_stats"))
print("e2e", d["e2e"]["value"], "eager", d["e2e"].get("value_eager"), "opt", d["e2e"].get("value_incl_optimizer"), d["e2e"].get("graph"))
print("value_path", d.get("value_path"))
print("dropin", d.get("dropin", {}).get("value"), d.get("dropin", {}).get("e2e")); print("binning", d.get("binning_chain_ms_per_pass"))
print("stages", {k: v["ms_per_launch"] for k, v in d["stages"].items()})
print("clocks", d["clocks"], "launches", d["gpu_launches"])
PY
