#!/bin/bash
# MUFU.RCP in the backward recurrence: gradient tests + bench
mkdir -p gpurun_out
echo "== GPU suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_ssim.json 2> gpurun_out/r2_ssim.err; tail -3 gpurun_out/r2_ssim.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_ssim.json").read().strip().splitlines()[-1])
print("value", d["value"], "eager", d.get("value_eager"), "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "opt", d["e2e"].get("value_incl_optimizer"), "dropin", d.get("dropin", {}).get("value"))
print("stages", {k: round(v["ms_per_launch"] * 1000, 1) for k, v in d["stages"].items()})
PY
cat gpurun_out/parity_cfg3_fused.json | head -c 1500
