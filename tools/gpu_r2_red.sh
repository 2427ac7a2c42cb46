#!/bin/bash
mkdir -p gpurun_out
echo "== tests (vector red default)"; timeout 900 python -m pytest tests -m gpu -q -k "fused or graphed or hair_image_loss_on" 2>&1 | tail -4
for m in vec scalar; do
echo "== HGS_BWD_RED=$m"
HGS_BWD_RED=$m timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_red_$m.json 2> gpurun_out/r2_red_$m.err; tail -3 gpurun_out/r2_red_$m.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_red_$m.json").read().strip().splitlines()[-1])
print("value", d["value"], "eager", d.get("value_eager"), "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "bwd", d["stages"]["composite_bwd"]["ms_per_launch"], "pre_bwd", d["stages"]["preprocess_bwd"]["ms_per_launch"], "pre_fwd", d["stages"]["preprocess_fwd"]["ms_per_launch"])
PY
done
