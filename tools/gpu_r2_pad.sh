#!/bin/bash
mkdir -p gpurun_out
for pad in 0 30000 70000; do
echo "== HGS_FWD_SMEM_PAD=$pad"
HGS_FWD_SMEM_PAD=$pad timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_pad_$pad.json 2> gpurun_out/r2_pad_$pad.err; tail -3 gpurun_out/r2_pad_$pad.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_pad_$pad.json").read().strip().splitlines()[-1])
print("value", d["value"], "eager", d.get("value_eager"), "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "fwd", d["stages"]["composite_fwd"]["ms_per_launch"])
PY
done
