#!/bin/bash
mkdir -p gpurun_out
for pad in 0 30000; do
echo "== HGS_FWD_SMEM_PAD=$pad (2 comp branches)"
HGS_FWD_SMEM_PAD=$pad timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_pad2_$pad.json 2> gpurun_out/r2_pad2_$pad.err; tail -3 gpurun_out/r2_pad2_$pad.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_pad2_$pad.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "opt", d["e2e"]["value_incl_optimizer"])
PY
done
