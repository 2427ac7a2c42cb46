#!/bin/bash
# block-mask walk of the half-warp compositors (HGS_WALK=mask default | extent): GPU suite on the default, then A/B of bench.py
mkdir -p gpurun_out
echo "== full GPU suite (mask walk)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12
echo "== parity subset, HGS_WALK=extent"; HGS_WALK=extent timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not full_size" 2>&1 | tail -5
for m in mask extent; do
echo "== HGS_WALK=$m"
HGS_WALK=$m timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_walk_$m.json 2> gpurun_out/r2_walk_$m.err; tail -3 gpurun_out/r2_walk_$m.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_walk_$m.json").read().strip().splitlines()[-1])
print("value", d["value"], "eager", d.get("value_eager"), "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "opt", d["e2e"].get("value_incl_optimizer"), "dropin", d.get("dropin", {}).get("value"))
print("stages", {k: round(v["ms_per_launch"] * 1000, 1) for k, v in d["stages"].items()})
print("parity", json.dumps(d.get("parity"))[:600])
PY
done
