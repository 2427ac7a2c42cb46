#!/bin/bash
mkdir -p gpurun_out
echo "== batch graph test"; timeout 600 python -m pytest tests -m gpu -q -k "graphed" 2>&1 | tail -4
for prio in 1 0; do
echo "== HGS_GRAPH_PRIORITY=$prio"
HGS_GRAPH_PRIORITY=$prio timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_prio_$prio.json 2> gpurun_out/r2_prio_$prio.err; tail -3 gpurun_out/r2_prio_$prio.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_prio_$prio.json").read().strip().splitlines()[-1])
print("value", d["value"], "eager", d.get("value_eager"), "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "opt", d["e2e"]["value_incl_optimizer"], d["e2e"].get("graph","")[:60])
PY
done
echo "== priority + fwd pad 30000"
HGS_FWD_SMEM_PAD=30000 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_prio_pad.json 2> gpurun_out/r2_prio_pad.err; tail -3 gpurun_out/r2_prio_pad.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_prio_pad.json").read().strip().splitlines()[-1])
print("value", d["value"], "eager", d.get("value_eager"), "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "opt", d["e2e"]["value_incl_optimizer"])
PY
