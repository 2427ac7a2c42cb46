#!/bin/bash
mkdir -p gpurun_out
echo "== batch graph test (2 comp branches)"; timeout 600 python -m pytest tests -m gpu -q -k "graphed_batch" 2>&1 | tail -3
for b in 1 2 3; do
echo "== HGS_BATCH_COMP_BRANCHES=$b"
HGS_BATCH_COMP_BRANCHES=$b timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_br_$b.json 2> gpurun_out/r2_br_$b.err; tail -3 gpurun_out/r2_br_$b.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_br_$b.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], d["ms_per_step_stats"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step_stats"], "opt", d["e2e"]["value_incl_optimizer"])
PY
done
