#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for r in 1 2; do
  timeout 600 python bench.py --steps 48 --warmup 8 --no-cpu-baseline > gpurun_out/bench3_4x4_$r.json 2> gpurun_out/bench3_4x4_$r.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench3_4x4_$r.json").read().strip().splitlines()[-1])
print("4x4 run $r value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("value_incl_optimizer"),
      {k: round(v["ms_per_launch"], 4) for k, v in d["stages"].items() if k.startswith("composite")})
PY
done
echo "== e2e host time"; timeout 300 python tools/e2e_host_time.py cfg3 2>&1 | head -45 | tee gpurun_out/e2e_host_time.log
for w in cfg5 cfg2; do
  timeout 600 python bench.py --workload $w --steps 32 --warmup 5 --no-cpu-baseline > gpurun_out/bench3_$w.json 2> gpurun_out/bench3_$w.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench3_$w.json").read().strip().splitlines()[-1])
print("$w value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"],
      {k: round(v["ms_per_launch"], 4) for k, v in d["stages"].items() if k.startswith("composite")})
PY
done
