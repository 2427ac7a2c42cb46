#!/bin/bash
# Round-end evidence on one B200: parity suite, smoke(), both bench arms exactly as the driver runs them, ncu launch list.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench reference"; timeout 900 python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; tail -c 600 gpurun_out/final_ref.json
echo "== bench ours"; timeout 900 python bench.py > gpurun_out/final_ours.json 2> gpurun_out/final_ours.err; tail -3 gpurun_out/final_ours.err
python - <<PY
import json
d = json.loads(open("gpurun_out/final_ours.json").read().strip().splitlines()[-1])
print("value", d["value"], d.get("value_eager"), "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("value_eager"), d["e2e"].get("value_incl_optimizer"))
print("cpu", d.get("cpu_baseline")); print("roofline", d["roofline"]); print("clocks", d["clocks"], "launches", d["gpu_launches"])
PY
echo "== ncu launch list of bench.py"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches_bench.csv
