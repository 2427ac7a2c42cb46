#!/bin/bash
# Round-2 end evidence on one B200: GPU suite, smoke(), both bench arms exactly as the driver runs them, cfg2 / cfg5 lines.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench reference"; timeout 1200 python bench.py --impl reference > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; tail -c 400 gpurun_out/r2_final_ref.json; tail -2 gpurun_out/r2_final_ref.err
echo "== bench ours"; timeout 1200 python bench.py > gpurun_out/r2_final_ours.json 2> gpurun_out/r2_final_ours.err; tail -3 gpurun_out/r2_final_ours.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_final_ours.json").read().strip().splitlines()[-1])
print("value", d["value"], d.get("value_eager"), "ms", d["ms_per_step"], d.get("ms_per_step_stats"), "e2e", d["e2e"]["value"], d["e2e"].get("value_eager"), d["e2e"].get("value_incl_optimizer"))
print("cpu", d.get("cpu_baseline")); print("roofline", d["roofline"]); print("clocks", d["clocks"], "launches", d["gpu_launches"])
print("parity", json.dumps(d.get("parity"))[:1200])
print("dropin", json.dumps(d.get("dropin"))[:600])
r = json.loads(open("gpurun_out/r2_final_ref.json").read().strip().splitlines()[-1])
print("ref value", r.get("value"), "e2e", r.get("e2e"), "fallback", r.get("fallback"), "kind", r.get("cpu_baseline", {}).get("kind"))
PY
for w in cfg2 cfg5; do
  echo "== $w"; timeout 900 python bench.py --workload $w --impl reference --no-cpu-baseline > gpurun_out/r2_final_ref_$w.json 2> gpurun_out/r2_final_ref_$w.err
  timeout 900 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2_final_ours_$w.json 2> gpurun_out/r2_final_ours_$w.err; tail -2 gpurun_out/r2_final_ours_$w.err
  python - <<PY
import json
for arm in ("ours", "ref"):
    try:
        d = json.loads(open("gpurun_out/r2_final_%s_$w.json" % arm).read().strip().splitlines()[-1])
        print(arm, "$w", "value", d.get("value"), "ms", d.get("ms_per_step"), "e2e", d.get("e2e", {}).get("value"), "roofline", d.get("roofline"), "parity_ok", (d.get("parity") or {}).get("ok"))
    except Exception as e:
        print(arm, "$w", "failed", e)
PY
done
