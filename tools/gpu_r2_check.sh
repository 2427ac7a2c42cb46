#!/bin/bash
# round-2 check on one B200: parity suite, smoke, sort A/B vs CUB, both bench arms (short)
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== sort A/B (ballot ranking)"; timeout 300 python tools/sort_bench.py cfg3 cfg5 2>&1 | tee gpurun_out/sort_ab.jsonl | tail -4
echo "== sort A/B (match.any ranking)"; HGS_SORT_RANK=match timeout 300 python tools/sort_bench.py cfg3 cfg5 2>&1 | tee gpurun_out/sort_ab_match.jsonl | tail -4
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_ref.json 2> gpurun_out/r2_ref.err; tail -5 gpurun_out/r2_ref.err; tail -c 1500 gpurun_out/r2_ref.json
echo "== bench ours"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_ours.json 2> gpurun_out/r2_ours.err; tail -5 gpurun_out/r2_ours.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_ours.json").read().strip().splitlines()[-1])
print("value", d["value"], d.get("value_eager"), "ms", d["ms_per_step"], d.get("ms_per_step_stats"), "e2e", d["e2e"]["value"], d["e2e"].get("value_eager"), d["e2e"].get("value_incl_optimizer"))
print("dropin", d.get("dropin")); print("binning", d.get("binning_chain_ms_per_pass"))
print("stages", {k: v["ms_per_launch"] for k, v in d["stages"].items()})
print("parity", json.dumps(d.get("parity")))
print("cpu", d.get("cpu_baseline", {}).get("value")); print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"]); print("clocks", d["clocks"], "launches", d["gpu_launches"])
PY
