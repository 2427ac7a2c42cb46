#!/bin/bash
mkdir -p gpurun_out
echo "== parity with the bulk-staged forward"; HGS_FWD_STAGING=bulk timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "forward_and_backward or golden or tiny or multichannel or block_shapes" 2>&1 | tail -6
for st in regs bulk; do
echo "== HGS_FWD_STAGING=$st"
HGS_FWD_STAGING=$st timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_stage_$st.json 2> gpurun_out/r2_stage_$st.err; tail -3 gpurun_out/r2_stage_$st.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_stage_$st.json").read().strip().splitlines()[-1])
print("value", d["value"], "eager", d.get("value_eager"), "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "fwd us", d["stages"]["composite_fwd"]["ms_per_launch"], "dropin", d["dropin"]["value"])
PY
done
