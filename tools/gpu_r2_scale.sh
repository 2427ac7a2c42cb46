#!/bin/bash
# multi-GPU check as the driver launches it: torchrun, one rank per GPU
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_scale_$N.json 2> gpurun_out/r2_scale_$N.err
tail -5 gpurun_out/r2_scale_$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_scale_$N.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value", d["value"], "eager", d.get("value_eager"), "ms", d["ms_per_step"], d.get("ms_per_step_stats"), "e2e", d["e2e"]["value"], "opt", d["e2e"]["value_incl_optimizer"], "dropin", d.get("dropin"))
print(d.get("allreduce")); print("allreduce_ms", d.get("allreduce_ms")); print("per rank", d.get("ms_per_step_stats", {}).get("per_rank_ms_per_step"), "e2e per rank", d["e2e"].get("ms_per_step_stats", {}).get("per_rank_ms_per_step")); print(d["config"]["views"], d["e2e"].get("graph"))
PY
