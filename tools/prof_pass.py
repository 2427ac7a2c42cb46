"""Minimal driver for ncu captures: a few forward+backward rasterization passes of one workload through the
drop-in `_C` entry points (no e2e glue, no CPU baseline).  Usage: python tools/prof_pass.py cfg3 3 [ref]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "hair-gs_b200"), os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
use_ref = len(sys.argv) > 3 and sys.argv[3] == "ref"
dev = torch.device("cuda:0")
if use_ref:
    import refload
    backend = refload.ref_dgr()
else:
    import diff_gaussian_rasterization._C as backend
import types  # noqa: E402
args = types.SimpleNamespace(views_per_step=1, gpus=1, graph_mode="view")
h = bench.Harness(bench.WORKLOADS[name], args, dev, backend, 1, 0)
fused = len(sys.argv) > 3 and sys.argv[3] == "fused"
if fused:
    h.setup_fused()
for it in range(iters):
    (h.step_resident_fused if fused else h.step_resident)(it)
torch.cuda.synchronize()
print("done", name, "P", h.P, "N", h.last_N)
