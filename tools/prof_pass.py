"""Minimal driver for ncu captures: a few forward+backward rasterization passes of one workload through the
drop-in `_C` entry points (no e2e glue, no CPU baseline).  Usage: python tools/prof_pass.py cfg3 3 [ref|fused|batch]
`batch`: the end-to-end step bench.py's headline `e2e` times (H2D of 8 views' targets, unpack, ONE batch-graph replay with the
image loss, D2H of the losses) — 2 untimed steps to page everything in, then `iters` steps."""
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
batch = len(sys.argv) > 3 and sys.argv[3] == "batch"
args = types.SimpleNamespace(views_per_step=8 if batch else 1, gpus=1, graph_mode="batch" if batch else "view")
h = bench.Harness(bench.WORKLOADS[name], args, dev, backend, 1, 0)
fused = len(sys.argv) > 3 and sys.argv[3] == "fused"
if fused:
    h.setup_fused()
if batch:
    h.setup_e2e(fused=True, graph=True)
    assert h.ebatch is not None, h.graph_note
    for it in range(iters + 2):
        h.step_e2e(it)
        torch.cuda.synchronize()
    print("done batch", name, "P", h.P)
    sys.exit(0)
for it in range(iters):
    (h.step_resident_fused if fused else h.step_resident)(it)
torch.cuda.synchronize()
print("done", name, "P", h.P, "N", h.last_N)
