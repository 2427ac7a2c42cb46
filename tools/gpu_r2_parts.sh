#!/bin/bash
# upper/lower bounds of the two-branch batch graph: each branch alone vs both (resident loop only)
mkdir -p gpurun_out
cat > /tmp/parts.py <<'PY'
import os, sys, json, math, torch
ROOT = os.getcwd()
sys.path[:0] = [os.path.join(ROOT, "hair-gs_b200"), os.path.join(ROOT, "tests")]
from hairgs_b200 import fused, graphs, models, scenes
dev = torch.device("cuda:0")
H = W = 1024; V = 8
sc = scenes.strand_scene(10000, 100, seed=0).to(dev)
cams = scenes.orbit_cameras(16, W, H, device=dev)
model = models.StrandModel(sc).to(dev)
names = {"endpoints": "_endpoints", "width": "_width", "opacity": "_opacity", "mask": "_mask", "features": "_features_dc"}
sink = fused.GradSink({k: torch.zeros_like(getattr(model, n)) for k, n in names.items()})
bg7 = torch.zeros(7, device=dev)
g = torch.Generator().manual_seed(1)
dL7 = (torch.randn(7, H, W, generator=g) / (H * W)).to(dev)
cap, bits = graphs.measure_plan(model, cams, bg7)
flat = lambda c: torch.cat([c.world_view_transform.reshape(-1), c.full_proj_transform.reshape(-1), c.camera_center.reshape(-1)])
res = {}
for parts in ("both", "bin", "both", "comp"):
    os.environ["HGS_BATCH_PARTS"] = "both"
    b = graphs.GraphedStrandBatch(model, sink, bg7, H, W, cams[0].FoVx, cams[0].FoVy, cap, bits, V, dimage=dL7)
    for v in range(V): b.cam_buf[v].copy_(flat(cams[v]))
    torch.cuda.synchronize()
    b._batch(False)            # full warm-up so that the comp-only variant finds binned workspaces
    torch.cuda.synchronize()
    os.environ["HGS_BATCH_PARTS"] = parts
    b.capture(warmup=1, accumulate_variant=False)
    for _ in range(5): b.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): b.replay()
    e1.record(); torch.cuda.synchronize()
    res.setdefault(parts, []).append(round(e0.elapsed_time(e1) / 20, 4))
    del b
print(json.dumps({"ms_per_8view_step": res}))
PY
python /tmp/parts.py 2>&1 | tail -3
