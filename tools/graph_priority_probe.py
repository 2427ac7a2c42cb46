"""Diagnostic: which priority do the kernel nodes of the captured two-branch batch graph carry?  (cuda-python, tool only)"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "hair-gs_b200"), os.path.join(ROOT, "tests")]
import torch
from cuda.bindings import runtime as rt
from hairgs_b200 import fused, graphs, models, scenes
dev = torch.device("cuda:0")
H, W, V = 256, 256, 2
sc = scenes.strand_scene(300, 40, seed=0).to(dev)
cams = scenes.orbit_cameras(4, W, H, device=dev)
model = models.StrandModel(sc).to(dev)
names = {"endpoints": "_endpoints", "width": "_width", "opacity": "_opacity", "mask": "_mask", "features": "_features_dc"}
sink = fused.GradSink({k: torch.zeros_like(getattr(model, n)) for k, n in names.items()})
bg7 = torch.zeros(7, device=dev)
dL7 = torch.randn(7, H, W, device=dev)
cap, bits = graphs.measure_plan(model, cams, bg7)
print("stream priority range", rt.cudaDeviceGetStreamPriorityRange())
b = graphs.GraphedStrandBatch(model, sink, bg7, H, W, cams[0].FoVx, cams[0].FoVy, cap, bits, V, dimage=dL7)
print("bin stream priority", b.bin_stream.priority)
flat = lambda c: torch.cat([c.world_view_transform.reshape(-1), c.full_proj_transform.reshape(-1), c.camera_center.reshape(-1)])
for v in range(V): b.cam_buf[v].copy_(flat(cams[v]))
b.capture(accumulate_variant=False)
g = b.graphs[False]
raw = g.raw_cuda_graph()
err, _, n = rt.cudaGraphGetNodes(raw, 0)
err, nodes, n = rt.cudaGraphGetNodes(raw, n)
print("nodes", n, err)
cnt = collections.Counter()
for node in nodes:
    err, ty = rt.cudaGraphNodeGetType(node)
    if ty == rt.cudaGraphNodeType.cudaGraphNodeTypeKernel:
        err, val = rt.cudaGraphKernelNodeGetAttribute(node, rt.cudaLaunchAttributeID.cudaLaunchAttributePriority)
        cnt[("kernel", int(val.priority), str(err))] += 1
    else:
        cnt[(str(ty), None, "")] += 1
for k, v in sorted(cnt.items(), key=str): print(k, v)
