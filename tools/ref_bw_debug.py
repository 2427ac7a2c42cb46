import faulthandler, os, sys
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "hair-gs_b200"), os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)
import torch
import refload
mode = sys.argv[1] if len(sys.argv) > 1 else "refonly"
if mode == "ours_first":
    from hairgs_b200 import _lib
    _lib.load()
import common
dev = torch.device("cuda:0")
d = common.blob_inputs(5000, 128, 128, dev)
C = refload.ref_dgr()
N, color, radii, geom, binning, img = C.rasterize_gaussians(*common.fwd_args(d))
torch.cuda.synchronize()
print("fwd ok", N, flush=True)
dL = torch.randn_like(color)
g = C.rasterize_gaussians_backward(*common.bwd_args(d, radii, dL, geom, N, binning, img))
torch.cuda.synchronize()
print("bwd ok", [float(x.abs().max()) for x in g if x.numel()], flush=True)
