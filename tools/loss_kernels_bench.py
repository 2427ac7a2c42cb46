"""Time the fused Hair-GS image loss launches alone (hgs_hair_image_loss: ssim_fwd, ssim_bwd, pointwise, finish) at 1024^2,
L2 flushed between calls by rotating over 8 input sets (8 x 60 MB > 126 MB)."""
import math
import os
import sys
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hair-gs_b200"))
from hairgs_b200 import losses, scenes  # noqa: E402

dev = torch.device("cuda:0")
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = torch.Generator(device="cuda").manual_seed(0)
sets = []
for _ in range(8):
    sets.append(dict(image7=torch.rand(7, H, W, generator=g, device=dev), gt_rgb=torch.rand(3, H, W, generator=g, device=dev),
                     gt_mask=(torch.rand(H, W, generator=g, device=dev) < 0.5).float(),
                     gt_theta=torch.rand(H, W, generator=g, device=dev) * math.pi, conf=torch.rand(H, W, generator=g, device=dev),
                     omask=torch.rand(H, W, generator=g, device=dev) < 0.5))
wvt = scenes.orbit_cameras(4, W, H, device=dev)[1].world_view_transform.float().contiguous()
rot = losses.view_rot_of(wvt)
lam = dict(lambda_dssim=0.2, lambda_mask=0.01, lambda_orientation=100.0)


def call(k):
    s = sets[k % 8]
    return losses.hair_image_loss(s["image7"], s["gt_rgb"], s["gt_mask"], s["gt_theta"], s["conf"], rot, orient_mask=s["omask"], **lam)


for k in range(10):
    call(k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(80):
    call(k)
e1.record()
torch.cuda.synchronize()
print(f"hair_image_loss (value + dL/dimage): {e0.elapsed_time(e1) / 80 * 1000:.1f} us per call at {H}x{W}")
