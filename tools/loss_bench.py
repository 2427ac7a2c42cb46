"""Time the fused Hair-GS image loss (hgs_hair_image_loss) against the torch composition at 1024^2 (value + backward)."""
import math
import os
import sys
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hair-gs_b200"))
from hairgs_b200 import losses, scenes  # noqa: E402

dev = torch.device("cuda:0")
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = torch.Generator(device="cuda").manual_seed(0)
image7 = torch.rand(7, H, W, generator=g, device=dev)
gt_rgb = torch.rand(3, H, W, generator=g, device=dev)
gt_mask = (torch.rand(H, W, generator=g, device=dev) < 0.5).float()
gt_theta = torch.rand(H, W, generator=g, device=dev) * math.pi
conf = torch.rand(H, W, generator=g, device=dev)
omask = torch.rand(H, W, generator=g, device=dev) < 0.5
wvt = scenes.orbit_cameras(4, W, H, device=dev)[1].world_view_transform.float().contiguous()
rot = losses.view_rot_of(wvt)


def fused():
    a = image7.clone().requires_grad_(True)
    loss, _ = losses.hair_image_loss(a, gt_rgb, gt_mask, gt_theta, conf, rot, orient_mask=omask)
    loss.backward()


def composed():
    a = image7.clone().requires_grad_(True)
    loss, _ = losses.hair_image_loss_torch(a[:3], a[3], a[4:7], gt_rgb, gt_mask, gt_theta, conf, wvt, orient_mask=omask)
    loss.backward()


for name, fn in (("fused", fused), ("torch", composed)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 20:.3f} ms per loss+backward at {H}x{W}")
