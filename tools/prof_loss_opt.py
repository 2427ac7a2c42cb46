"""ncu driver for the N3/N4 kernels: a few fused image-loss evaluations and Adam steps at cfg3 size."""
import math
import os
import sys
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hair-gs_b200"))
from hairgs_b200 import losses, optim, scenes  # noqa: E402

dev = torch.device("cuda:0")
H = W = 1024
g = torch.Generator(device="cuda").manual_seed(0)
image7 = torch.rand(7, H, W, generator=g, device=dev, requires_grad=True)
gt_rgb = torch.rand(3, H, W, generator=g, device=dev)
gt_mask = (torch.rand(H, W, generator=g, device=dev) < 0.5).float()
gt_theta = torch.rand(H, W, generator=g, device=dev) * math.pi
conf = torch.rand(H, W, generator=g, device=dev)
rot = losses.view_rot_of(scenes.orbit_cameras(4, W, H, device=dev)[1].world_view_transform.float())
for _ in range(3):
    loss, _ = losses.hair_image_loss(image7, gt_rgb, gt_mask, gt_theta, conf, rot, orient_mask=gt_mask > 0.5)
shapes = {"_endpoints": (1000000, 3), "_width": (990000, 1), "_opacity": (990000, 1), "_mask": (990000, 1),
          "_features_dc": (990000, 1, 3)}
params = {k: torch.nn.Parameter(torch.randn(*s, device=dev)) for k, s in shapes.items()}
fa = optim.FlatAdam([{"params": [p], "lr": 1e-3, "name": k} for k, p in params.items()])
fa.grads.flat.normal_()
for _ in range(3):
    fa.step()
torch.cuda.synchronize()
print("done")
