import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "hair-gs_b200"), os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)
import torch
import common
from hairgs_b200 import fused, models, scenes
from gaussian_renderer import render
dev = torch.device("cuda:0")
for (M, D, mod, W, H) in [(1, 0, 1.0, 384, 320), (4, 1, 1.0, 384, 320), (1, 0, 0.8, 384, 320), (1, 0, 1.0, 250, 190), (4, 0, 1.0, 384, 320)]:
    sc = scenes.strand_scene(300, 40, seed=8, sh_coeffs=M).to(dev)
    cam = scenes.orbit_cameras(4, W, H, device=dev)[1]
    torch.manual_seed(5)
    w7 = torch.randn(7, H, W, device=dev)
    bg7 = torch.tensor([0.1, 0.2, 0.3, 0.0, 0.5, 0.4, 0.6], device=dev)
    m1 = models.StrandModel(sc, sh_degree=D).to(dev)
    out = fused.render_strands(cam, m1, bg7, scaling_modifier=mod)
    (out["image7"] * w7).sum().backward()
    m2 = models.StrandModel(sc, sh_degree=D).to(dev)
    r_rgb = render(cam, m2, bg7[0:3], scaling_modifier=mod)
    r_mask = render(cam, m2, bg7[3:4].repeat(3), scaling_modifier=mod, override_color=m2.get_mask.repeat(1, 3))
    r_ori = render(cam, m2, bg7[4:7], scaling_modifier=mod, override_color=m2.get_orientation)
    ((r_rgb["render"] * w7[0:3]).sum() + (r_mask["render"][0:1] * w7[3:4]).sum() + (r_ori["render"] * w7[4:7]).sum()).backward()
    ref_img = torch.cat([r_rgb["render"], r_mask["render"][0:1], r_ori["render"]]).detach()
    diff = (out["image7"].detach() - ref_img).abs()
    print(f"M={M} D={D} mod={mod} {W}x{H}: pix>tol {int((diff > 1e-4).sum())} max {float(diff.max()):.2e} radii diff {int((out['radii'] != r_rgb['radii']).sum())}",
          {n: f"{common.rel_err(p1.grad, p2.grad):.1e}" for (n, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()) if p1.grad is not None and p1.numel()})
