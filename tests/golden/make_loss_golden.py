"""Golden vectors for the image-space loss, recorded by IMPORTING the reference's own
loss/losses.py in the build container (CPU) and running its pure-torch functions on seeded inputs:
  l1_loss (:16-17), ssim (:43-84), bidirectional_angle_difference (:87-103).
The module's other imports (pytorch3d, the CUDA rasterizer, scene.*, c_utils) are not installed here and are not
needed by these three functions, so they are stubbed before the import.  orientation_loss_rast / mask_loss_rast call
the CUDA renderer and cannot run here; their post-render arithmetic is restated in
hairgs_b200.losses.hair_image_loss_torch and pinned through these pieces (+ torch's own BCEWithLogitsLoss).

    python tests/golden/make_loss_golden.py      # writes tests/golden/loss_ref.npz (needs /root/reference)
"""
import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_losses():
    for name, attrs in {"pytorch3d": [], "pytorch3d.ops": ["knn_points"], "gaussian_renderer": ["render"],
                        "scene": [], "scene.cameras": ["Camera"],
                        "scene.hair_gaussian_model": ["HairGaussianModel", "GaussianModel"],
                        "c_utils": ["filter_strand_list_segments"]}.items():
        m = types.ModuleType(name)
        for a in attrs:
            setattr(m, a, type(a, (), {}))
        sys.modules.setdefault(name, m)
    spec = importlib.util.spec_from_file_location("ref_losses", os.path.join(REF, "loss", "losses.py"))
    mod = importlib.util.module_from_spec(spec)
    # two function signatures build a default `bg` tensor on "cuda" at import time (losses.py:228,296); there is no GPU
    # in the build container, so device="cuda" falls back to the CPU for the duration of the import only
    real_tensor = torch.tensor

    def tensor_cpu(*a, **k):
        if str(k.get("device", "")).startswith("cuda") and not torch.cuda.is_available():
            k["device"] = "cpu"
        return real_tensor(*a, **k)
    torch.tensor = tensor_cpu
    try:
        spec.loader.exec_module(mod)
    finally:
        torch.tensor = real_tensor
    return mod


def main():
    ref = load_reference_losses()
    g = torch.Generator().manual_seed(20261017)
    out = {}
    for tag, (H, W) in {"a": (40, 52), "b": (17, 33), "c": (64, 64)}.items():
        base = torch.nn.functional.interpolate(torch.rand(1, 3, H // 4 + 2, W // 4 + 2, generator=g), size=(H, W),
                                               mode="bilinear")[0]
        img = (base + 0.1 * torch.rand(3, H, W, generator=g)).clamp(0, 1)
        gt = (base + 0.1 * torch.rand(3, H, W, generator=g)).clamp(0, 1)
        a1 = torch.rand(H, W, generator=g) * math.pi
        a2 = torch.rand(H, W, generator=g) * math.pi
        img_r = img.clone().requires_grad_(True)
        s = ref.ssim(img_r, gt)
        l1 = ref.l1_loss(img_r, gt)
        (0.8 * l1 + 0.2 * (1.0 - s)).backward()
        out.update({f"{tag}_img": img.numpy(), f"{tag}_gt": gt.numpy(), f"{tag}_a1": a1.numpy(), f"{tag}_a2": a2.numpy(),
                    f"{tag}_ssim": np.float64(s.item()), f"{tag}_l1": np.float64(l1.item()),
                    f"{tag}_grad": img_r.grad.numpy(),
                    f"{tag}_angle_diff": ref.bidirectional_angle_difference(a1, a2).numpy()})
    np.savez_compressed(os.path.join(HERE, "loss_ref.npz"), **out)
    print("wrote loss_ref.npz:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.startswith("a_")})


if __name__ == "__main__":
    main()
