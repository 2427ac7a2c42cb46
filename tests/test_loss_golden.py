"""The torch restatement of Hair-GS's image loss (hairgs_b200.losses.hair_image_loss_torch — the checker the GPU tests
compare the fused kernel with) against vectors recorded from the reference's OWN loss/losses.py functions
(tests/golden/make_loss_golden.py: l1_loss, ssim, bidirectional_angle_difference, and the autograd gradient of
0.8 l1 + 0.2 (1 - ssim))."""
import math
import os

import numpy as np
import pytest
import torch

from hairgs_b200 import losses

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_ref.npz"))
TAGS = ["a", "b", "c"]


def _case(tag, device):
    t = lambda k: torch.from_numpy(GOLD[f"{tag}_{k}"]).to(device)
    return t("img"), t("gt"), t("a1"), t("a2"), t("grad"), t("angle_diff"), float(GOLD[f"{tag}_ssim"]), float(GOLD[f"{tag}_l1"])


@pytest.mark.parametrize("tag", TAGS)
def test_torch_restatement_matches_reference_functions(tag):
    img, gt, a1, a2, grad, adiff, ssim, l1 = _case(tag, "cpu")
    H, W = img.shape[1:]
    x = img.clone().requires_grad_(True)
    zeros = torch.zeros(H, W)
    # orientation plane chosen so that theta == a1 exactly is not needed here: the angle term is checked separately below
    loss, terms = losses.hair_image_loss_torch(x, zeros, torch.zeros(3, H, W), gt, zeros, a2, zeros, torch.eye(4),
                                               lambda_dssim=0.2, lambda_mask=0.0, lambda_orientation=0.0,
                                               orient_mask=torch.ones(H, W, dtype=torch.bool))
    assert abs(float(terms["l1"]) - l1) <= 1e-7
    assert abs((1.0 - float(terms["dssim"])) - ssim) <= 1e-6
    loss.backward()
    assert torch.allclose(x.grad, grad, rtol=1e-5, atol=1e-9)
    # bidirectional_angle_difference (losses.py:87-103) as used inside the orientation term
    d = math.pi / 2 - torch.abs(torch.abs(a1 - a2) - math.pi / 2)
    assert torch.equal(d, adiff)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_fused_loss_matches_reference_vectors(tag):
    """hgs_hair_image_loss on the GPU against the reference's own l1_loss / ssim values and autograd gradient
    (terms rel 2e-5, gradient 1e-4 of its max)."""
    dev = torch.device("cuda:0")
    img, gt, a1, a2, grad, adiff, ssim, l1 = _case(tag, dev)
    H, W = img.shape[1:]
    # an orientation plane whose view-space angle is a1: o = (sin a1, cos a1, 0) with the identity view rotation
    orient = torch.stack([torch.sin(a1), torch.cos(a1), torch.zeros_like(a1)])
    image7 = torch.cat([img, torch.zeros(1, H, W, device=dev), orient]).contiguous().requires_grad_(True)
    conf = torch.ones(H, W, device=dev)
    rot = [1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0]
    loss, terms = losses.hair_image_loss(image7, gt, torch.zeros(H, W, device=dev), a2, conf, rot, lambda_dssim=0.2,
                                         lambda_mask=0.0, lambda_orientation=0.0)
    t = terms.tolist()
    assert abs(t[1] - l1) <= 2e-5 * max(1.0, l1)
    assert abs((1.0 - t[2]) - ssim) <= 2e-5
    # orientation term with unit confidence == mean bidirectional angle difference of the reference
    assert abs(t[4] - float(adiff.mean())) <= 1e-5
    loss.backward()
    g = image7.grad[:3]
    assert float((g - grad).abs().max()) <= 1e-4 * float(grad.abs().max())
