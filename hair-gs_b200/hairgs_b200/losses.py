"""Image-space losses on the device (SURVEY.md §8f row N3 — first piece).

weighted_l1(image[C,H,W], target[C,H,W], weights[C]) = sum_c weights[c] * sum |image_c - target_c| with its gradient
produced in the same kernel (hgs_weighted_l1).  `l1_groups` builds the weights that make it equal to a sum of Hair-GS
l1_loss terms (loss/losses.py:16-17: mean absolute error) over channel groups, e.g. RGB, mask, orientation of the fused
strand pass:  l1_groups([(0, 3, 1.0), (3, 4, 0.01), (4, 7, 1.0)], H, W).
"""
import torch

from . import _lib as L


class _WeightedL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, target, weights):
        lib = L.load()
        if not image.is_cuda:
            raise L.HgsError("weighted_l1: image must be a CUDA tensor (no CPU path)")
        dev = image.device
        img = L.f32c(image, "image", dev)
        tgt = L.f32c(target, "target", dev)
        w = L.f32c(weights, "weights", dev)
        if img.shape != tgt.shape or img.dim() != 3 or w.numel() != img.shape[0]:
            raise L.HgsError("weighted_l1: image/target must be [C,H,W] and weights [C]")
        C, HW = img.shape[0], img.shape[1] * img.shape[2]
        loss = torch.empty((), dtype=torch.float32, device=dev)
        grad = torch.empty_like(img)
        with torch.cuda.device(dev):
            L.check(lib.hgs_weighted_l1(C, HW, img.data_ptr(), tgt.data_ptr(), w.data_ptr(), loss.data_ptr(),
                                        grad.data_ptr(), L.stream_ptr(dev)), "weighted_l1")
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        return grad * grad_out, None, None


def weighted_l1(image, target, weights):
    return _WeightedL1.apply(image, target, weights)


def l1_groups(groups, H, W, device):
    """weights[C] such that weighted_l1 == sum over (c0, c1, lam) of lam * mean(|image[c0:c1] - target[c0:c1]|)."""
    C = max(c1 for _, c1, _ in groups)
    w = torch.zeros(C, dtype=torch.float32)
    for c0, c1, lam in groups:
        w[c0:c1] = lam / ((c1 - c0) * H * W)
    return w.to(device)
