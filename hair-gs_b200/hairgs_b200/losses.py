"""Image-space losses on the device (SURVEY.md §8f row N3 — first piece).

weighted_l1(image[C,H,W], target[C,H,W], weights[C]) = sum_c weights[c] * sum |image_c - target_c| with its gradient
produced in the same kernel (hgs_weighted_l1).  `l1_groups` builds the weights that make it equal to a sum of Hair-GS
l1_loss terms (loss/losses.py:16-17: mean absolute error) over channel groups, e.g. RGB, mask, orientation of the fused
strand pass:  l1_groups([(0, 3, 1.0), (3, 4, 0.01), (4, 7, 1.0)], H, W).
"""
import ctypes

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


def pack_targets(gt_rgb, gt_mask, gt_theta, confidence):
    """Host-side storage format of one view's targets (what a data loader keeps in pinned memory): rgbm uint8 [H,W,4]
    (image bytes + mask 0/255) and theta_conf float16 [H,W,2] — 8 bytes per pixel instead of the 24 of six float planes."""
    rgb = (gt_rgb.clamp(0, 1) * 255.0).round().to(torch.uint8)
    m = (gt_mask > 0.5).to(torch.uint8) * 255
    rgbm = torch.cat([rgb, m[None]], 0).permute(1, 2, 0).contiguous()
    tc = torch.stack([gt_theta, confidence], -1).to(torch.float16).contiguous()
    return rgbm, tc


def unpack_targets(rgbm, theta_conf, out=None):
    """Device side of pack_targets (hgs_unpack_targets): rgbm [V,H,W,4] / [H,W,4] uint8 and theta_conf [V,H,W,2] / [H,W,2]
    float16 CUDA tensors -> float32 [V,6,H,W] / [6,H,W] = (r, g, b, mask, theta, confidence) planes, one launch."""
    lib = L.load()
    if not rgbm.is_cuda or not theta_conf.is_cuda:
        raise L.HgsError("unpack_targets: CUDA tensors only (no CPU path)")
    dev = rgbm.device
    batched = rgbm.dim() == 4
    V = rgbm.shape[0] if batched else 1
    H, W = rgbm.shape[-3], rgbm.shape[-2]
    if rgbm.dtype != torch.uint8 or rgbm.shape[-1] != 4 or theta_conf.dtype != torch.float16 or \
            tuple(theta_conf.shape) != tuple(rgbm.shape[:-1]) + (2,) or not rgbm.is_contiguous() or not theta_conf.is_contiguous():
        raise L.HgsError("unpack_targets: expected contiguous rgbm uint8 [...,H,W,4] and theta_conf float16 [...,H,W,2]")
    shape = (V, 6, H, W) if batched else (6, H, W)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=dev)
    elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev:
        raise L.HgsError("unpack_targets: out must be a contiguous float32 tensor of shape " + str(shape))
    with torch.cuda.device(dev):
        L.check(lib.hgs_unpack_targets(V, H * W, rgbm.data_ptr(), theta_conf.data_ptr(), out.data_ptr(), L.stream_ptr(dev)),
                "unpack_targets")
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Hair-GS image-space loss of one view (loss/losses.py:319-346, image terms), fused: hgs_hair_image_loss
# ---------------------------------------------------------------------------------------------------------------------
def hair_image_loss_raw(image7, gt_rgb, gt_mask, gt_theta, confidence, orient_mask, view_rot, lambdas, bg_orient):
    """The fused loss without autograd: -> (terms[8], dL_dimage7[7,H,W]) for d(total)/d(image7); lambdas = (l1, dssim,
    mask, orientation) weights.  See hair_image_loss for the arguments."""
    lib = L.load()
    if not image7.is_cuda:
        raise L.HgsError("hair_image_loss: image7 must be a CUDA tensor (no CPU path)")
    dev = image7.device
    img = L.f32c(image7, "image7", dev)
    if img.dim() != 3 or img.shape[0] != 7:
        raise L.HgsError("hair_image_loss: image7 must be [7,H,W] (rgb | mask | orientation)")
    H, W = img.shape[1], img.shape[2]
    gt = L.f32c(gt_rgb, "gt_rgb", dev)
    gm = L.f32c(gt_mask, "gt_mask", dev)
    th = L.f32c(gt_theta, "gt_theta", dev)
    cf = L.f32c(confidence, "confidence", dev)
    if gt.shape != (3, H, W) or gm.shape != (H, W) or th.shape != (H, W) or cf.shape != (H, W):
        raise L.HgsError("hair_image_loss: targets must be gt_rgb[3,H,W], gt_mask/gt_theta/confidence[H,W]")
    om = None
    if orient_mask is not None:
        if orient_mask.shape != (H, W) or orient_mask.device != dev:
            raise L.HgsError("hair_image_loss: orient_mask must be a [H,W] tensor on the image's device")
        if orient_mask.dtype == torch.bool:
            om = orient_mask.contiguous().view(torch.uint8)
        else:
            om = (orient_mask != 0).view(torch.uint8)
    a = L.HairLoss()
    a.height, a.width = H, W
    a.image7, a.gt_rgb, a.gt_mask, a.gt_theta, a.confidence = (img.data_ptr(), gt.data_ptr(), gm.data_ptr(),
                                                             th.data_ptr(), cf.data_ptr())
    a.orient_mask = om.data_ptr() if om is not None else None
    vm = None
    if torch.is_tensor(view_rot):
        # the camera's world_view_transform on the device: read by the kernel, nothing baked into the launch
        if not view_rot.is_cuda or view_rot.device != dev or view_rot.dtype != torch.float32 or view_rot.numel() != 16:
            raise L.HgsError("hair_image_loss: a tensor view_rot must be the [4,4] float32 world_view_transform on the "
                             "image's device")
        vm = view_rot.contiguous()
        a.view_matrix_dev = vm.data_ptr()
    else:
        a.view_matrix_dev = None
        for i, v in enumerate(view_rot):
            a.view_rot[i] = float(v)
    for i, v in enumerate(bg_orient):
        a.bg_orient[i] = float(v)
    a.l_l1, a.l_dssim, a.l_mask, a.l_orient = [float(v) for v in lambdas]
    terms = torch.empty(8, dtype=torch.float32, device=dev)
    scratch = torch.empty(9 * H * W, dtype=torch.float32, device=dev)
    grad = torch.empty_like(img)
    a.terms, a.scratch, a.dL_dimage = terms.data_ptr(), scratch.data_ptr(), grad.data_ptr()
    with torch.cuda.device(dev):
        L.check(lib.hgs_hair_image_loss(ctypes.byref(a), L.stream_ptr(dev)), "hair_image_loss")
    return terms, grad


class _HairImageLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image7, gt_rgb, gt_mask, gt_theta, confidence, orient_mask, view_rot, lambdas, bg_orient):
        terms, grad = hair_image_loss_raw(image7, gt_rgb, gt_mask, gt_theta, confidence, orient_mask, view_rot, lambdas,
                                          bg_orient)
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(terms)
        return terms[0], terms

    @staticmethod
    def backward(ctx, grad_total, _grad_terms):
        (grad,) = ctx.saved_tensors
        return (grad * grad_total,) + (None,) * 8


def view_rot_of(world_view_transform):
    """The nine floats of camera.world_view_transform[:3,:3] (host list; read once per camera, not per step)."""
    return [float(v) for v in world_view_transform[:3, :3].reshape(-1).tolist()]


def hair_image_loss(image7, gt_rgb, gt_mask, gt_theta, confidence, view_rot, lambda_dssim=0.2, lambda_mask=0.01,
                    lambda_orientation=100.0, orient_mask=None, bg_orient=(0.0, 0.0, 0.0)):
    """loss, terms = the image-space part of Hair-GS's loss_function on the fused strand render.

    image7: [7,H,W] from rasterize_strands (rgb | mask logit | world orientation).  view_rot: view_rot_of(camera
    .world_view_transform) — nine host floats — or camera.world_view_transform itself as a CUDA [4,4] tensor (the
    kernel then reads the rotation from device memory: required inside a captured CUDA graph, hairgs_b200.graphs).
    terms (no grad): [total, l1, dssim, mask, orientation, n_orientation_pixels, -, -].
    The lambda defaults are the reference's OptimizationParams (arguments/__init__.py:84-86).  Two documented deviations
    (INTEGRATION.md): an empty orientation mask gives an orientation term of 0 (the reference's mean over nothing is NaN),
    and the orientation channel uses the reference's own `norm >= 1e-7` collapsed-segment rule of get_orientation.
    """
    lambdas = (max(0.0, 1.0 - lambda_dssim), lambda_dssim, lambda_mask, lambda_orientation)
    return _HairImageLoss.apply(image7, gt_rgb, gt_mask, gt_theta, confidence, orient_mask, view_rot, lambdas, bg_orient)


def hair_image_loss_torch(rgb, mask_plane, orientation, gt_rgb, gt_mask, gt_theta, confidence, world_view_transform,
                          lambda_dssim=0.2, lambda_mask=0.01, lambda_orientation=100.0, orient_mask=None,
                          bg_orient=(0.0, 0.0, 0.0)):
    """The same loss written with torch ops the way Hair-GS composes it (loss/losses.py:16-17, 24-84, 87-103, 244-288,
    311-316, 336-346).  It is what the three-render reference arm evaluates in bench.py and what the tests compare the
    fused kernel with; the product path does not call it."""
    import math
    import torch.nn.functional as F
    g = torch.tensor([math.exp(-((x - 5) ** 2) / float(2 * 1.5 ** 2)) for x in range(11)])
    g = (g / g.sum()).unsqueeze(1)
    window = g.mm(g.t()).float()[None, None].expand(3, 1, 11, 11).contiguous().to(rgb)

    def conv(t):
        return F.conv2d(t, window, padding=5, groups=3)
    mu1, mu2 = conv(rgb), conv(gt_rgb)
    mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = conv(rgb * rgb) - mu1_sq
    s2 = conv(gt_rgb * gt_rgb) - mu2_sq
    s12 = conv(rgb * gt_rgb) - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim = (((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))).mean()
    terms = {"l1": torch.abs(rgb - gt_rgb).mean(), "dssim": 1.0 - ssim}
    loss = max(0.0, 1.0 - lambda_dssim) * terms["l1"] + lambda_dssim * terms["dssim"]
    terms["mask"] = F.binary_cross_entropy_with_logits(mask_plane, gt_mask)
    loss = loss + lambda_mask * terms["mask"]
    ow = orientation.permute(1, 2, 0)
    v = ow.flatten(0, 1) @ world_view_transform[:3, :3]
    p = v[:, :2]
    p = p / (torch.norm(p, dim=1, keepdim=True) + 1e-7)
    x, y = p[:, 0], p[:, 1]
    y = torch.where(y < 1e-7, y + 1e-7, y)
    th = torch.atan2(x, y)
    th = torch.where(th < 0, th + math.pi, th).reshape(ow.shape[:2])
    m = orient_mask if orient_mask is not None else torch.any(ow != torch.tensor(bg_orient).to(ow), dim=2)
    d = math.pi / 2 - torch.abs(torch.abs(th[m] - gt_theta[m]) - math.pi / 2)
    terms["orientation"] = (d * confidence[m]).mean()
    loss = loss + lambda_orientation * terms["orientation"]
    return loss, terms
