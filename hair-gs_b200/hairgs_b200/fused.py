"""Fused strand render path (SURVEY.md §8f rows N1 + N2): what Hair-GS computes per training view with THREE render()
calls — RGB via SH, mask via override_color=get_mask.repeat(1,3), orientation via override_color=get_orientation
(train.py:146-155, loss/losses.py:246-249,311-312,341-346) — plus the ~25 torch kernels of the strand getters
(scene/hair_gaussian_model.py:134-206) per call, done here in ONE rasterization pass:

  * the preprocess kernel derives the Gaussians from (endpoints, endpoint_pairs, width) and applies the sigmoid /
    exp activations itself (hgs_strands_forward_stage_a);
  * geometry (cull, scan, keys, sort, ranges) runs once and seven channels are composited together;
  * one backward pass returns gradients w.r.t. the RAW parameters (end points scattered to their joints, width,
    opacity / mask logits, SH features) and the screen-space mean gradients Hair-GS uses for densification statistics.

Equivalence with the three-pass drop-in path is tested to the north_star tolerances
(tests/test_gpu_parity.py::test_fused_strands_equals_three_pass_dropin).
"""
import ctypes
import math

import torch

from . import _lib as L
from diff_gaussian_rasterization import _C as _dgr  # capacity hints / pinned read-back ring are shared


# HGS_BWD_RED=scalar: the round-1 accumulation (four arrays, scalar red.global.add.f32); default: interleaved records +
# red.global.add.v4.f32 (A/B: profiles/r2_bwd_vector_red.md)
BWD_VECTOR_RED = __import__("os").environ.get("HGS_BWD_RED", "vec") != "scalar"


def _prep(endpoints, endpoint_pairs, width, opacity_logit, mask_logit, features, s):
    if not endpoints.is_cuda:
        raise L.HgsError("endpoints must be a CUDA tensor: this rasterizer has no CPU path")
    dev = endpoints.device
    if endpoint_pairs.dtype != torch.int64 or not endpoint_pairs.is_cuda:
        raise L.HgsError("endpoint_pairs must be a CUDA int64 tensor [P,2]")
    keep = dict(endpoints=L.f32c(endpoints, "endpoints", dev), width=L.f32c(width, "width", dev),
                opacity_logit=L.f32c(opacity_logit, "opacity_logit", dev), mask_logit=L.f32c(mask_logit, "mask_logit", dev),
                features=L.f32c(features, "features", dev), background=L.f32c(s["bg"], "bg", dev),
                viewmatrix=L.f32c(s["viewmatrix"], "viewmatrix", dev), projmatrix=L.f32c(s["projmatrix"], "projmatrix", dev),
                cam_pos=L.f32c(s["campos"], "campos", dev), endpoint_pairs=endpoint_pairs.contiguous())
    if keep["background"] is None or keep["background"].numel() != 7:
        raise L.HgsError("the fused strand pass needs a 7-channel background (rgb, mask, orientation)")
    P = endpoint_pairs.shape[0]
    M = features.shape[1] if features.numel() else 0
    prm = L.RasterParams(P=P, D=int(s["sh_degree"]), M=int(M), width=int(s["image_width"]), height=int(s["image_height"]),
                         channels=7, tan_fovx=float(s["tanfovx"]), tan_fovy=float(s["tanfovy"]),
                         scale_modifier=float(s["scale_modifier"]), prefiltered=0, debug=int(bool(s["debug"])))
    inp = L.StrandInputs(num_endpoints=int(endpoints.shape[0]), **{k: L.ptr(v) for k, v in keep.items()})
    return dev, prm, inp, keep


def strands_forward(endpoints, endpoint_pairs, width, opacity_logit, mask_logit, features, settings):
    """The forward pass without autograd: -> (image[7,H,W], radii[P], state) where state = (geom, binning, img, capacity,
    num_rendered) is what strands_backward needs.  num_rendered is -1 under a launch plan (not read back)."""
    lib = L.load()
    dev, prm, inp, keep = _prep(endpoints, endpoint_pairs, width, opacity_logit, mask_logit, features, settings)
    P, H, W = prm.P, prm.height, prm.width
    u8 = dict(dtype=torch.uint8, device=dev)
    plan = settings.get("plan")
    if plan is not None:
        prm.slice_base, prm.slice_shift, prm.sort_mode = plan.slice_base, plan.slice_shift, int(plan.sort_mode)
    else:
        prm.slice_base, prm.slice_shift = _dgr.slice_params((dev.index, P, H, W, 7))
        prm.sort_mode = _dgr.sort_mode_for((dev.index, P, H, W, 7))
    with torch.cuda.device(dev):
        stream = L.stream_ptr(dev)
        image = torch.empty((7, H, W), dtype=torch.float32, device=dev)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        geom = torch.empty((lib.hgs_geom_bytes(P, 7, W, H),), **u8)
        img = torch.empty((lib.hgs_image_bytes(W, H),), **u8)
        L.check(lib.hgs_strands_forward_stage_a(ctypes.byref(prm), ctypes.byref(inp), geom.data_ptr(),
                                                radii.data_ptr(), stream), "strands stage A")
        if plan is not None:
            # fixed launch plan (hairgs_b200.graphs): no host synchronisation and no host-side decision depends on
            # this view's instance count, so the whole pass can be captured in a CUDA graph.  The count, the
            # overflow flags and the depth range land in plan.host (pinned) when the work has run; plan.check()
            # validates them afterwards.
            L.check(lib.hgs_forward_read_num_rendered(geom.data_ptr(), P, plan.host.data_ptr(), stream),
                    "read num_rendered")
            prm.sort_depth_bits = int(plan.depth_bits)
            prm.sort_mode = int(plan.sort_mode)
            cap = int(plan.capacity)
            binning = torch.empty((lib.hgs_binning_bytes(cap, 7),), **u8)
            L.check(lib.hgs_strands_forward_stage_b(ctypes.byref(prm), ctypes.byref(inp), geom.data_ptr(),
                                                    binning.data_ptr(), img.data_ptr(), cap, image.data_ptr(), stream),
                    "strands stage B")
            return image, radii, (geom, binning, img, cap, -1)
        host = _dgr._pinned_triplet()
        L.check(lib.hgs_forward_read_num_rendered(geom.data_ptr(), P, host.data_ptr(), stream), "read num_rendered")
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        key = (dev.index, P, H, W, 7)
        cap = _dgr._capacity_hint.get(key) if _dgr.SYNC_FREE else None
        prm.sort_depth_bits = _dgr._depth_bits_hint.get(key, 0) if _dgr.SYNC_FREE else 0
        binning = None
        if cap is not None:
            binning = torch.empty((lib.hgs_binning_bytes(cap, 7),), **u8)
            L.check(lib.hgs_strands_forward_stage_b(ctypes.byref(prm), ctypes.byref(inp), geom.data_ptr(),
                                                    binning.data_ptr(), img.data_ptr(), cap, image.data_ptr(), stream),
                    "strands stage B")
        ready.synchronize()
        N, overflow = int(host[0]), int(host[2])
        if (overflow & 1) != 0 or N < 0:
            raise L.HgsError("instance count overflows int32")
        need = _dgr._depth_range_bits(host)
        _dgr.note_depth_range(key, host)
        if (overflow & 4) and prm.slice_shift > 0:
            _dgr._sort_mode_hint[key] = L.SORT_GLOBAL   # see diff_gaussian_rasterization._C.rasterize_gaussians
        if prm.sort_mode == L.SORT_TILE:
            fits = (overflow & 4) == 0
        else:
            fits = prm.sort_depth_bits in (0, 32) or need <= prm.sort_depth_bits
        if cap is None or N > cap or not fits:
            if cap is None or N > cap:
                cap = N
                binning = torch.empty((lib.hgs_binning_bytes(N, 7),), **u8)
            if overflow & 4:
                prm.sort_mode = L.SORT_GLOBAL
            prm.sort_depth_bits = _dgr._next_depth_bits(H, W, need) if _dgr.SYNC_FREE else 0
            L.check(lib.hgs_strands_forward_stage_b(ctypes.byref(prm), ctypes.byref(inp), geom.data_ptr(),
                                                    binning.data_ptr() if cap > 0 else None, img.data_ptr(), cap,
                                                    image.data_ptr(), stream), "strands stage B")
        _dgr._capacity_hint[key] = _dgr._next_capacity(_dgr._capacity_hint.get(key), N)
        _dgr._depth_bits_hint[key] = max(_dgr._depth_bits_hint.get(key, 0), _dgr._next_depth_bits(H, W, need))
    return image, radii, (geom, binning, img, cap, N)


def strands_backward(endpoints, endpoint_pairs, width, opacity_logit, mask_logit, features, settings, state, grad_image):
    """The backward pass without autograd.  -> (dL_dendpoints, dL_dwidth, dL_dopacity_logit, dL_dmask_logit, dL_dfeatures,
    dL_dmean2D); with settings["grad_sink"] the first five are the sink's tensors (written or accumulated in place)."""
    lib = L.load()
    geom, binning, img, capacity, _ = state
    dev, prm, inp, keep = _prep(endpoints, endpoint_pairs, width, opacity_logit, mask_logit, features, settings)
    P, M, E = prm.P, prm.M, endpoints.shape[0]
    f32 = dict(dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        dpix = L.f32c(grad_image, "grad_image", dev)
        vec = BWD_VECTOR_RED
        if vec:
            # one interleaved 64-byte accumulation record per Gaussian (hgs_strand_grads.acc16: four red.global.add.v4.f32
            # per instance and pixel block instead of 13 scalar atomics) + the [P,3] screen-space mean gradients written
            # from it by the preprocess backward
            acc = torch.empty((P * 19,), **f32)
            d_mean2D = acc[16 * P:].view(P, 3)
        else:
            acc = torch.empty((P * 15,), **f32)  # mean2D 3 | conic 4 | opacity 1 | colour 7 : one memset in the library
            d_mean2D = acc[:3 * P].view(P, 3)
        sink = settings.get("grad_sink")
        if sink is None:
            out = torch.empty((3 * E + 3 * P + 3 * M * P,), **f32)
            d_end = out[:3 * E].view(E, 3)
            d_width = out[3 * E:3 * E + P].view(P, 1)
            d_opac = out[3 * E + P:3 * E + 2 * P].view(P, 1)
            d_mask = out[3 * E + 2 * P:3 * E + 3 * P].view(P, 1)
            d_feat = out[3 * E + 3 * P:].view(P, M, 3)
            accumulate = 0
        else:
            # the kernel deposits the parameter gradients straight into the caller's tensors (slices of a flat
            # gradient bucket): no autograd AccumulateGrad adds, no bucket clear before the first view
            d_end, d_width, d_opac, d_mask, d_feat = (sink.tensors[k] for k in ("endpoints", "width", "opacity", "mask",
                                                                                 "features"))
            for t, n in ((d_end, 3 * E), (d_width, P), (d_opac, P), (d_mask, P), (d_feat, 3 * M * P)):
                if t.numel() != n or t.dtype != torch.float32 or t.device != dev or not t.is_contiguous():
                    raise L.HgsError("grad_sink: tensors must be contiguous float32 on the render device, shaped like "
                                     "the parameters")
            accumulate = 1 if sink.accumulate else 0
            sink.accumulate = True      # later views of the same optimiser step add to the first
        grads = L.StrandGrads(dL_dmean2D=d_mean2D.data_ptr(), dL_dconic=None if vec else acc[3 * P:].data_ptr(),
                              dL_dopacity=None if vec else acc[7 * P:].data_ptr(),
                              dL_dcolor=None if vec else acc[8 * P:].data_ptr(),
                              dL_dendpoints=d_end.data_ptr(), dL_dwidth=d_width.data_ptr(),
                              dL_dopacity_logit=d_opac.data_ptr(), dL_dmask_logit=d_mask.data_ptr(),
                              dL_dfeatures=d_feat.data_ptr(), accumulate=accumulate, acc16=acc.data_ptr() if vec else None)
        L.check(lib.hgs_strands_backward(ctypes.byref(prm), ctypes.byref(inp), int(capacity), geom.data_ptr(),
                                         L.ptr(binning), img.data_ptr(), dpix.data_ptr(), ctypes.byref(grads),
                                         L.stream_ptr(dev)), "strands backward")
    return d_end, d_width.view_as(width), d_opac.view_as(opacity_logit), d_mask.view_as(mask_logit), d_feat, d_mean2D


class _RasterizeStrands(torch.autograd.Function):
    @staticmethod
    def forward(ctx, endpoints, endpoint_pairs, width, opacity_logit, mask_logit, features, means2D, settings):
        image, radii, state = strands_forward(endpoints, endpoint_pairs, width, opacity_logit, mask_logit, features, settings)
        geom, binning, img, cap, N = state
        ctx.settings, ctx.capacity, ctx.num_rendered = settings, cap, N
        ctx.save_for_backward(endpoints, endpoint_pairs, width, opacity_logit, mask_logit, features, geom, binning, img)
        ctx.mark_non_differentiable(radii)
        return image, radii

    @staticmethod
    def backward(ctx, grad_image, _):
        endpoints, endpoint_pairs, width, opacity_logit, mask_logit, features, geom, binning, img = ctx.saved_tensors
        d_end, d_width, d_opac, d_mask, d_feat, d_mean2D = strands_backward(
            endpoints, endpoint_pairs, width, opacity_logit, mask_logit, features, ctx.settings,
            (geom, binning, img, ctx.capacity, ctx.num_rendered), grad_image)
        if ctx.settings.get("grad_sink") is not None:
            return (None, None, None, None, None, None, d_mean2D, None)
        return (d_end, None, d_width, d_opac, d_mask, d_feat, d_mean2D, None)


class GradSink:
    """Where the strand backward deposits the parameter gradients instead of returning them through autograd:
    tensors = {"endpoints": [E,3], "width": [P,1], "opacity": [P,1], "mask": [P,1], "features": [P,M,3]} — typically the
    slices of a multiview.GradBucket / FlatAdam gradient buffer that the parameters' .grad already point at.  The first
    backward after begin_step() overwrites them (no clear needed), later ones add."""

    def __init__(self, tensors):
        self.tensors = tensors
        self.accumulate = False

    def begin_step(self):
        self.accumulate = False


def rasterize_strands(endpoints, endpoint_pairs, width, opacity_logit, mask_logit, features, means2D, *, image_height,
                      image_width, tanfovx, tanfovy, bg, scale_modifier, viewmatrix, projmatrix, sh_degree, campos,
                      debug=False, grad_sink=None, plan=None):
    """-> (image[7,H,W] = rgb | mask | orientation, radii[P]).  plan: a graphs.LaunchPlan (fixed capacity / depth bits, no
    host synchronisation) or None (capacity hints + asynchronous read-back, re-run of stage B when they did not fit)."""
    settings = dict(image_height=image_height, image_width=image_width, tanfovx=tanfovx, tanfovy=tanfovy, bg=bg,
                    scale_modifier=scale_modifier, viewmatrix=viewmatrix, projmatrix=projmatrix, sh_degree=sh_degree,
                    campos=campos, debug=debug, grad_sink=grad_sink, plan=plan)
    return _RasterizeStrands.apply(endpoints, endpoint_pairs, width, opacity_logit, mask_logit, features, means2D,
                                   settings)


def render_strands(viewpoint_camera, pc, bg_color, scaling_modifier=1.0, debug=False, grad_sink=None, plan=None):
    """One-pass counterpart of the three render() calls of a Hair-GS Stage-III iteration.  `pc` is a
    HairGaussianModel-like object (hairgs_b200.models.StrandModel): _endpoints, endpoint_pairs, _width, _opacity, _mask,
    get_features, active_sh_degree.  bg_color: [7].  Returns render / mask / orientation images plus the usual
    viewspace_points / visibility_filter / radii entries of gaussian_renderer.render()."""
    P = pc.endpoint_pairs.shape[0]
    screenspace_points = torch.zeros((P, 3), dtype=pc._endpoints.dtype, requires_grad=True, device=pc._endpoints.device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    image, radii = rasterize_strands(
        pc._endpoints, pc.endpoint_pairs, pc._width, pc._opacity, pc._mask, pc.get_features, screenspace_points,
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5), tanfovy=math.tan(viewpoint_camera.FoVy * 0.5), bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform, sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center, debug=debug, grad_sink=grad_sink, plan=plan)
    return {"render": image[0:3], "mask": image[3:4], "orientation": image[4:7], "image7": image,
            "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii}
