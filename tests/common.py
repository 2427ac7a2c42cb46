"""Shared helpers for the parity tests: seeded inputs in the `_C.rasterize_gaussians` argument order, and
runners that expose the intermediate buffers of our library and of the reference build side by side."""
import math

import torch

from hairgs_b200 import _lib as L
from hairgs_b200 import scenes
import diff_gaussian_rasterization._C as ours_C

import refload

EMPTY = lambda: torch.Tensor([])  # noqa: E731  (what GaussianRasterizer.forward substitutes for absent inputs)


def blob_inputs(P, W, H, device, seed=0, sh_degree=3, M=16, view=0, n_views=4, colors=None, scale_mul=1.0,
                use_cov=False):
    sc = scenes.blob_scene(P, seed=seed, sh_coeffs=M).to(device)
    cam = scenes.orbit_cameras(n_views, W, H, device=device)[view]
    d = dict(background=torch.tensor([0.1, 0.2, 0.3], device=device), means3D=sc.means3D, colors=EMPTY(),
             opacity=sc.opacities, scales=sc.scales * scale_mul, rotations=sc.rotations, scale_modifier=1.0,
             cov3D_precomp=EMPTY(), viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
             tan_fovx=cam.tanfovx, tan_fovy=cam.tanfovy, image_height=H, image_width=W, sh=sc.shs, degree=sh_degree,
             campos=cam.camera_center, prefiltered=False, debug=False)
    if colors is not None:
        d["colors"], d["sh"] = colors, EMPTY()
    return d


def strand_inputs(S, V, W, H, device, seed=0, view=0, n_views=4, colors=None):
    sc = scenes.strand_scene(S, V, seed=seed).to(device)
    cam = scenes.orbit_cameras(n_views, W, H, device=device)[view]
    means, scales, rot, orient = scenes.strand_gaussians(sc.endpoints, sc.endpoint_pairs, sc.width)
    d = dict(background=torch.zeros(3, device=device), means3D=means.contiguous(), colors=EMPTY(),
             opacity=torch.sigmoid(sc.opacity_logit), scales=scales.contiguous(), rotations=rot.contiguous(),
             scale_modifier=1.0, cov3D_precomp=EMPTY(), viewmatrix=cam.world_view_transform,
             projmatrix=cam.full_proj_transform, tan_fovx=cam.tanfovx, tan_fovy=cam.tanfovy, image_height=H,
             image_width=W, sh=torch.cat((sc.features_dc, sc.features_rest), 1).contiguous(), degree=0,
             campos=cam.camera_center, prefiltered=False, debug=False)
    if colors == "orientation":
        d["colors"], d["sh"] = orient.contiguous(), EMPTY()
    elif colors == "mask":
        d["colors"], d["sh"] = torch.sigmoid(sc.mask_logit).repeat(1, 3).contiguous(), EMPTY()
    return d


FWD_ORDER = ("background", "means3D", "colors", "opacity", "scales", "rotations", "scale_modifier", "cov3D_precomp",
             "viewmatrix", "projmatrix", "tan_fovx", "tan_fovy", "image_height", "image_width", "sh", "degree",
             "campos", "prefiltered", "debug")


def fwd_args(d):
    return tuple(d[k] for k in FWD_ORDER)


def bwd_args(d, radii, dL, geom, R, binning, img):
    return (d["background"], d["means3D"], radii, d["colors"], d["scales"], d["rotations"], d["scale_modifier"],
            d["cov3D_precomp"], d["viewmatrix"], d["projmatrix"], d["tan_fovx"], d["tan_fovy"], dL, d["sh"],
            d["degree"], d["campos"], geom, R, binning, img, d["debug"])


def ours_forward(d):
    """Returns (N, color, radii, buffers, views) with every intermediate in reference layout."""
    N, color, radii, geom, binning, img = ours_C.rasterize_gaussians(*fwd_args(d))
    dev, prm, inp, keep = ours_C._prep(*fwd_args(d))
    views = {}
    for name, what in (("depths", L.VIEW_DEPTHS), ("means2D", L.VIEW_MEANS2D), ("conic_opacity", L.VIEW_CONIC_OPACITY),
                       ("rgb", L.VIEW_RGB), ("tiles_touched", L.VIEW_TILES_TOUCHED),
                       ("point_offsets", L.VIEW_POINT_OFFSETS), ("clamped", L.VIEW_CLAMPED),
                       ("point_list_keys", L.VIEW_KEYS_SORTED), ("point_list", L.VIEW_POINT_LIST),
                       ("ranges", L.VIEW_RANGES), ("accum_alpha", L.VIEW_FINAL_T), ("n_contrib", L.VIEW_N_CONTRIB)):
        views[name] = L.state_view(what, prm, inp, N, geom, binning, img)
    if d["scales"].numel():
        views["cov3D"] = L.state_view(L.VIEW_COV3D, prm, inp, N, geom, binning, img)
    return N, color, radii, (geom, binning, img), views


def ref_forward(d):
    C = refload.ref_dgr()
    N, color, radii, geom, binning, img = C.rasterize_gaussians(*fwd_args(d))
    P = d["means3D"].shape[0]
    H, W = d["image_height"], d["image_width"]
    views = {}
    g = refload.unpack_geom(geom, P)
    vis = radii > 0
    for k in ("depths", "means2D", "conic_opacity", "rgb", "cov3D", "clamped"):
        t = g[k].clone()
        t[~vis] = 0  # the reference leaves culled entries uninitialised (torch.empty)
        views[k] = t
    views["tiles_touched"] = g["tiles_touched"].clone()
    views["point_offsets"] = g["point_offsets"].clone()
    b = refload.unpack_binning(binning, N)
    views["point_list_keys"] = b["point_list_keys"].clone()
    views["point_list"] = b["point_list"].clone()
    im = refload.unpack_image(img, H, W)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    views["ranges"] = im["ranges"][:T].clone()
    views["accum_alpha"] = im["accum_alpha"].view(H, W).clone()
    views["n_contrib"] = im["n_contrib"].view(H, W).clone()
    return N, color, radii, (geom, binning, img), views


def bits_equal(a, b):
    """Bit-for-bit equality count for float tensors (compares the raw words, so -0 != +0, NaN == same NaN)."""
    if a.dtype == torch.float32:
        a, b = a.contiguous().view(torch.int32), b.contiguous().view(torch.int32)
    return int((a != b).sum().item())


def rel_err(a, b):
    """max|a-b| / max|b| — the gradient tolerance metric (rel <= 1e-3)."""
    den = b.abs().max().item()
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)
