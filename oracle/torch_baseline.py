"""TEST / MEASUREMENT INFRASTRUCTURE — the torch-CPU baseline of SURVEY.md §8(d), never imported by the product.

The reference rasterizer is CUDA-only; what the reference CAN run on a CPU is its torch side.  BASELINE.json's north_star
therefore defines the CPU baseline as "its torch-side preprocessing (utils/sh.py, covariance build) plus a naive torch
compositor".  This module is that baseline, item by item as §8(d) lists them, in plain torch fp32 on the host cores:

  (i)   eval_sh at the active degree + clamp        utils/sh.py:55-118, gaussian_renderer/__init__.py:93-102
  (ii)  covariance build L·Lᵀ -> 6 floats            utils/transform.py:7-42, utils/general.py:71-84
  (iii) strand parameterisation (getters)            scene/hair_gaussian_model.py:134-206
  (iv)  EWA projection, radius, tile rectangle       DGR/cuda_rasterizer/forward.cu:74-113, 216-255, auxiliary.h:46-58
  (v)   naive compositor: per tile, depth-sorted list, dense [256, K] alpha, cumprod transmittance with the
        0.99 / 1/255 / 1e-4 rules of forward.cu:336-351, blend; backward through autograd

(vi) of §8(d) (`compute_metrics`, loss/metrics.py) needs the reference tree at run time and is not part of the render
path; it is not restated.  The compositor is checked against the C oracle in tests/test_torch_baseline.py (pixels and
autograd gradients), so the baseline times a computation with the reference's semantics, not a look-alike.
"""
import os
import sys
import time

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(os.path.dirname(_HERE), "hair-gs_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
from hairgs_b200 import sh as _sh  # noqa: E402  (torch helpers restating utils/sh.py, utils/transform.py)

TILE = 16


def colours_from_sh(degree, shs, means3D, campos):
    """(i) gaussian_renderer/__init__.py:93-102: directions from the camera centre, eval_sh, +0.5, clamp at 0."""
    dirs = means3D - campos.reshape(1, 3)
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    rgb = _sh.eval_sh(degree, shs.transpose(1, 2), dirs)
    return torch.clamp_min(rgb + 0.5, 0.0)


def covariance6(scales, scale_modifier, rotations):
    """(ii) scene/gaussian_model.py:61-65 -> [P,6] (xx, xy, xz, yy, yz, zz)."""
    return _sh.build_covariance_from_scaling_rotation(scales, scale_modifier, rotations)


def project_ewa(means3D, cov6, viewmatrix, projmatrix, tan_fovx, tan_fovy, W, H):
    """(iv) forward.cu:182-255: frustum cull (view z <= 0.2), EWA covariance + 0.3 low-pass, conic, 3-sigma radius, pixel
    centre, tile rectangle.  Matrices are the reference's tensors (row vectors: p_view = [p,1] @ viewmatrix)."""
    P = means3D.shape[0]
    ones = torch.ones(P, 1, dtype=means3D.dtype)
    ph = torch.cat([means3D, ones], 1)
    pv = ph @ viewmatrix
    hom = ph @ projmatrix
    p_w = 1.0 / (hom[:, 3] + 0.0000001)
    proj = hom[:, :2] * p_w[:, None]
    depth = pv[:, 2]
    in_front = depth > 0.2
    fx, fy = W / (2.0 * tan_fovx), H / (2.0 * tan_fovy)
    tz = torch.where(in_front, depth, torch.ones_like(depth))
    limx, limy = 1.3 * tan_fovx, 1.3 * tan_fovy
    tx = torch.clamp(pv[:, 0] / tz, -limx, limx) * tz
    ty = torch.clamp(pv[:, 1] / tz, -limy, limy) * tz
    zero = torch.zeros_like(tz)
    J = torch.stack([torch.stack([fx / tz, zero, -(fx * tx) / (tz * tz)], 1),
                     torch.stack([zero, fy / tz, -(fy * ty) / (tz * tz)], 1)], 1)          # [P,2,3]
    Wm = viewmatrix[:3, :3].t()                                                            # world -> view rotation
    Sigma = torch.stack([torch.stack([cov6[:, 0], cov6[:, 1], cov6[:, 2]], 1),
                         torch.stack([cov6[:, 1], cov6[:, 3], cov6[:, 4]], 1),
                         torch.stack([cov6[:, 2], cov6[:, 4], cov6[:, 5]], 1)], 1)         # [P,3,3]
    JW = J @ Wm
    cov2 = JW @ Sigma @ JW.transpose(1, 2)
    a, b, c = cov2[:, 0, 0] + 0.3, cov2[:, 0, 1], cov2[:, 1, 1] + 0.3
    det = a * c - b * b
    ok = in_front & (det != 0)
    det_s = torch.where(ok, det, torch.ones_like(det))
    conic = torch.stack([c / det_s, -b / det_s, a / det_s], 1)
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam)).detach()
    px = ((proj[:, 0] + 1.0) * W - 1.0) * 0.5
    py = ((proj[:, 1] + 1.0) * H - 1.0) * 0.5
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    with torch.no_grad():
        def cl(v, hi):
            return torch.clamp(v.to(torch.int64), 0, hi)       # int() truncation as the reference's cast
        rmin = torch.stack([cl((px - radius) / TILE, gx), cl((py - radius) / TILE, gy)], 1)
        rmax = torch.stack([cl((px + radius + TILE - 1) / TILE, gx), cl((py + radius + TILE - 1) / TILE, gy)], 1)
        touched = (rmax[:, 0] - rmin[:, 0]) * (rmax[:, 1] - rmin[:, 1])
        touched = torch.where(ok, touched, torch.zeros_like(touched))
    return dict(depth=depth, means2D=torch.stack([px, py], 1), conic=conic, radii=torch.where(touched > 0, radius, 0 * radius),
                rmin=rmin, rmax=rmax, touched=touched, grid=(gx, gy))


def tile_lists(pr):
    """duplicateWithKeys + stable (tile | depth) sort + identifyTileRanges (rasterizer_impl.cu:70-138) with torch ops:
    -> (point_list [N], ranges [T,2])."""
    gx, gy = pr["grid"]
    touched = pr["touched"]
    ids = torch.repeat_interleave(torch.arange(touched.shape[0]), touched)
    first = torch.cumsum(touched, 0) - touched
    k = torch.arange(ids.shape[0]) - first[ids]
    w = (pr["rmax"][:, 0] - pr["rmin"][:, 0])[ids]
    ty = pr["rmin"][ids, 1] + k // torch.clamp_min(w, 1)
    tx = pr["rmin"][ids, 0] + k % torch.clamp_min(w, 1)
    tile = ty * gx + tx
    depth_bits = pr["depth"].detach().contiguous().view(torch.int32)[ids].to(torch.int64) & 0xffffffff
    keys = (tile << 32) | depth_bits
    keys_sorted, order = torch.sort(keys, stable=True)
    point_list = ids[order]
    tiles_sorted = keys_sorted >> 32
    T = gx * gy
    counts = torch.bincount(tiles_sorted, minlength=T)
    ends = torch.cumsum(counts, 0)
    ranges = torch.stack([ends - counts, ends], 1)
    return point_list, ranges


def composite_tile(x0, y0, W, H, idx, means2D, conic, opacity, colors, bg):
    """(v) one 16x16 tile against its depth-sorted list `idx` [K]: dense [256, K] evaluation of forward.cu:323-357.
    -> (pixels [C, 16, 16], final_T [16, 16]).  Differentiable; rows of pixels outside the image are still computed
    (the caller crops)."""
    ys, xs = torch.meshgrid(torch.arange(y0, y0 + TILE, dtype=torch.float32), torch.arange(x0, x0 + TILE, dtype=torch.float32),
                            indexing="ij")
    px, py = xs.reshape(-1, 1), ys.reshape(-1, 1)                              # [256,1]
    m, cn = means2D[idx], conic[idx]
    dx, dy = m[:, 0][None, :] - px, m[:, 1][None, :] - py                      # [256,K]
    power = -0.5 * (cn[:, 0][None, :] * dx * dx + cn[:, 2][None, :] * dy * dy) - cn[:, 1][None, :] * dx * dy
    alpha = torch.clamp_max(opacity[idx].reshape(1, -1) * torch.exp(power), 0.99)
    skip = (power > 0) | (alpha < 1.0 / 255.0)
    alpha = torch.where(skip, torch.zeros_like(alpha), alpha)
    T_incl = torch.cumprod(1.0 - alpha, dim=1)                                 # transmittance AFTER each list entry
    T_excl = torch.cat([torch.ones_like(T_incl[:, :1]), T_incl[:, :-1]], 1)
    live = T_incl >= 0.0001                # forward.cu:346-351: the entry that would push T below 1e-4 ends the pixel
    w = torch.where(live, alpha * T_excl, torch.zeros_like(alpha))
    pix = w @ colors[idx]                                                      # [256,C]
    # transmittance left when the pixel stopped: after its last live entry
    final_T = torch.where(live, T_incl, torch.ones_like(T_incl)).min(dim=1).values
    pix = pix + final_T[:, None] * bg.reshape(1, -1)
    C = colors.shape[1]
    return pix.t().reshape(C, TILE, TILE), final_T.reshape(TILE, TILE)


def render(d, tiles=None):
    """The whole view through (i)/(ii)/(iv)/(v).  d: the drop-in's argument dict (tests/common.py) with CPU tensors.
    tiles: iterable of tile ids to composite (default: all).  -> (image [C,H,W], radii [P], info dict)."""
    W, H = int(d["image_width"]), int(d["image_height"])
    means3D, opacity = d["means3D"], d["opacity"]
    t = {}
    t0 = time.perf_counter()
    if d["colors"].numel():
        colors = d["colors"]
    else:
        colors = colours_from_sh(int(d["degree"]), d["sh"], means3D, d["campos"])
    t["sh"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    cov6 = d["cov3D_precomp"] if d["cov3D_precomp"].numel() else covariance6(d["scales"], float(d["scale_modifier"]),
                                                                              d["rotations"])
    t["covariance"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    pr = project_ewa(means3D, cov6, d["viewmatrix"], d["projmatrix"], float(d["tan_fovx"]), float(d["tan_fovy"]), W, H)
    t["projection"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    point_list, ranges = tile_lists(pr)
    t["binning"] = time.perf_counter() - t0
    gx, gy = pr["grid"]
    C = colors.shape[1]
    bg = d["background"]
    image = bg.reshape(C, 1, 1).expand(C, gy * TILE, gx * TILE).clone()
    todo = range(gx * gy) if tiles is None else tiles
    t0 = time.perf_counter()
    entries = 0
    for tl in todo:
        r0, r1 = int(ranges[tl, 0]), int(ranges[tl, 1])
        if r1 <= r0:
            continue
        x0, y0 = (tl % gx) * TILE, (tl // gx) * TILE
        pix, _ = composite_tile(x0, y0, W, H, point_list[r0:r1], pr["means2D"], pr["conic"], opacity, colors, bg)
        image[:, y0:y0 + TILE, x0:x0 + TILE] = pix
        entries += r1 - r0
    t["composite"] = time.perf_counter() - t0
    info = dict(times=t, num_rendered=int(point_list.shape[0]), entries_composited=entries, ranges=ranges)
    return image[:, :H, :W], pr["radii"].to(torch.int32), info


def time_view(d, strand=None, max_tiles=96, threads=None, seed=0):
    """Bounded timing of one training view (forward + autograd backward) for bench.py's cpu_baseline: items (i)-(iv) on
    all Gaussians, item (v) on `max_tiles` randomly chosen non-empty tiles, extrapolated to the view by tile-list entries.
    strand = (endpoints, endpoint_pairs, width) adds item (iii) (hairgs_b200.scenes.strand_gaussians).
    -> dict(ms_per_view, items={...ms}, cores, sample)."""
    if threads:
        torch.set_num_threads(int(threads))
    cores = torch.get_num_threads()
    items = {}
    base = dict(d)

    def inputs():
        """Fresh leaves (and, for strands, a fresh parameterisation graph) for one forward + backward pass."""
        x = dict(base)
        if strand is not None:
            from hairgs_b200 import scenes
            ep = strand[0].clone().requires_grad_(True)
            wd = strand[2].clone().requires_grad_(True)
            t0 = time.perf_counter()
            means, scales, rot, _ = scenes.strand_gaussians(ep, strand[1], wd)
            items["strand_parameterisation"] = time.perf_counter() - t0
            x["means3D"], x["scales"], x["rotations"] = means, scales, rot
        else:
            for k in ("means3D", "scales", "rotations"):
                x[k] = base[k].clone().requires_grad_(True)
        x["opacity"] = base["opacity"].clone().requires_grad_(True)
        if base["sh"].numel():
            x["sh"] = base["sh"].clone().requires_grad_(True)
        return x

    # pass 0 (no tiles, no grad): preprocessing + binning, to learn the tile lists
    with torch.no_grad():
        _, _, info = render(inputs(), tiles=[])
    ranges = info["ranges"]
    nonempty = torch.nonzero(ranges[:, 1] > ranges[:, 0]).flatten()
    g = torch.Generator().manual_seed(seed)
    pick = nonempty[torch.randperm(nonempty.numel(), generator=g)[:max_tiles]].tolist()
    total_entries = int((ranges[:, 1] - ranges[:, 0]).sum())

    def fwd_bwd(tiles):
        x = inputs()
        t0 = time.perf_counter()
        image, _, inf = render(x, tiles=tiles)
        t1 = time.perf_counter()
        dL = torch.randn(image.shape, generator=g)
        t2 = time.perf_counter()
        (image * dL).sum().backward()
        return inf, t1 - t0, time.perf_counter() - t2

    # two passes: the sample of tiles, and a single tile — their difference separates the compositor's cost per list
    # entry from the preprocessing (which every pass pays in full, forward and backward)
    info, fwd_a, bwd_a = fwd_bwd(pick)
    info_b, fwd_b, bwd_b = fwd_bwd(pick[:1])
    de = max(1, info["entries_composited"] - info_b["entries_composited"])
    comp_fwd_per_entry = max(0.0, info["times"]["composite"] - info_b["times"]["composite"]) / de
    comp_bwd_per_entry = max(0.0, bwd_a - bwd_b) / de
    pre_fwd = fwd_b - info_b["times"]["composite"]
    pre_bwd = max(0.0, bwd_b - comp_bwd_per_entry * info_b["entries_composited"])
    for k, v in info_b["times"].items():
        if k != "composite":
            items[k] = v
    items["preprocess_backward"] = pre_bwd
    items["composite_forward_extrapolated"] = comp_fwd_per_entry * total_entries
    items["composite_backward_extrapolated"] = comp_bwd_per_entry * total_entries
    ms = 1e3 * (items.get("strand_parameterisation", 0.0) + pre_fwd + pre_bwd + items["composite_forward_extrapolated"] +
                items["composite_backward_extrapolated"])
    return dict(ms_per_view=ms, items={k: round(1e3 * v, 2) for k, v in items.items()}, cores=cores,
                sample=f"(i)-(iv) on all {base['opacity'].shape[0]} Gaussians, forward and autograd backward; naive "
                       f"compositor on {len(pick)} of {nonempty.numel()} non-empty tiles ({info['entries_composited']} of "
                       f"{total_entries} list entries), forward and backward, extrapolated by entries; torch "
                       f"{torch.__version__} fp32, {cores} threads")
