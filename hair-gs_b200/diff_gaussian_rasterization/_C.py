"""`diff_gaussian_rasterization._C` — same three entry points, argument order and return tuples as the
reference's pybind module (submodules/diff-gaussian-rasterization/ext.cpp:15-19, signatures
rasterize_points.h:18-66), implemented over the C ABI of libhairgs_rast.so (sm_100a kernels).

Differences a caller can observe (all documented in DESIGN.md):
  * geomBuffer / binningBuffer / imgBuffer are opaque uint8 tensors with OUR layout;
  * colors_precomp may carry 1..8 channels (the reference is fixed at 3); out_color follows;
  * kernels run on torch's current stream, not the legacy default stream.
"""
import ctypes

import torch

from hairgs_b200 import _lib as L

NUM_CHANNELS = 3
SYNC_FREE = True        # False: size the binning workspace exactly, after a blocking read-back (reference behaviour)
_capacity_hint = {}     # (device, P, H, W) -> instance capacity to try first
_depth_bits_hint = {}   # (device, P, H, W) -> sort_depth_bits to try first (depth-range compaction of the sort keys)
_sort_mode_hint = {}    # (device, P, H, W) -> L.SORT_GLOBAL once a view had a tile list too long for the in-tile sort
# HGS_SORT_MODE=global|tile: binning formulation.  Default: the global radix sort.  The tile-partitioned formulation
# (csrc/tilesort.cu) is bit-identical and fully tested but measured SLOWER on B200 (profiles/r2_tilesort.md): sorting the
# lists with a shared-memory bitonic network costs more than the five one-wave radix passes it replaces.
DEFAULT_SORT_MODE = L.SORT_TILE if __import__("os").environ.get("HGS_SORT_MODE", "global") == "tile" else L.SORT_GLOBAL


def sort_mode_for(key):
    return _sort_mode_hint.get(key, DEFAULT_SORT_MODE)


_slice_hint = {}        # (device, P, H, W) -> (smallest, largest) depth bit pattern seen: depth slices of the tile lists


def slice_params(key):
    """(slice_base, slice_shift) of hgs_raster_params for this scene: HGS_TILE_SLICES slices over the depth range of the views
    seen so far; (0, 0) = no hint yet (everything in slice 0).  Pure performance hints - any values give the same output."""
    lo_hi = _slice_hint.get(key)
    if lo_hi is None:
        return 0, 0
    lo, hi = lo_hi
    return int(lo), max((int(hi - lo)).bit_length() - (L.TILE_SLICES.bit_length() - 1), 1)


def note_depth_range(key, host):
    dmax, dmin = int(host[3]) & 0xffffffff, (~int(host[4])) & 0xffffffff
    if dmax >= dmin and dmax < 0x7f800000:
        old = _slice_hint.get(key)
        _slice_hint[key] = (dmin, dmax) if old is None else (min(old[0], dmin), max(old[1], dmax))
_pinned_pool = []
_pinned_next = 0


def _next_capacity(old, n):
    """Capacity to try next time for this (device, P, H, W): 25 % head-room over the count just seen, in 256 Ki
    steps, and never shrinking — views of one scene then settle on ONE workspace size, which the caching allocator
    serves from the same block every call (sizes that wander per view make it fall back to cudaMalloc)."""
    want = ((int(n * 1.25) + 4096 + 262143) // 262144) * 262144
    return max(old or 0, want)


def _tile_bits(H, W):
    # tile_id_bits in csrc/hgs_common.cuh (the reference's getHigherMsb of the tile count, rasterizer_impl.cu:300)
    return (((W + 15) // 16) * ((H + 15) // 16)).bit_length()


def _depth_range_bits(host):
    """Bits needed for (depth_max - depth_min) of the visible Gaussians, from the read-back words 3-4."""
    dmax, dmin = int(host[3]) & 0xffffffff, (~int(host[4])) & 0xffffffff
    return max(1, (dmax - dmin).bit_length()) if dmax >= dmin else 1


def _next_depth_bits(H, W, need):
    """Hint for the next call: as many depth bits as fit in the number of 8-bit passes `need` bits take anyway."""
    tb = _tile_bits(H, W)
    passes = (tb + need + 7) // 8
    return min(32, passes * 8 - tb)


def _pinned_triplet():
    """Small ring of pinned int32[8] buffers for the async (num_rendered, -, overflow, depth_max, ~depth_min) read-back."""
    global _pinned_next
    if len(_pinned_pool) < 64:
        _pinned_pool.append(torch.zeros(8, dtype=torch.int32).pin_memory())
        return _pinned_pool[-1]
    _pinned_next = (_pinned_next + 1) % 64
    return _pinned_pool[_pinned_next]


def _prep(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
          projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos, prefiltered, debug):
    if means3D.ndimension() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:57-59
    if not means3D.is_cuda:
        raise L.HgsError("means3D must be a CUDA tensor: this rasterizer has no CPU path")
    dev = means3D.device
    P = means3D.size(0)
    keep = {}
    for name, t in (("background", background), ("means3D", means3D), ("colors_precomp", colors),
                    ("opacities", opacity), ("scales", scales), ("rotations", rotations),
                    ("cov3D_precomp", cov3D_precomp), ("viewmatrix", viewmatrix), ("projmatrix", projmatrix),
                    ("shs", sh), ("cam_pos", campos)):
        keep[name] = L.f32c(t, name, dev)
    M = 0
    if sh is not None and sh.numel() != 0:
        M = sh.size(1)
    channels = NUM_CHANNELS
    if keep["colors_precomp"] is not None:
        channels = keep["colors_precomp"].size(-1) if keep["colors_precomp"].dim() > 1 else 1
    prm = L.RasterParams(P=P, D=int(degree), M=int(M), width=int(image_width), height=int(image_height),
                         channels=int(channels), tan_fovx=float(tan_fovx), tan_fovy=float(tan_fovy),
                         scale_modifier=float(scale_modifier), prefiltered=int(bool(prefiltered)),
                         debug=int(bool(debug)))
    inp = L.RasterInputs(**{k: L.ptr(v) for k, v in keep.items()})
    return dev, prm, inp, keep


def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                        prefiltered, debug):
    """RasterizeGaussiansCUDA (rasterize_points.cu:35-115): returns
    (num_rendered, out_color[C,H,W], radii[P] int32, geomBuffer, binningBuffer, imgBuffer)."""
    lib = L.load()
    dev, prm, inp, keep = _prep(background, means3D, colors, opacity, scales, rotations, scale_modifier,
                                cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height,
                                image_width, sh, degree, campos, prefiltered, debug)
    P, H, W, C = prm.P, prm.height, prm.width, prm.channels
    prm.slice_base, prm.slice_shift = slice_params((dev.index, P, H, W))
    prm.sort_mode = sort_mode_for((dev.index, P, H, W))     # stage A counts the lists of the tile-partitioned formulation
    with torch.cuda.device(dev):
        stream = L.stream_ptr(dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        out_color = torch.empty((C, H, W), dtype=torch.float32, device=dev)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        geom = torch.empty((lib.hgs_geom_bytes(P, C, W, H),), **u8)
        img = torch.empty((lib.hgs_image_bytes(W, H),), **u8)
        if P == 0:
            # rasterize_points.cu:81 short-circuit: zero outputs, empty scratch
            out_color.zero_()
            return 0, out_color, radii, geom, torch.empty((0,), **u8), img
        L.check(lib.hgs_forward_stage_a(ctypes.byref(prm), ctypes.byref(inp), geom.data_ptr(), radii.data_ptr(),
                                        stream), "forward stage A")
        # The reference blocks on the instance count in the middle of the pass (rasterizer_impl.cu:281) and the GPU
        # idles until the host has launched the rest.  Here the count stays on the device: stage B is enqueued at
        # once into a binning workspace sized from the last count seen for this (device, P, W, H), and the host only
        # waits for the small async read-back AFTERWARDS (the GPU is busy with stage B by then).  Too small a
        # guess -> stage B is simply run again with the exact size.
        host = _pinned_triplet()
        L.check(lib.hgs_forward_read_num_rendered(geom.data_ptr(), P, host.data_ptr(), stream), "read num_rendered")
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        key = (dev.index, P, H, W)
        cap = _capacity_hint.get(key) if SYNC_FREE else None
        prm.sort_depth_bits = _depth_bits_hint.get(key, 0) if SYNC_FREE else 0
        binning = None
        if cap is not None:
            binning = torch.empty((lib.hgs_binning_bytes(cap, C),), **u8)
            L.check(lib.hgs_forward_stage_b(ctypes.byref(prm), ctypes.byref(inp), geom.data_ptr(), binning.data_ptr(),
                                            img.data_ptr(), cap, radii.data_ptr(), out_color.data_ptr(), stream),
                    "forward stage B")
        ready.synchronize()
        N, overflow = int(host[0]), int(host[2])
        if (overflow & 1) != 0 or N < 0:
            raise L.HgsError("instance count overflows int32")
        need = _depth_range_bits(host)
        note_depth_range(key, host)
        if (overflow & 4) and prm.slice_shift > 0:
            # a (tile, slice) list is longer than HGS_TILE_SORT_MAX (counted in stage A) although the depth slices were in
            # effect: this scene is a global-sort scene from now on (without slice hints - first view - only this pass is)
            _sort_mode_hint[key] = L.SORT_GLOBAL
        if prm.sort_mode == L.SORT_TILE:
            fits = (overflow & 4) == 0
        else:
            fits = prm.sort_depth_bits in (0, 32) or need <= prm.sort_depth_bits
        if cap is None or N > cap or not fits:
            if cap is None or N > cap:
                binning = torch.empty((lib.hgs_binning_bytes(N, C),), **u8)
                cap_b = N
            else:
                cap_b = cap
            if overflow & 4:
                prm.sort_mode = L.SORT_GLOBAL
            prm.sort_depth_bits = _next_depth_bits(H, W, need) if SYNC_FREE else 0
            L.check(lib.hgs_forward_stage_b(ctypes.byref(prm), ctypes.byref(inp), geom.data_ptr(),
                                            binning.data_ptr() if cap_b > 0 else None, img.data_ptr(), cap_b,
                                            radii.data_ptr(), out_color.data_ptr(), stream), "forward stage B")
        # next guess: 25 % head-room, rounded up to the 4096-instance granularity hgs_binning_capacity inverts
        _capacity_hint[key] = _next_capacity(_capacity_hint.get(key), N)
        _depth_bits_hint[key] = max(_depth_bits_hint.get(key, 0), _next_depth_bits(H, W, need))
    del keep
    return N, out_color, radii, geom, binning, img


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                 cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, sh,
                                 degree, campos, geomBuffer, R, binningBuffer, imageBuffer, debug):
    """RasterizeGaussiansBackwardCUDA (rasterize_points.cu:117-196): returns (dL_dmeans2D[P,3],
    dL_dcolors[P,C], dL_dopacity[P,1], dL_dmeans3D[P,3], dL_dcov3D[P,6], dL_dsh[P,M,3], dL_dscales[P,3],
    dL_drotations[P,4])."""
    lib = L.load()
    H, W = dL_dout_color.size(1), dL_dout_color.size(2)
    dev, prm, inp, keep = _prep(background, means3D, colors, None, scales, rotations, scale_modifier, cov3D_precomp,
                                viewmatrix, projmatrix, tan_fovx, tan_fovy, H, W, sh, degree, campos, False, debug)
    P, M, C = prm.P, prm.M, dL_dout_color.size(0)
    prm.channels = C
    f32 = dict(dtype=torch.float32, device=dev)
    if P == 0:
        return (torch.zeros((0, 3), **f32), torch.zeros((0, C), **f32), torch.zeros((0, 1), **f32),
                torch.zeros((0, 3), **f32), torch.zeros((0, 6), **f32), torch.zeros((0, M, 3), **f32),
                torch.zeros((0, 3), **f32), torch.zeros((0, 4), **f32))
    with torch.cuda.device(dev):
        dpix = L.f32c(dL_dout_color, "dL_dout_color", dev)
        # one allocation for the compositor's accumulation targets (cleared with a single memset inside
        # the library), one for the arrays preprocess_bwd writes exactly once
        acc = torch.empty((P * (3 + 4 + 1 + C),), **f32)
        dL_dmeans2D = acc[:3 * P].view(P, 3)
        dL_dconic = acc[3 * P:7 * P].view(P, 4)
        dL_dopacity = acc[7 * P:8 * P].view(P, 1)
        dL_dcolors = acc[8 * P:].view(P, C)
        rest = torch.empty((P * (3 + 6 + 3 * M + 3 + 4),), **f32)
        o = 0
        dL_drotations = rest[o:o + 4 * P].view(P, 4); o += 4 * P   # first: keeps float4 stores 16-B aligned
        dL_dmeans3D = rest[o:o + 3 * P].view(P, 3); o += 3 * P
        dL_dcov3D = rest[o:o + 6 * P].view(P, 6); o += 6 * P
        dL_dscales = rest[o:o + 3 * P].view(P, 3); o += 3 * P
        dL_dsh = rest[o:o + 3 * M * P].view(P, M, 3)
        grads = L.RasterGrads(dL_dmean2D=dL_dmeans2D.data_ptr(), dL_dconic=dL_dconic.data_ptr(),
                              dL_dopacity=dL_dopacity.data_ptr(), dL_dcolor=dL_dcolors.data_ptr(),
                              dL_dmean3D=dL_dmeans3D.data_ptr(), dL_dcov3D=dL_dcov3D.data_ptr(),
                              dL_dsh=L.ptr(dL_dsh), dL_dscale=dL_dscales.data_ptr(),
                              dL_drot=dL_drotations.data_ptr())
        cap = lib.hgs_binning_capacity(binningBuffer.numel(), C, int(R))
        L.check(int(cap), "binning capacity")
        L.check(lib.hgs_rasterize_backward(ctypes.byref(prm), ctypes.byref(inp), int(cap), L.ptr(radii),
                                           geomBuffer.data_ptr(), L.ptr(binningBuffer), imageBuffer.data_ptr(),
                                           dpix.data_ptr(), ctypes.byref(grads), L.stream_ptr(dev)),
                "backward")
    del keep
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations


def mark_visible(means3D, viewmatrix, projmatrix):
    """markVisible (rasterize_points.cu:198-217): bool[P], True where view-space z > 0.2."""
    lib = L.load()
    if not means3D.is_cuda:
        raise L.HgsError("means3D must be a CUDA tensor: this rasterizer has no CPU path")
    dev = means3D.device
    P = means3D.size(0)
    present = torch.zeros((P,), dtype=torch.bool, device=dev)
    if P != 0:
        m = L.f32c(means3D, "means3D", dev)
        v = L.f32c(viewmatrix, "viewmatrix", dev)
        p = L.f32c(projmatrix, "projmatrix", dev)
        with torch.cuda.device(dev):
            L.check(lib.hgs_mark_visible(P, m.data_ptr(), v.data_ptr(), p.data_ptr(), present.data_ptr(),
                                         L.stream_ptr(dev)), "mark_visible")
    return present
