"""ctypes binding of libhairgs_rast.so — the C ABI declared in include/hairgs_rast.h.

PyTorch is used only for device memory and streams; every compute call goes through the shared
library.  There is NO fallback: if the library is missing or the device is not CUDA this raises.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# HGS_LIBRARY: alternative build of the same library (A/B kernel experiments); default is the in-tree build
LIB_PATH = os.environ.get("HGS_LIBRARY") or os.path.join(os.path.dirname(_HERE), "lib", "libhairgs_rast.so")

HGS_MAX_CHANNELS = 8

(VIEW_DEPTHS, VIEW_MEANS2D, VIEW_CONIC_OPACITY, VIEW_RGB, VIEW_TILES_TOUCHED, VIEW_POINT_OFFSETS,
 VIEW_CLAMPED, VIEW_KEYS_SORTED, VIEW_POINT_LIST, VIEW_RANGES, VIEW_FINAL_T, VIEW_N_CONTRIB,
 VIEW_KEYS_UNSORTED, VIEW_COV3D) = range(14)


class RasterParams(ctypes.Structure):
    _fields_ = [("P", c_int32), ("D", c_int32), ("M", c_int32), ("width", c_int32), ("height", c_int32),
                ("channels", c_int32), ("tan_fovx", c_float), ("tan_fovy", c_float),
                ("scale_modifier", c_float), ("prefiltered", c_int32), ("debug", c_int32),
                ("sort_depth_bits", c_int32), ("sort_mode", c_int32), ("slice_base", c_int32), ("slice_shift", c_int32)]

SORT_TILE, SORT_GLOBAL = 0, 1
TILE_SLICES = 16


class RasterInputs(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ("background", "means3D", "shs", "colors_precomp", "opacities", "scales",
                                        "rotations", "cov3D_precomp", "viewmatrix", "projmatrix", "cam_pos")]


class RasterGrads(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ("dL_dmean2D", "dL_dconic", "dL_dopacity", "dL_dcolor", "dL_dmean3D",
                                        "dL_dcov3D", "dL_dsh", "dL_dscale", "dL_drot")]


class StrandInputs(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ("background", "endpoints", "endpoint_pairs", "width", "opacity_logit",
                                        "mask_logit", "features", "viewmatrix", "projmatrix", "cam_pos")] + \
               [("num_endpoints", c_int64)]


class StrandGrads(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ("dL_dmean2D", "dL_dconic", "dL_dopacity", "dL_dcolor", "dL_dendpoints",
                                        "dL_dwidth", "dL_dopacity_logit", "dL_dmask_logit", "dL_dfeatures")] + \
               [("accumulate", c_int32), ("acc16", c_void_p)]


ALLOC_FN = ctypes.CFUNCTYPE(c_void_p, c_void_p, c_size_t)

_lib = None


def load():
    """Load (once) and return the shared library; raise loudly when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python hair-gs_b200/build.py` "
            "(nvcc, sm_100a). There is no CPU / PyTorch fallback for the rasterizer.")
    lib = ctypes.CDLL(LIB_PATH)
    P = ctypes.POINTER
    lib.hgs_abi_version.restype = c_int
    lib.hgs_last_error.restype = c_char_p
    lib.hgs_geom_bytes.restype = c_size_t
    lib.hgs_geom_bytes.argtypes = [c_int32, c_int32, c_int32, c_int32]
    lib.hgs_image_bytes.restype = c_size_t
    lib.hgs_image_bytes.argtypes = [c_int32, c_int32]
    lib.hgs_binning_bytes.restype = c_size_t
    lib.hgs_binning_bytes.argtypes = [c_int64, c_int32]
    lib.hgs_binning_capacity.restype = c_int64
    lib.hgs_binning_capacity.argtypes = [c_size_t, c_int32, c_int64]
    lib.hgs_sort_bytes.restype = c_size_t
    lib.hgs_sort_bytes.argtypes = [c_int64]
    lib.hgs_knn_bytes.restype = c_size_t
    lib.hgs_knn_bytes.argtypes = [c_int32]
    lib.hgs_rasterize_forward.restype = c_int
    lib.hgs_rasterize_forward.argtypes = [ALLOC_FN, c_void_p, ALLOC_FN, c_void_p, ALLOC_FN, c_void_p,
                                          P(RasterParams), P(RasterInputs), c_void_p, c_void_p, c_void_p]
    lib.hgs_forward_stage_a.restype = c_int
    lib.hgs_forward_stage_a.argtypes = [P(RasterParams), P(RasterInputs), c_void_p, c_void_p, c_void_p]
    lib.hgs_forward_read_num_rendered.restype = c_int
    lib.hgs_forward_read_num_rendered.argtypes = [c_void_p, c_int32, c_void_p, c_void_p]
    lib.hgs_forward_stage_b.restype = c_int
    lib.hgs_forward_stage_b.argtypes = [P(RasterParams), P(RasterInputs), c_void_p, c_void_p, c_void_p, c_int64,
                                        c_void_p, c_void_p, c_void_p]
    lib.hgs_forward_stage_b_binning.restype = c_int
    lib.hgs_forward_stage_b_binning.argtypes = [P(RasterParams), c_void_p, c_void_p, c_void_p, c_int64, c_void_p]
    lib.hgs_forward_stage_b_composite.restype = c_int
    lib.hgs_forward_stage_b_composite.argtypes = [P(RasterParams), c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                                  c_void_p]
    lib.hgs_graph_instantiate.restype = c_int
    lib.hgs_graph_instantiate.argtypes = [c_void_p, c_int32, P(c_void_p)]
    lib.hgs_graph_launch.restype = c_int
    lib.hgs_graph_launch.argtypes = [c_void_p, c_void_p]
    lib.hgs_graph_exec_destroy.restype = c_int
    lib.hgs_graph_exec_destroy.argtypes = [c_void_p]
    lib.hgs_rasterize_backward.restype = c_int
    lib.hgs_rasterize_backward.argtypes = [P(RasterParams), P(RasterInputs), c_int64, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p, P(RasterGrads), c_void_p]
    lib.hgs_mark_visible.restype = c_int
    lib.hgs_mark_visible.argtypes = [c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.hgs_dist2_knn3.restype = c_int
    lib.hgs_dist2_knn3.argtypes = [c_int32, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.hgs_sort_pairs.restype = c_int
    lib.hgs_sort_pairs.argtypes = [c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.hgs_state_view.restype = c_int64
    lib.hgs_state_view.argtypes = [c_int, P(RasterParams), P(RasterInputs), c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p]
    lib.hgs_strands_forward_stage_a.restype = c_int
    lib.hgs_strands_forward_stage_a.argtypes = [P(RasterParams), P(StrandInputs), c_void_p, c_void_p, c_void_p]
    lib.hgs_strands_forward_stage_b.restype = c_int
    lib.hgs_strands_forward_stage_b.argtypes = [P(RasterParams), P(StrandInputs), c_void_p, c_void_p, c_void_p, c_int64,
                                                c_void_p, c_void_p]
    lib.hgs_strands_backward.restype = c_int
    lib.hgs_strands_backward.argtypes = [P(RasterParams), P(StrandInputs), c_int64, c_void_p, c_void_p, c_void_p,
                                         c_void_p, P(StrandGrads), c_void_p]
    lib.hgs_strands_backward_parts.restype = c_int
    lib.hgs_strands_backward_parts.argtypes = [P(RasterParams), P(StrandInputs), c_int64, c_void_p, c_void_p, c_void_p,
                                               c_void_p, P(StrandGrads), c_int32, c_void_p]
    lib.hgs_weighted_l1.restype = c_int
    lib.hgs_weighted_l1.argtypes = [c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.hgs_hair_image_loss.restype = c_int
    lib.hgs_hair_image_loss.argtypes = [P(HairLoss), c_void_p]
    lib.hgs_unpack_targets.restype = c_int
    lib.hgs_unpack_targets.argtypes = [c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.hgs_adam_step.restype = c_int
    lib.hgs_adam_step.argtypes = [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, P(c_int64), P(c_float), c_int32,
                                  c_float, c_float, c_float, c_float, c_int32, c_void_p]
    lib.hgs_densify_stats.restype = c_int
    lib.hgs_densify_stats.argtypes = [c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.hgs_merge_count.restype = c_int
    lib.hgs_merge_count.argtypes = [c_int32, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_double, ctypes.c_double, c_int32,
                                    c_int32, c_void_p, c_void_p]
    lib.hgs_merge_fill.restype = c_int
    lib.hgs_merge_fill.argtypes = [c_int32, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_double, ctypes.c_double, c_int32,
                                   c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.hgs_merge_greedy.restype = c_int
    lib.hgs_merge_greedy.argtypes = [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.hgs_debug_set_stats.restype = c_int
    lib.hgs_debug_set_stats.argtypes = [c_void_p]
    lib.hgs_debug_set_composite_blocks.restype = c_int
    lib.hgs_debug_set_composite_blocks.argtypes = [c_int]
    lib.hgs_profile_enable.restype = c_int
    lib.hgs_profile_enable.argtypes = [c_int]
    lib.hgs_profile_collect.restype = c_int
    lib.hgs_profile_collect.argtypes = [c_void_p, c_void_p]
    lib.hgs_stage_name.restype = c_char_p
    lib.hgs_stage_name.argtypes = [c_int]
    if lib.hgs_abi_version() != 4:
        raise ImportError("libhairgs_rast.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


class HairLoss(ctypes.Structure):
    """hgs_hair_loss (include/hairgs_rast.h)."""
    _fields_ = [("height", c_int32), ("width", c_int32), ("image7", c_void_p), ("gt_rgb", c_void_p),
                ("gt_mask", c_void_p), ("gt_theta", c_void_p), ("confidence", c_void_p), ("orient_mask", c_void_p),
                ("view_rot", ctypes.c_float * 9), ("bg_orient", ctypes.c_float * 3), ("l_l1", ctypes.c_float),
                ("l_dssim", ctypes.c_float), ("l_mask", ctypes.c_float), ("l_orient", ctypes.c_float),
                ("terms", c_void_p), ("scratch", c_void_p), ("dL_dimage", c_void_p), ("view_matrix_dev", c_void_p)]


class HgsError(RuntimeError):
    pass


def check(status, what="hairgs_rast"):
    if status < 0:
        msg = load().hgs_last_error().decode(errors="replace")
        raise HgsError(f"{what} failed ({status}): {msg}")
    return status


def ptr(t):
    """Device pointer of a tensor; None (NULL) for absent / empty tensors, as the reference does
    (rasterize_points.cu passes the null data_ptr of 0-element tensors)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def f32c(t, name, device):
    """Contiguous float32 CUDA view of an input (rasterize_points.cu:94-111 applies .contiguous();
    data<float>() rejects other dtypes)."""
    if t is None or t.numel() == 0:
        return None
    if not t.is_cuda:
        raise HgsError(f"{name} must be a CUDA tensor (no CPU path)")
    if t.dtype != torch.float32:
        raise HgsError(f"{name} must be float32, got {t.dtype}")
    if t.device != device:
        raise HgsError(f"{name} is on {t.device}, expected {device}")
    return t.contiguous()


_VIEW_SPEC = {
    VIEW_DEPTHS: (torch.float32, lambda p, n: (p.P,)),
    VIEW_MEANS2D: (torch.float32, lambda p, n: (p.P, 2)),
    VIEW_CONIC_OPACITY: (torch.float32, lambda p, n: (p.P, 4)),
    VIEW_RGB: (torch.float32, lambda p, n: (p.P, p.channels)),
    VIEW_TILES_TOUCHED: (torch.int32, lambda p, n: (p.P,)),
    VIEW_POINT_OFFSETS: (torch.int32, lambda p, n: (p.P,)),
    VIEW_CLAMPED: (torch.uint8, lambda p, n: (p.P, 3)),
    VIEW_KEYS_SORTED: (torch.int64, lambda p, n: (n,)),
    VIEW_POINT_LIST: (torch.int32, lambda p, n: (n,)),
    VIEW_RANGES: (torch.int32, lambda p, n: (((p.width + 15) // 16) * ((p.height + 15) // 16), 2)),
    VIEW_FINAL_T: (torch.float32, lambda p, n: (p.height, p.width)),
    VIEW_N_CONTRIB: (torch.int32, lambda p, n: (p.height, p.width)),
    VIEW_COV3D: (torch.float32, lambda p, n: (p.P, 6)),
}


def state_view(what, prm, inp, num_rendered, geom, binning, img):
    """Copy one sub-array of the opaque workspaces out in the REFERENCE's element layout (hgs_state_view)."""
    lib = load()
    dtype, shape_fn = _VIEW_SPEC[what]
    shape = shape_fn(prm, int(num_rendered))
    out = torch.zeros(shape, dtype=dtype, device=geom.device)
    if out.numel() == 0:
        return out
    cap = num_rendered
    if binning is not None and binning.numel() > 0:
        cap = check(int(lib.hgs_binning_capacity(binning.numel(), prm.channels, int(num_rendered))), "binning capacity")
    with torch.cuda.device(geom.device):
        n = lib.hgs_state_view(what, ctypes.byref(prm), ctypes.byref(inp), int(num_rendered), int(cap), ptr(geom), ptr(binning),
                               ptr(img), out.data_ptr(), stream_ptr(geom.device))
    check(int(n), "state_view")
    assert int(n) == out.numel() * out.element_size(), (n, out.shape)
    return out
