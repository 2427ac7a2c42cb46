"""Test-only access to oracle/_ref: the UNMODIFIED reference rasterizer / simple-knn built for sm_100a by
oracle/build_ref.py.  Gives the reference `_C` modules plus unpackers that replay the reference's
GeometryState/BinningState/ImageState carve-up (rasterizer_impl.cu:155-194, rasterizer_impl.h:22-30) so
intermediate buffers can be compared bit for bit.  Never imported by the product package."""
import importlib.util
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def _load(name):
    so = os.path.join(REF_DIR, name, name + ".so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def ref_dgr():
    if "dgr" not in _cache:
        # train.py:278 -> utils/general.py:119-124: DISTWAR butterfly backward, threshold 8 (latched at first backward)
        os.environ.setdefault("BW_IMPLEMENTATION", "1")
        os.environ.setdefault("BALANCE_THRESHOLD", "8")
        _cache["dgr"] = _load("ref_dgr_C")
    return _cache["dgr"]


def ref_knn():
    if "knn" not in _cache:
        _cache["knn"] = _load("ref_knn_C")
    return _cache["knn"]


def _al(x, a=128):
    return (x + a - 1) // a * a


def unpack_geom(buf, P):
    """geomBuffer -> dict of tensors (views).  Order: depths, clamped, internal_radii, means2D, cov3D,
    conic_opacity, rgb, tiles_touched, [scan temp], point_offsets (located from the end)."""
    assert buf.data_ptr() % 128 == 0
    out, off = {}, 0

    def take(name, nbytes, dtype, shape):
        nonlocal off
        off = _al(off)
        out[name] = buf[off:off + nbytes].view(dtype).view(shape)
        off += nbytes

    take("depths", 4 * P, torch.float32, (P,))
    take("clamped", 3 * P, torch.uint8, (P, 3))
    take("internal_radii", 4 * P, torch.int32, (P,))
    take("means2D", 8 * P, torch.float32, (P, 2))
    take("cov3D", 24 * P, torch.float32, (P, 6))
    take("conic_opacity", 16 * P, torch.float32, (P, 4))
    take("rgb", 12 * P, torch.float32, (P, 3))
    take("tiles_touched", 4 * P, torch.int32, (P,))
    end = buf.numel() - 128
    out["point_offsets"] = buf[end - 4 * P:end].view(torch.int32)
    return out


def unpack_binning(buf, N):
    out, off = {}, 0

    def take(name, nbytes, dtype):
        nonlocal off
        off = _al(off)
        out[name] = buf[off:off + nbytes].view(dtype)
        off += nbytes

    take("point_list", 4 * N, torch.int32)
    take("point_list_unsorted", 4 * N, torch.int32)
    take("point_list_keys", 8 * N, torch.int64)
    take("point_list_keys_unsorted", 8 * N, torch.int64)
    return out


def unpack_image(buf, H, W):
    out, off = {}, 0
    n = H * W

    def take(name, nbytes, dtype):
        nonlocal off
        off = _al(off)
        out[name] = buf[off:off + nbytes].view(dtype)
        off += nbytes

    take("accum_alpha", 4 * n, torch.float32)
    take("n_contrib", 4 * n, torch.int32)
    take("ranges", 8 * n, torch.int32)
    out["ranges"] = out["ranges"].view(-1, 2)
    return out
