"""numpy/ctypes front end of the C oracle (oracle/rast_oracle.c).  TEST INFRASTRUCTURE ONLY: importable from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs — never from hair-gs_b200/."""
import ctypes
import os
import subprocess
from ctypes import POINTER, c_float, c_int, c_int32, c_int64, c_uint8, c_uint32, c_uint64, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liborc.so")


class Params(ctypes.Structure):
    _fields_ = [("P", c_int32), ("D", c_int32), ("M", c_int32), ("W", c_int32), ("H", c_int32), ("C", c_int32),
                ("tan_fovx", c_float), ("tan_fovy", c_float), ("scale_modifier", c_float)]


def build(force=False):
    src = os.path.join(HERE, "rast_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.run(["make", "-C", HERE, f"CC={cc}"], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        fp = POINTER(c_float)
        L.orc_forward.restype = c_void_p
        L.orc_forward.argtypes = [POINTER(Params)] + [fp] * 11 + [fp, POINTER(c_int32)]
        L.orc_backward.restype = None
        L.orc_backward.argtypes = [c_void_p] + [fp] * 10 + [fp] + [fp] * 9
        L.orc_free.argtypes = [c_void_p]
        L.orc_num_rendered.restype = c_int64
        L.orc_num_rendered.argtypes = [c_void_p]
        L.orc_array.restype = c_void_p
        L.orc_array.argtypes = [c_void_p, c_int]
        L.orc_higher_msb.restype = c_uint32
        L.orc_higher_msb.argtypes = [c_uint32]
        L.orc_sort_pairs.argtypes = [c_int64, c_int, POINTER(c_uint64), POINTER(c_uint32)]
        L.orc_mark_visible.argtypes = [c_int, fp, fp, POINTER(c_uint8)]
        L.orc_knn3.argtypes = [c_int, fp, fp]
        L.orc_num_threads.restype = c_int
        L.orc_set_threads.argtypes = [c_int]
        _lib = L
    return _lib


def _f(a):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.size == 0:
        return None
    return a


def _p(a):
    return a.ctypes.data_as(POINTER(c_float)) if a is not None else None


_ARR = {"depths": (0, np.float32, lambda P, N, T, HW, C: (P,)), "means2D": (1, np.float32, lambda P, N, T, HW, C: (P, 2)),
        "conic_opacity": (2, np.float32, lambda P, N, T, HW, C: (P, 4)), "rgb": (3, np.float32, lambda P, N, T, HW, C: (P, C)),
        "tiles_touched": (4, np.uint32, lambda P, N, T, HW, C: (P,)), "point_offsets": (5, np.uint32, lambda P, N, T, HW, C: (P,)),
        "clamped": (6, np.uint8, lambda P, N, T, HW, C: (P, 3)), "point_list_keys": (7, np.uint64, lambda P, N, T, HW, C: (N,)),
        "point_list": (8, np.uint32, lambda P, N, T, HW, C: (N,)), "ranges": (9, np.uint32, lambda P, N, T, HW, C: (T, 2)),
        "accum_alpha": (10, np.float32, lambda P, N, T, HW, C: (HW,)), "n_contrib": (11, np.uint32, lambda P, N, T, HW, C: (HW,)),
        "keys_unsorted": (12, np.uint64, lambda P, N, T, HW, C: (N,)), "cov3D": (13, np.float32, lambda P, N, T, HW, C: (P, 6)),
        "radii": (14, np.int32, lambda P, N, T, HW, C: (P,)), "vals_unsorted": (15, np.uint32, lambda P, N, T, HW, C: (N,))}


class Forward:
    """Result of one oracle forward pass; keeps the C state alive for backward()."""

    def __init__(self, d):
        L = lib()
        self.inputs = {k: _f(d.get(k)) for k in ("background", "means3D", "sh", "colors", "opacity", "scales",
                                                  "rotations", "cov3D_precomp", "viewmatrix", "projmatrix", "campos")}
        i = self.inputs
        P = i["means3D"].shape[0] if i["means3D"] is not None else 0
        M = i["sh"].shape[1] if i["sh"] is not None else 0
        C = i["colors"].shape[-1] if i["colors"] is not None else 3
        H, W = int(d["image_height"]), int(d["image_width"])
        self.prm = Params(P=P, D=int(d["degree"]), M=M, W=W, H=H, C=C, tan_fovx=float(d["tan_fovx"]),
                          tan_fovy=float(d["tan_fovy"]), scale_modifier=float(d["scale_modifier"]))
        self.color = np.zeros((C, H, W), np.float32)
        self.radii = np.zeros((max(P, 1),), np.int32)
        self.state = L.orc_forward(ctypes.byref(self.prm), _p(i["background"]), _p(i["means3D"]), _p(i["sh"]),
                                   _p(i["colors"]), _p(i["opacity"]), _p(i["scales"]), _p(i["rotations"]),
                                   _p(i["cov3D_precomp"]), _p(i["viewmatrix"]), _p(i["projmatrix"]), _p(i["campos"]),
                                   _p(self.color), self.radii.ctypes.data_as(POINTER(c_int32)))
        self.radii = self.radii[:P]
        self.N = int(L.orc_num_rendered(self.state))
        self.P, self.C, self.H, self.W, self.M = P, C, H, W, M

    def array(self, name):
        idx, dt, shp = _ARR[name]
        T = ((self.W + 15) // 16) * ((self.H + 15) // 16)
        shape = shp(self.P, self.N, T, self.H * self.W, self.C)
        n = int(np.prod(shape))
        if n == 0:
            return np.zeros(shape, dt)
        ptr = lib().orc_array(self.state, idx)
        buf = (ctypes.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dt).reshape(shape).copy()

    def backward(self, dL_dpix):
        L = lib()
        i = self.inputs
        P, C, M = self.P, self.C, self.M
        g = dict(dL_dmeans2D=np.zeros((P, 3), np.float32), dL_dconic=np.zeros((P, 4), np.float32),
                 dL_dopacity=np.zeros((P, 1), np.float32), dL_dcolors=np.zeros((P, C), np.float32),
                 dL_dmeans3D=np.zeros((P, 3), np.float32), dL_dcov3D=np.zeros((P, 6), np.float32),
                 dL_dsh=np.zeros((P, M, 3), np.float32), dL_dscales=np.zeros((P, 3), np.float32),
                 dL_drotations=np.zeros((P, 4), np.float32))
        dpix = _f(dL_dpix)
        L.orc_backward(self.state, _p(i["background"]), _p(i["means3D"]), _p(i["sh"]), _p(i["colors"]), _p(i["scales"]),
                       _p(i["rotations"]), _p(i["cov3D_precomp"]), _p(i["viewmatrix"]), _p(i["projmatrix"]),
                       _p(i["campos"]), _p(dpix), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]),
                       _p(g["dL_dcolors"]), _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]),
                       _p(g["dL_dsh"]) if M > 0 else None, _p(g["dL_dscales"]), _p(g["dL_drotations"]))
        return g

    def close(self):
        if self.state:
            lib().orc_free(self.state)
            self.state = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def higher_msb(n):
    return int(lib().orc_higher_msb(int(n)))


def sort_pairs(keys, vals, end_bit):
    k = np.ascontiguousarray(keys, np.uint64).copy()
    v = np.ascontiguousarray(vals, np.uint32).copy()
    lib().orc_sort_pairs(k.size, int(end_bit), k.ctypes.data_as(POINTER(c_uint64)), v.ctypes.data_as(POINTER(c_uint32)))
    return k, v


def mark_visible(means3D, viewmatrix):
    m, v = _f(means3D), _f(viewmatrix)
    out = np.zeros((m.shape[0],), np.uint8)
    lib().orc_mark_visible(m.shape[0], _p(m), _p(v), out.ctypes.data_as(POINTER(c_uint8)))
    return out.astype(bool)


def knn3(points):
    p = _f(points)
    out = np.zeros((p.shape[0],), np.float32)
    lib().orc_knn3(p.shape[0], _p(p), _p(out))
    return out


def num_threads():
    return int(lib().orc_num_threads())


def set_threads(n):
    lib().orc_set_threads(int(n))
