"""hairgs_b200 — host-side helpers of the B200-native Hair-GS render path.

The drop-in packages live beside this one (`diff_gaussian_rasterization`, `simple_knn`); this package
holds the ctypes binding of the C ABI and the callers either side of the path (synthetic scenes,
strand parameterisation, multi-view / multi-GPU driver).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
