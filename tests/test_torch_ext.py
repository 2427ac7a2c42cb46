"""The compiled torch binding over the C ABI (hair-gs_b200/torch_ext/hgs_torch_ext.cpp — what INTEGRATION.md §2 tells a
maintainer to build if they prefer pybind over the shipped ctypes binding): it loads on CPU, exports the reference's entry
points, and on the GPU gives what the ctypes `_C` gives (and, through it, what the reference gives)."""
import os
import sys

import pytest
import torch

import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hair-gs_b200", "torch_ext"))


def _ext():
    import build as ext_build
    if not os.path.exists(os.path.join(ROOT, "hair-gs_b200", "lib", "hgs_torch_ext.so")):
        pytest.skip("hgs_torch_ext.so not built (python hair-gs_b200/torch_ext/build.py)")
    return ext_build.load_module()


def test_torch_ext_loads_and_exports_the_reference_entry_points():
    m = _ext()
    for name in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible", "distCUDA2"):
        assert callable(getattr(m, name))
    d = common.blob_inputs(10, 32, 32, "cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.rasterize_gaussians(*common.fwd_args(d))


@pytest.mark.gpu
def test_torch_ext_equals_ctypes_binding():
    import diff_gaussian_rasterization as dgr
    import diff_gaussian_rasterization._C as ours_C
    import simple_knn._C as knn
    m = _ext()
    dev = torch.device("cuda:0")
    d = common.blob_inputs(20000, 256, 192, dev, seed=71)
    N0, c0, r0, g0, b0, i0 = ours_C.rasterize_gaussians(*common.fwd_args(d))
    N1, c1, r1, g1, b1, i1 = m.rasterize_gaussians(*common.fwd_args(d))
    assert N1 == N0 and torch.equal(c1, c0) and torch.equal(r1, r0)
    torch.manual_seed(2)
    dL = torch.randn_like(c0)
    ga = ours_C.rasterize_gaussians_backward(*common.bwd_args(d, r0, dL, g0, N0, b0, i0))
    gb = m.rasterize_gaussians_backward(*common.bwd_args(d, r1, dL, g1, N1, b1, i1))
    for a, b in zip(ga, gb):
        assert a.shape == b.shape
        if a.numel():
            assert common.rel_err(b, a) <= 1e-5
    assert torch.equal(m.mark_visible(d["means3D"], d["viewmatrix"], d["projmatrix"]),
                       ours_C.mark_visible(d["means3D"], d["viewmatrix"], d["projmatrix"]))
    pts = d["means3D"][:5000].contiguous()
    assert torch.equal(m.distCUDA2(pts), knn.distCUDA2(pts))
    # the same autograd.Function runs on it (drop-in for the reference's `_C`)
    settings = dgr.GaussianRasterizationSettings(192, 256, d["tan_fovx"], d["tan_fovy"], d["background"], 1.0, d["viewmatrix"],
                                                 d["projmatrix"], 3, d["campos"], False, False)
    dgr._RasterizeGaussians.backend = m
    try:
        leaves = {k: d[k].clone().requires_grad_(True) for k in ("means3D", "opacity", "scales", "rotations", "sh")}
        color, radii = dgr.GaussianRasterizer(settings)(means3D=leaves["means3D"], means2D=torch.zeros_like(leaves["means3D"], requires_grad=True),
                                                        opacities=leaves["opacity"], shs=leaves["sh"], scales=leaves["scales"],
                                                        rotations=leaves["rotations"])
        (color * dL).sum().backward()
    finally:
        dgr._RasterizeGaussians.backend = dgr._C
    assert torch.equal(color.detach(), c0) and common.rel_err(leaves["means3D"].grad, ga[3]) <= 1e-5
