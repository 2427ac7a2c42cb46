"""CPU-side suite (no GPU): the C-ABI library loads and exports every symbol include/hairgs_rast.h declares,
the host-side mirror of the reference interface behaves like the reference's Python layer, the product fails
loudly without CUDA, and the input builders (scenes, cameras, strand parameterisation) are sound."""
import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch

import common
from hairgs_b200 import _lib as L
from hairgs_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hairgs_rast.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(hgs_[a-z0-9_]+)\s*\(", hdr)) - {"hgs_alloc_fn"})
    assert len(names) >= 15
    lib = ctypes.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hairgs_rast.h but not exported"
    assert L.load().hgs_abi_version() == 4


def test_struct_layouts_match_header():
    assert ctypes.sizeof(L.RasterParams) == 15 * 4
    assert ctypes.sizeof(L.RasterInputs) == 11 * 8
    assert ctypes.sizeof(L.RasterGrads) == 9 * 8


def test_workspace_size_queries():
    lib = L.load()
    assert lib.hgs_geom_bytes(0, 3, 64, 64) > 0
    a, b = lib.hgs_geom_bytes(1000, 3, 64, 64), lib.hgs_geom_bytes(2000, 3, 64, 64)
    assert b > a and lib.hgs_geom_bytes(1000, 7, 64, 64) > a and lib.hgs_geom_bytes(1000, 3, 1024, 1024) > a
    assert lib.hgs_image_bytes(1024, 1024) >= 1024 * 1024 * 8 + 4096 * 8
    assert lib.hgs_binning_bytes(1 << 20, 3) >= (1 << 20) * 72
    assert lib.hgs_binning_bytes(0, 3) >= 0 and lib.hgs_sort_bytes(10) > 0 and lib.hgs_knn_bytes(100) > 0


def test_argument_validation_without_gpu():
    """Errors the reference raises on the host are raised before any device work (rasterize_points.cu:57-59,
    rasterizer_impl.cu:242-245); status codes and messages come back through hgs_last_error()."""
    lib = L.load()
    prm = L.RasterParams(P=10, D=0, M=0, width=64, height=64, channels=5, tan_fovx=1.0, tan_fovy=1.0,
                         scale_modifier=1.0, prefiltered=0, debug=0)
    inp = L.RasterInputs()
    st = lib.hgs_forward_stage_a(ctypes.byref(prm), ctypes.byref(inp), None, None, None)
    assert st == -1 and b"For non-RGB, provide precomputed Gaussian colors!" in lib.hgs_last_error()
    prm.channels = 3
    st = lib.hgs_forward_stage_a(ctypes.byref(prm), ctypes.byref(inp), None, None, None)
    assert st == -1 and b"missing required input" in lib.hgs_last_error()
    with pytest.raises(L.HgsError):
        L.check(st, "stage a")
    assert lib.hgs_mark_visible(-1, None, None, None, None, None) == -1
    assert lib.hgs_dist2_knn3(5, None, None, None, None) == -1
    assert lib.hgs_sort_pairs(10, 99, None, None, None, None, None, None) == -1


def test_python_surface_mirrors_reference():
    import diff_gaussian_rasterization as dgr
    fields = ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix",
              "projmatrix", "sh_degree", "campos", "prefiltered", "debug")
    assert dgr.GaussianRasterizationSettings._fields == fields
    import inspect
    sig = inspect.signature(dgr.GaussianRasterizer.forward)
    assert list(sig.parameters)[1:] == ["means3D", "means2D", "opacities", "shs", "colors_precomp", "scales",
                                        "rotations", "cov3D_precomp"]
    assert list(inspect.signature(dgr.rasterize_gaussians).parameters) == [
        "means3D", "means2D", "sh", "colors_precomp", "opacities", "scales", "rotations", "cov3Ds_precomp",
        "raster_settings"]
    assert len(inspect.signature(dgr._C.rasterize_gaussians).parameters) == 19
    assert len(inspect.signature(dgr._C.rasterize_gaussians_backward).parameters) == 21
    assert len(inspect.signature(dgr._C.mark_visible).parameters) == 3
    import simple_knn._C as knn
    assert callable(knn.distCUDA2)
    d = common.blob_inputs(10, 32, 32, "cpu")
    s = dgr.GaussianRasterizationSettings(32, 32, 1.0, 1.0, d["background"], 1.0, d["viewmatrix"], d["projmatrix"], 3,
                                          d["campos"], False, False)
    r = dgr.GaussianRasterizer(s)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(d["means3D"], None, d["opacity"], shs=d["sh"], colors_precomp=d["sh"][:, 0], scales=d["scales"],
          rotations=d["rotations"])
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
        r(d["means3D"], None, d["opacity"], shs=d["sh"])


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors / missing CUDA — never fall back."""
    import diff_gaussian_rasterization as dgr
    import simple_knn._C as knn
    d = common.blob_inputs(10, 32, 32, "cpu")
    with pytest.raises(L.HgsError, match="no CPU path"):
        dgr._C.rasterize_gaussians(*common.fwd_args(d))
    with pytest.raises(L.HgsError, match="no CPU path"):
        dgr._C.mark_visible(d["means3D"], d["viewmatrix"], d["projmatrix"])
    with pytest.raises(L.HgsError, match="no CPU path"):
        knn.distCUDA2(d["means3D"])
    src = ""
    pkg = os.path.join(ROOT, "hair-gs_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh")):
                src += open(os.path.join(dp, f)).read()
    assert "pyoracle" not in src and "rast_oracle" not in src and "oracle/" not in src.replace("oracle/_ref", "")


def test_scene_generators_are_deterministic_and_shaped():
    a, b = scenes.strand_scene(50, 21, seed=3), scenes.strand_scene(50, 21, seed=3)
    assert torch.equal(a.endpoints, b.endpoints) and a.endpoint_pairs.shape == (50 * 20, 2)
    assert a.endpoints.shape == (50 * 21, 3) and a.features_dc.shape == (1000, 1, 3)
    # consecutive segments share joints
    assert torch.equal(a.endpoint_pairs[0], torch.tensor([0, 1])) and torch.equal(a.endpoint_pairs[1], torch.tensor([1, 2]))
    c = scenes.blob_scene(100, seed=1)
    assert torch.allclose(c.rotations.norm(dim=1), torch.ones(100), atol=1e-6)
    cams = scenes.orbit_cameras(5, 64, 48)
    assert len(cams) == 5
    for cam in cams:
        # scene/cameras.py:93-108 conventions
        w2c = cam.world_view_transform.t()
        assert torch.allclose(w2c[:3, :3] @ w2c[:3, :3].t(), torch.eye(3), atol=1e-5)
        assert torch.allclose(cam.camera_center, -(w2c[:3, :3].t() @ w2c[:3, 3]), atol=1e-5)
        assert math.isclose(cam.tanfovx, 1.0, rel_tol=1e-6) and math.isclose(cam.tanfovy, 0.75, rel_tol=1e-6)
        # the look-at point projects to the image centre, in front of the camera
        p = torch.tensor([0.0, -0.05, 0.0, 1.0]) @ cam.full_proj_transform
        assert abs(float(p[0] / p[3])) < 1e-5 and abs(float(p[1] / p[3])) < 1e-5 and float(p[3]) > 0.2


def test_matrix_to_quaternion_round_trip():
    """pytorch3d's matrix_to_quaternion is not in the reference tree (parity unpinned): pin it by R(q) round trip
    using the rotation formula the rasterizer itself uses (forward.cu:133-138, transposed storage)."""
    g = torch.Generator().manual_seed(0)
    q = torch.randn(2000, 4, generator=g, dtype=torch.float64)
    q = q / q.norm(dim=1, keepdim=True)
    r, x, y, z = q.unbind(1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).view(-1, 3, 3)
    q2 = scenes.matrix_to_quaternion(R)
    q_std = torch.where(q[:, :1] < 0, -q, q)
    assert torch.allclose(q2, q_std, atol=1e-9)
    assert bool((q2[:, 0] >= 0).all())


def test_strand_parameterisation():
    """scene/hair_gaussian_model.py:134-201: mean = segment centre, scale_x = |e1-e0|/2 * k clamped, scale_yz = exp(w),
    the quaternion rotates +x onto the segment direction, collapsed segments get the identity."""
    sc = scenes.strand_scene(30, 11, seed=2)
    ep = sc.endpoints.double()
    ep[sc.endpoint_pairs[5, 1]] = ep[sc.endpoint_pairs[5, 0]]  # collapse one segment
    means, scales, rot, orient = scenes.strand_gaussians(ep, sc.endpoint_pairs, sc.width.double())
    e0, e1 = ep[sc.endpoint_pairs[:, 0]], ep[sc.endpoint_pairs[:, 1]]
    assert torch.allclose(means, (e0 + e1) / 2)
    dist = (e1 - e0).norm(dim=1)
    assert torch.allclose(scales[:, 0], torch.clamp(dist / 2 * scenes.DIST_TO_SCALE, min=1e-7))
    assert torch.allclose(scales[:, 1:], torch.exp(sc.width.double()).repeat(1, 2))
    assert torch.equal(rot[5], torch.tensor([1.0, 0, 0, 0], dtype=torch.float64))
    assert torch.equal(orient[5], torch.tensor([1.0, 0, 0], dtype=torch.float64))
    ok = dist > 1e-7
    r, x, y, z = rot[ok].unbind(1)
    first_col = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y + r * z), 2 * (x * z - r * y)], -1)  # R @ (1,0,0)
    assert torch.allclose(first_col, (e1 - e0)[ok] / dist[ok, None], atol=1e-6)
    assert torch.allclose(orient[ok], (e1 - e0)[ok] / dist[ok, None])
    assert torch.allclose(rot[ok].norm(dim=1), torch.ones(int(ok.sum()), dtype=torch.float64), atol=1e-6)


def test_fused_strand_entry_validation_without_gpu():
    lib = L.load()
    prm = L.RasterParams(P=10, D=0, M=1, width=64, height=64, channels=3, tan_fovx=1.0, tan_fovy=1.0,
                         scale_modifier=1.0, prefiltered=0, debug=0)
    inp = L.StrandInputs()
    assert lib.hgs_strands_forward_stage_a(ctypes.byref(prm), ctypes.byref(inp), None, None, None) == -1
    assert b"7 channels" in lib.hgs_last_error()
    prm.channels = 7
    assert lib.hgs_strands_forward_stage_a(ctypes.byref(prm), ctypes.byref(inp), None, None, None) == -1
    assert b"missing required strand input" in lib.hgs_last_error()
    from hairgs_b200 import fused, models
    sc = scenes.strand_scene(5, 6, seed=1)
    cam = scenes.orbit_cameras(2, 32, 32)[0]
    with pytest.raises(L.HgsError, match="no CPU path"):
        fused.render_strands(cam, models.StrandModel(sc), torch.zeros(7))
    assert lib.hgs_binning_capacity(lib.hgs_binning_bytes(12345, 3), 3, 12345) == 12345
    assert lib.hgs_binning_capacity(lib.hgs_binning_bytes(8 * 4096, 7), 7, 999) == 8 * 4096
    assert lib.hgs_binning_capacity(lib.hgs_binning_bytes(8 * 4096, 7) + 256, 7, 999) < 0


def test_graph_replay_host_logic_without_gpu():
    """hairgs_b200.graphs: no CPU path, and the deferred validation of a launch plan (pure host logic on the read-back
    words: num_rendered, -, overflow, depth_max bits, ~depth_min bits)."""
    from hairgs_b200 import fused, graphs, models
    sc = scenes.strand_scene(5, 6, seed=1)
    m = models.StrandModel(sc)
    with pytest.raises(L.HgsError, match="no CPU path"):
        graphs.GraphedStrandStep(m, None, torch.zeros(7), 32, 32, 1.0, 1.0, 4096, 32)

    def plan(capacity, bits, n, overflow, dmax, dmin, sort_mode=L.SORT_GLOBAL):
        p = graphs.LaunchPlan.__new__(graphs.LaunchPlan)
        p.capacity, p.depth_bits, p.sort_mode = capacity, bits, sort_mode
        p.host = torch.tensor([n, 0, overflow, dmax, ~dmin, 0, 0, 0], dtype=torch.int64).to(torch.int32)
        return p

    lo, hi = 0x3E4CCCCD, 0x3F4CCCCD          # float bits of 0.2 and 0.8: a 2^24 range -> 25 bits
    assert plan(1000, 25, 900, 0, hi, lo).check() == 900
    assert plan(1000, 32, 1000, 0, hi, lo).check() == 1000
    with pytest.raises(graphs.HgsPlanError, match="tile instances"):
        plan(1000, 25, 1001, 0, hi, lo).check()
    with pytest.raises(graphs.HgsPlanError, match="depth bits"):
        plan(1000, 24, 900, 0, hi, lo).check()
    with pytest.raises(graphs.HgsPlanError, match="int32"):
        plan(1000, 25, 900, 1, hi, lo).check()
    # in-tile sort plans: the depth range is irrelevant, a tile list longer than HGS_TILE_SORT_MAX (bit 2) is not
    assert plan(1000, 24, 900, 0, hi, lo, L.SORT_TILE).check() == 900
    with pytest.raises(graphs.HgsPlanError, match="HGS_TILE_SORT_MAX"):
        plan(1000, 32, 900, 4, hi, lo, L.SORT_TILE).check()
    assert plan(1000, 32, 900, 4, hi, lo, L.SORT_GLOBAL).check() == 900
    assert issubclass(graphs.HgsPlanError, L.HgsError)
    assert fused.GradSink({}).accumulate is False
