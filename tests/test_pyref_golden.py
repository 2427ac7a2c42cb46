"""Pins the torch-side rows of the hot path (merge search of §8f N4) against vectors recorded
from the reference's OWN Python (tests/golden/make_pyref_golden.py ran utils/sh.py, utils/transform.py, utils/general.py,
utils/graphics.py, scene/hair_gaussian_model.py through oracle/ref_python.py).  The CPU tests need neither the reference
nor a GPU; the GPU tests run the product kernels / render() switches against the same vectors and, where oracle/_ref/pyref
travelled to the box, against the reference's Python executed live."""
import math
import os

import numpy as np
import pytest
import torch

import common  # noqa: F401  (sys.path set-up)
from hairgs_b200 import scenes, sh

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pyref.npz")


@pytest.fixture(scope="module")
def z():
    return np.load(GOLD)


def t(a):
    return torch.tensor(np.asarray(a))


def test_eval_sh_matches_reference(z):
    """A23: hairgs_b200.sh.eval_sh vs utils/sh.py:55-118 — same expression order, so bit-equal on the same torch build;
    tolerance 1 ulp-ish (2e-7 relative to the O(1) result) to stay robust across torch versions."""
    for deg in range(4):
        got = sh.eval_sh(deg, t(z["sh_in"]), t(z["sh_dirs"])).numpy()
        assert np.abs(got - z[f"sh_out_deg{deg}"]).max() <= 2e-7, deg
    assert np.array_equal(sh.RGB2SH(t(z["rgb_in"])).numpy(), z["rgb2sh"])
    assert np.array_equal(sh.SH2RGB(t(z["rgb_in"])).numpy(), z["sh2rgb"])


def test_covariance_helpers_match_reference(z):
    """A23: build_rotation / build_scaling_rotation / strip_symmetric / covariance vs utils/transform.py:7-42,
    utils/general.py:71-84, scene/gaussian_model.py:61-65."""
    s, r = t(z["cov_scales"]), t(z["cov_rots"])
    assert np.abs(sh.build_rotation(r).numpy() - z["build_rotation"]).max() <= 3e-7
    ref = z["build_scaling_rotation"]
    assert np.abs(sh.build_scaling_rotation(s, r).numpy() - ref).max() <= 3e-7 * np.abs(ref).max()
    L = sh.build_scaling_rotation(s, r)
    ref = z["strip_symmetric"]
    assert np.abs(sh.strip_symmetric(L @ L.transpose(1, 2)).numpy() - ref).max() <= 1e-6 * np.abs(ref).max()
    ref = z["covariance_mod0.7"]
    assert np.abs(sh.build_covariance_from_scaling_rotation(s, 0.7, r).numpy() - ref).max() <= 1e-6 * np.abs(ref).max()


def test_strand_getters_match_reference(z):
    """A21: scenes.strand_* / models.StrandModel vs HairGaussianModel's properties (scene/hair_gaussian_model.py:134-206,
    utils/transform.py:54-86), including collapsed segments and a segment pointing along -x.  The reference selects with
    boolean masks, ours with torch.where: identical values required."""
    from hairgs_b200 import models
    e, p, w = t(z["strand_endpoints"]), t(z["strand_pairs"]), t(z["strand_width"])
    assert np.array_equal(scenes.strand_xyz(e, p).numpy(), z["strand_xyz"])
    assert np.array_equal(scenes.strand_scaling(e, p, w).numpy(), z["strand_scaling"])
    assert np.array_equal(scenes.strand_orientation(e, p).numpy(), z["strand_orientation"])
    q, qr = scenes.strand_rotation(e, p).numpy(), z["strand_rotation"]
    collapsed = np.linalg.norm(z["strand_endpoints"][z["strand_pairs"][:, 1]] - z["strand_endpoints"][z["strand_pairs"][:, 0]], axis=1) <= 1e-7
    assert collapsed.sum() >= 1 and np.array_equal(qr[collapsed], np.tile([1.0, 0, 0, 0], (collapsed.sum(), 1)))
    assert np.abs(q - qr).max() <= 1e-6
    sc = scenes.StrandScene(e, p, w, t(z["strand_opacity_logit"]), t(z["strand_mask_logit"]),
                            t(z["strand_features"])[:, :1], t(z["strand_features"])[:, 1:], 40)
    m = models.StrandModel(sc)
    with torch.no_grad():
        for mod, key in ((0.5, "strand_covariance_default"), (1.0, "strand_covariance_1.0")):
            ref = z[key]
            assert np.abs(m.get_covariance(mod).numpy() - ref).max() <= 2e-6 * np.abs(ref).max(), key
        assert np.array_equal(m.get_opacity.numpy(), z["strand_opacity"])
        assert np.array_equal(m.get_mask.numpy(), z["strand_mask"])
        assert np.array_equal(m.get_features.numpy(), z["strand_features"])


def test_camera_matrices_match_reference(z):
    """A24: scenes.camera_from_w2c vs getWorld2View2 / getProjectionMatrix (utils/graphics.py:38-71) composed as
    scene/cameras.py:93-108."""
    w2c = np.eye(4)
    w2c[:3, :3], w2c[:3, 3] = z["cam_R"].T, z["cam_T"]
    fovx, fovy = z["cam_fov"]
    cam = scenes.camera_from_w2c(w2c, 640, 480, fovx, fovy)
    wv_ref = torch.tensor(z["getWorld2View2"]).transpose(0, 1)
    assert np.array_equal(cam.world_view_transform.numpy(), wv_ref.numpy())
    proj_ref = torch.tensor(z["getProjectionMatrix"]).transpose(0, 1)
    full_ref = wv_ref.unsqueeze(0).bmm(proj_ref.unsqueeze(0)).squeeze(0)
    assert np.array_equal(cam.full_proj_transform.numpy(), full_ref.numpy())
    assert np.array_equal(cam.camera_center.numpy(), wv_ref.inverse()[3, :3].numpy())


def merge_inputs(z, tag):
    """The bookkeeping compute_endpoint_pair_to_merge does before its search (scene/hair_gaussian_model.py:1261-1281):
    strand ends = endpoint ids that appear once, restricted to foreground segments; direction = end -> neighbouring joint."""
    e, p = z[f"merge_{tag}_endpoints"], z[f"merge_{tag}_pairs"]
    fg = np.logical_and(z[f"merge_{tag}_opacity"][:, 0] >= 0.005, z[f"merge_{tag}_mask"][:, 0] >= 0.25)  # gaussian_model.py:727-733
    ids, counts = np.unique(p, return_counts=True)
    ends = ids[counts == 1]
    ends = ends[np.isin(ends, p[fg].reshape(-1))]
    neighbour = {}
    for a, b in p:
        neighbour.setdefault(int(a), int(b))
        neighbour.setdefault(int(b), int(a))
    nb = np.array([neighbour[int(i)] for i in ends])
    d = e[nb] - e[ends]
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    comp = z[f"merge_{tag}_complementary"]
    dist_th, angle, bidir, max_nn = z[f"merge_{tag}_cfg"]
    return e[ends], d.astype(np.float32), ends.astype(np.int64), comp[ends], comp, float(dist_th), float(angle), bool(bidir), int(max_nn)


def same_pairs(got, ref):
    """Row-for-row equality of the merge list up to the ORIENTATION of a row.  Every candidate (a, b) has a mirror (b, a)
    at exactly the same distance; the reference orders them with an unstable torch.sort (:1339), so which of the two
    survives remove_duplicate_endpoint_rows is an artefact of the sort implementation (it differs between torch's CPU and
    CUDA sorts).  The set of merged end pairs and their order by distance are the contract; ours keeps search order
    (stable sort)."""
    return got.shape == ref.shape and np.array_equal(np.sort(got, axis=1), np.sort(ref, axis=1))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_merge_oracle_matches_reference(z, tag):
    """Pins oracle/merge_oracle.py (the checker of the device merge search) on what the reference's own
    compute_endpoint_pair_to_merge returned for three scenes (plain, bidirectional, max_num_nn=2)."""
    from oracle import merge_oracle
    pts, dirs, gid, other, comp, dist_th, angle, bidir, max_nn = merge_inputs(z, tag)
    p1, p2, d = merge_oracle.merge_candidates(pts, dirs, gid, other, dist_th, angle, bidir, max_nn)
    order = np.argsort(d, kind="stable")
    keep = merge_oracle.greedy_filter(p1[order], p2[order], comp)
    got = np.stack([p1[order][keep], p2[order][keep]], 1)
    ref = z[f"merge_{tag}_result"]
    assert ref.shape[0] >= 20
    assert same_pairs(got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_merge_search_matches_reference_golden(z, tag):
    """N4: the device merge search (hgs_merge_count/_fill/_greedy) returns exactly the pairs the reference returned."""
    from hairgs_b200 import merge
    dev = torch.device("cuda:0")
    pts, dirs, gid, other, comp, dist_th, angle, bidir, max_nn = merge_inputs(z, tag)
    g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    pairs = merge.endpoint_pairs_to_merge(g(pts), g(dirs), g(gid), g(other), g(comp), dist_th, angle, bidir, max_nn)
    assert same_pairs(pairs.cpu().numpy(), z[f"merge_{tag}_result"])


@pytest.mark.gpu
def test_render_python_switches_match_default_path():
    """A1/A23: render(convert_SHs_python=True) and render(compute_cov3D_python=True) (gaussian_renderer/__init__.py:82-104)
    against the default in-kernel path: radii equal, pixels <= 1e-4, parameter gradients rel <= 1e-3."""
    from gaussian_renderer import render
    from hairgs_b200 import models
    dev = torch.device("cuda:0")
    sc = scenes.blob_scene(20000, seed=3, sh_coeffs=16).to(dev)
    cam = scenes.orbit_cameras(4, 320, 256, device=dev)[1]
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    torch.manual_seed(0)
    w = torch.randn(3, 256, 320, device=dev)
    outs = []
    for kw in ({}, {"convert_SHs_python": True}, {"compute_cov3D_python": True}):
        m = models.BlobModel(sc, sh_degree=3).to(dev)
        r = render(cam, m, bg, scaling_modifier=0.9, **kw)
        (r["render"] * w).sum().backward()
        outs.append((r["render"].detach(), r["radii"], {n: p.grad.clone() for n, p in m.named_parameters()}))
    (c0, r0, g0) = outs[0]
    for (c, r, g), name in zip(outs[1:], ("convert_SHs_python", "compute_cov3D_python")):
        assert torch.equal(r, r0), name
        assert float((c - c0).abs().max()) <= 1e-4, (name, float((c - c0).abs().max()))
        for n in g0:
            # the in-kernel covariance path returns dL/d(scale_modifier * scale) as "dL_dscales" (backward_distwar.cu:296-326
            # never multiplies by the modifier); the torch covariance differentiates through the multiplication
            k = 0.9 if (name == "compute_cov3D_python" and n == "_scaling") else 1.0
            assert common.rel_err(g[n], k * g0[n]) <= 1e-3, (name, n, common.rel_err(g[n], k * g0[n]))


@pytest.mark.gpu
def test_strand_getters_match_reference_live_on_gpu():
    """A21 on the device, against the reference's HairGaussianModel executed live (byte-compiled reference modules under
    oracle/_ref/pyref; skipped where they did not travel)."""
    from oracle import ref_python as rp
    if not rp.available():
        pytest.skip("oracle/_ref/pyref not built")
    ns = rp.load(with_cuda_ext=False)
    dev = torch.device("cuda:0")
    sc = scenes.strand_scene(500, 50, seed=9).to(dev)
    sc.endpoints[sc.endpoint_pairs[11, 1]] = sc.endpoints[sc.endpoint_pairs[11, 0]]
    hm = ns.hair_gaussian_model.HairGaussianModel(0, device="cuda")
    hm._endpoints, hm.endpoint_pairs, hm._width = sc.endpoints, sc.endpoint_pairs, sc.width
    assert torch.equal(scenes.strand_xyz(sc.endpoints, sc.endpoint_pairs), hm.get_xyz)
    assert torch.equal(scenes.strand_scaling(sc.endpoints, sc.endpoint_pairs, sc.width), hm.get_scaling)
    assert torch.equal(scenes.strand_orientation(sc.endpoints, sc.endpoint_pairs), hm.get_orientation)
    assert float((scenes.strand_rotation(sc.endpoints, sc.endpoint_pairs) - hm.get_rotation).abs().max()) <= 1e-6
