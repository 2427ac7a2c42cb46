"""Golden vectors for the torch-side rows of the hot path (SURVEY §8a A21, A23, A24 and the merge search of §8f N4),
recorded by RUNNING THE REFERENCE'S OWN PYTHON in the build container (CPU):

  utils/sh.py:55-126            eval_sh (degrees 0-3), RGB2SH, SH2RGB
  utils/transform.py:7-42       build_rotation, build_scaling_rotation
  utils/general.py:71-84        strip_symmetric (covariance 6-vector)
  scene/gaussian_model.py:61-65 build_covariance_from_scaling_rotation
  scene/hair_gaussian_model.py:134-206   HairGaussianModel.get_xyz / get_scaling / get_rotation / get_orientation /
                                         get_covariance (incl. a collapsed segment)
  utils/transform.py:54-86      calculate_rotation_from_vectors (through get_rotation)
  utils/graphics.py:38-71       getWorld2View2, getProjectionMatrix
  scene/hair_gaussian_model.py:1205-1362, 1410-1500   compute_strands_info + compute_endpoint_pair_to_merge

The modules are the byte-compiled reference files of oracle/ref_python.py.  The reference hard-codes device="cuda" in
build_rotation / strip_lowerdiag; with no GPU here those factory calls are redirected to the CPU for the duration of the
call (ref_python.cuda_as_cpu).  pytorch3d.transforms.matrix_to_quaternion is absent (un-vendored, un-pinned) and is the
restatement in hairgs_b200/scenes.py — so the quaternions in this file pin everything AROUND that function (R, masks,
order of operations); the function itself stays pinned by the R(q) round trip in tests/test_host_cpu.py.

    python tests/golden/make_pyref_golden.py        # writes tests/golden/pyref.npz (needs /root/reference)
"""
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "hair-gs_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def merge_scene(seed, S=120, V=6, extent=0.012):
    """S short strands in a small box so that many strand ends lie within the merge radius of each other."""
    rng = np.random.default_rng(seed)
    roots = rng.uniform(0, extent, (S, 3))
    d = rng.normal(size=(S, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    t = np.linspace(0, 0.004, V)[None, :, None]
    verts = roots[:, None, :] + t * d[:, None, :] + rng.normal(scale=1e-4, size=(S, V, 3))
    endpoints = verts.reshape(S * V, 3).astype(np.float32)
    base = (np.arange(S) * V)[:, None] + np.arange(V - 1)[None, :]
    pairs = np.stack([base, base + 1], -1).reshape(-1, 2).astype(np.int64)
    P = pairs.shape[0]
    opacity = rng.uniform(0.2, 0.9, (P, 1)).astype(np.float32)
    mask = rng.uniform(0.5, 0.99, (P, 1)).astype(np.float32)
    return endpoints, pairs, opacity, mask, roots.astype(np.float32)


def main():
    from oracle import ref_python as rp
    from hairgs_b200 import scenes
    ns = rp.load(with_cuda_ext=False)
    g = torch.Generator().manual_seed(20261018)
    out = {}

    # ---- A23: SH and covariance helpers ---------------------------------------------------------------------------
    N = 257
    sh = torch.randn(N, 3, 16, generator=g) * 0.4
    dirs = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=1)
    out["sh_in"], out["sh_dirs"] = sh.numpy(), dirs.numpy()
    for deg in range(4):
        out[f"sh_out_deg{deg}"] = ns.sh.eval_sh(deg, sh, dirs).numpy()
    rgb = torch.rand(N, 3, generator=g)
    out["rgb_in"], out["rgb2sh"], out["sh2rgb"] = rgb.numpy(), ns.sh.RGB2SH(rgb).numpy(), ns.sh.SH2RGB(rgb).numpy()
    scales = torch.exp(torch.randn(N, 3, generator=g) * 0.5 - 5.0)
    rots = torch.randn(N, 4, generator=g)        # un-normalised on purpose: build_rotation normalises
    with rp.cuda_as_cpu():
        out["cov_scales"], out["cov_rots"] = scales.numpy(), rots.numpy()
        out["build_rotation"] = ns.transform.build_rotation(rots).numpy()
        out["build_scaling_rotation"] = ns.transform.build_scaling_rotation(scales, rots).numpy()
        gm = ns.gaussian_model.GaussianModel(3, device="cpu")
        out["covariance_mod0.7"] = gm.covariance_activation(scales, 0.7, rots).numpy()
        L = ns.transform.build_scaling_rotation(scales, rots)
        out["strip_symmetric"] = ns.general.strip_symmetric(L @ L.transpose(1, 2)).numpy()

    # ---- A21: strand-aligned parameterisation ---------------------------------------------------------------------
    sc = scenes.strand_scene(40, 12, seed=33)
    sc.endpoints[sc.endpoint_pairs[7, 1]] = sc.endpoints[sc.endpoint_pairs[7, 0]]      # collapsed segment(s)
    sc.endpoints[sc.endpoint_pairs[100, 1]] = sc.endpoints[sc.endpoint_pairs[100, 0]] + torch.tensor([-2e-3, 0.0, 1e-9])  # ~ -x
    hm = ns.hair_gaussian_model.HairGaussianModel(0, device="cpu")
    hm._endpoints, hm.endpoint_pairs, hm._width = sc.endpoints.clone(), sc.endpoint_pairs.clone(), sc.width.clone()
    hm._opacity, hm._mask = sc.opacity_logit.clone(), sc.mask_logit.clone()
    hm._features_dc, hm._features_rest = sc.features_dc.clone(), sc.features_rest.clone()
    out["strand_endpoints"], out["strand_pairs"], out["strand_width"] = sc.endpoints.numpy(), sc.endpoint_pairs.numpy(), sc.width.numpy()
    out["strand_opacity_logit"], out["strand_mask_logit"] = sc.opacity_logit.numpy(), sc.mask_logit.numpy()
    with rp.cuda_as_cpu():
        out["strand_xyz"] = hm.get_xyz.numpy()
        out["strand_scaling"] = hm.get_scaling.numpy()
        out["strand_rotation"] = hm.get_rotation.numpy()
        out["strand_orientation"] = hm.get_orientation.numpy()
        out["strand_covariance_default"] = hm.get_covariance().numpy()
        out["strand_covariance_1.0"] = hm.get_covariance(1.0).numpy()
        out["strand_opacity"], out["strand_mask"] = hm.get_opacity.numpy(), hm.get_mask.numpy()
        out["strand_features"] = hm.get_features.numpy()

    # ---- A24: camera matrices -------------------------------------------------------------------------------------
    cam = scenes.orbit_cameras(5, 640, 480)[2]
    w2c = cam.world_view_transform.t().numpy().astype(np.float64)
    R, T = w2c[:3, :3].T.copy(), w2c[:3, 3].copy()      # the loaders hand the transposed rotation (scene/cameras.py:93)
    out["cam_R"], out["cam_T"] = R, T
    out["cam_fov"] = np.array([cam.FoVx, cam.FoVy])
    out["getWorld2View2"] = ns.graphics.getWorld2View2(R, T, np.array([0.0, 0.0, 0.0]), 1.0)
    out["getProjectionMatrix"] = ns.graphics.getProjectionMatrix(znear=0.01, zfar=100.0, fovX=cam.FoVx, fovY=cam.FoVy).numpy()

    # ---- N4: merge search -----------------------------------------------------------------------------------------
    for tag, seed, bidir, max_nn in (("a", 5, False, -1), ("b", 6, True, -1), ("c", 7, False, 2)):
        endpoints, pairs, opacity, mask, roots = merge_scene(seed)
        m = ns.hair_gaussian_model.HairGaussianModel(0, device="cpu")
        m._endpoints, m.endpoint_pairs = torch.tensor(endpoints), torch.tensor(pairs)
        m._opacity, m._mask = torch.logit(torch.tensor(opacity)), torch.logit(torch.tensor(mask))
        m.ref_strand_root = roots
        m.merge_dist_th, m.merge_angle_th = 2.5e-3, 50.0
        m.training_args = types.SimpleNamespace(bidirectional_merge=bidir)
        m.compute_strands_info()
        res = m.compute_endpoint_pair_to_merge(max_num_nn=max_nn)
        out[f"merge_{tag}_endpoints"], out[f"merge_{tag}_pairs"] = endpoints, pairs
        out[f"merge_{tag}_opacity"], out[f"merge_{tag}_mask"] = opacity, mask
        out[f"merge_{tag}_cfg"] = np.array([2.5e-3, 50.0, float(bidir), float(max_nn)])
        out[f"merge_{tag}_complementary"] = m.strands_info.strand_endpoint_id_to_complementary.astype(np.int64)
        out[f"merge_{tag}_result"] = res.numpy().astype(np.int64)
        print(f"merge_{tag}: {res.shape[0]} pairs to merge")
    np.savez_compressed(os.path.join(HERE, "pyref.npz"), **out)
    print("wrote pyref.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
