"""Generates the golden fixtures in tests/golden/*.npz by running the UNMODIFIED reference build
(oracle/_ref, compiled from /root/reference by oracle/build_ref.py) on a B200:

    gpurun -- python tests/golden/make_golden.py gpurun_out/golden     # then copy *.npz into tests/golden/

Each file holds the exact inputs (so nothing depends on regenerating them bit for bit), every intermediate
buffer of the reference forward pass (unpacked from geomBuffer / binningBuffer / imgBuffer), the image, and
the gradients of the reference backward for a stored dL_dout.  The reference tree itself ships no fixtures
(SURVEY.md §4); these pin the CPU oracle and, through it, the CPU-side tests.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "hair-gs_b200"), os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import common  # noqa: E402
import refload  # noqa: E402


def cases(dev):
    yield "blobs_sh3", common.blob_inputs(2000, 128, 100, dev, seed=11)
    yield "blobs_sh1", common.blob_inputs(1500, 96, 96, dev, seed=12, sh_degree=1, M=4, view=1)
    d = common.blob_inputs(600, 80, 64, dev, seed=13, scale_mul=6.0, view=2)
    g = torch.Generator().manual_seed(5)
    d["colors"], d["sh"] = torch.rand(600, 3, generator=g).to(dev), common.EMPTY()
    # push some points behind / near the camera to exercise the cull and the 1.3*tanfov clamp
    d["means3D"] = d["means3D"].clone()
    d["means3D"][:60] += (d["campos"] - d["means3D"][:60]) * torch.linspace(0.5, 1.3, 60, device=dev)[:, None]
    yield "blobs_colors_big", d
    yield "strands_rgb", common.strand_inputs(40, 26, 160, 120, dev, seed=14, view=3)
    yield "strands_orient", common.strand_inputs(40, 26, 160, 120, dev, seed=14, view=0, colors="orientation")
    # precomputed covariance path
    d = common.blob_inputs(800, 64, 64, dev, seed=15)
    s, q = d["scales"], d["rotations"]
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).view(-1, 3, 3)
    Lm = R * s[:, None, :]
    Sig = Lm @ Lm.transpose(1, 2)
    d["cov3D_precomp"] = torch.stack([Sig[:, 0, 0], Sig[:, 0, 1], Sig[:, 0, 2], Sig[:, 1, 1], Sig[:, 1, 2], Sig[:, 2, 2]], -1).contiguous()
    d["scales"], d["rotations"] = common.EMPTY(), common.EMPTY()
    yield "blobs_cov_precomp", d


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    C = refload.ref_dgr()
    assert C is not None, "oracle/_ref is not built"
    for name, d in cases(dev):
        N, color, radii, bufs, views = common.ref_forward(d)
        rng = np.random.default_rng(abs(hash(name)) % 1000 + 7)
        dL = torch.tensor(rng.standard_normal(tuple(color.shape)).astype(np.float32), device=dev)
        grads = C.rasterize_gaussians_backward(*common.bwd_args(d, radii, dL, bufs[0], N, bufs[1], bufs[2]))
        torch.cuda.synchronize()
        rec = {"in_" + k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}
        rec.update({"out_" + k: v.cpu().numpy() for k, v in views.items()})
        rec.update(out_num_rendered=np.int64(N), out_color=color.cpu().numpy(), out_radii=radii.cpu().numpy(),
                   in_dL_dout=dL.cpu().numpy())
        for gname, g in zip(("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh",
                             "dL_dscales", "dL_drotations"), grads):
            rec["grad_" + gname] = g.cpu().numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
        print(name, "P", d["means3D"].shape[0], "N", N, "visible", int((radii > 0).sum()))
    K = refload.ref_knn()
    rng = np.random.default_rng(3)
    pts = np.concatenate([rng.standard_normal((3000, 3)) * 0.2, rng.uniform(-1, 1, (1000, 3)),
                          np.repeat(rng.standard_normal((20, 3)), 3, 0)]).astype(np.float32)  # incl. exact duplicates
    out = K.distCUDA2(torch.tensor(pts, device=dev)).cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "knn.npz"), in_points=pts, out_dist2=out)
    print("knn", pts.shape[0])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
