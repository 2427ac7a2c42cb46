"""First-contact GPU diagnostics: run ours and the reference build on the same inputs and print
per-buffer mismatch statistics (not a test; see tests/ for the asserted version)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "hair-gs_b200"), os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)

import torch  # noqa: E402

import common  # noqa: E402
import refload  # noqa: E402
import diff_gaussian_rasterization._C as ours_C  # noqa: E402


def report(name, a, b, mask=None):
    if mask is not None:
        a, b = a[mask], b[mask]
    if a.shape != b.shape:
        print(f"  {name:18s} SHAPE MISMATCH {tuple(a.shape)} vs {tuple(b.shape)}")
        return
    nb = common.bits_equal(a, b)
    if a.dtype == torch.float32 and a.numel():
        d = (a - b).abs()
        finite = torch.isfinite(d)
        mx = d[finite].max().item() if finite.any() else float("nan")
        print(f"  {name:18s} n={a.numel():9d} bit-mismatch={nb:8d} max-abs={mx:.3e}")
    else:
        print(f"  {name:18s} n={a.numel():9d} mismatch={nb:8d}")
    return nb


def run_case(title, d, bwd=True):
    print(f"== {title}: P={d['means3D'].shape[0]} {d['image_width']}x{d['image_height']} D={d['degree']}")
    torch.cuda.synchronize()
    No, co, ro, bo, vo = common.ours_forward(d)
    torch.cuda.synchronize()
    Nr, cr, rr, br, vr = common.ref_forward(d)
    torch.cuda.synchronize()
    print(f"  num_rendered ours={No} ref={Nr}  visible ours={(ro > 0).sum().item()} ref={(rr > 0).sum().item()}")
    report("radii", ro, rr)
    vis = (rr > 0) & (ro > 0)
    for k in ("tiles_touched", "point_offsets"):
        report(k, vo[k], vr[k])
    for k in ("depths", "means2D", "conic_opacity", "rgb", "cov3D", "clamped"):
        if k in vo and k in vr:
            if k == "rgb" and d["colors"].numel():
                continue
            report(k, vo[k], vr[k], vis)
    if No == Nr:
        report("sorted keys", vo["point_list_keys"], vr["point_list_keys"])
        report("point_list", vo["point_list"], vr["point_list"])
    report("ranges", vo["ranges"], vr["ranges"])
    report("n_contrib", vo["n_contrib"], vr["n_contrib"])
    report("final_T", vo["accum_alpha"], vr["accum_alpha"])
    report("out_color", co, cr)
    if not bwd:
        return
    torch.manual_seed(1)
    dL = torch.randn_like(cr)
    go = ours_C.rasterize_gaussians_backward(*common.bwd_args(d, ro, dL, bo[0], No, bo[1], bo[2]))
    gr = refload.ref_dgr().rasterize_gaussians_backward(*common.bwd_args(d, rr, dL, br[0], Nr, br[1], br[2]))
    torch.cuda.synchronize()
    names = ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
             "dL_drotations")
    for n, a, b in zip(names, go, gr):
        if a.numel() == 0:
            continue
        fin = torch.isfinite(a).all().item() and torch.isfinite(b).all().item()
        print(f"  {n:14s} rel={common.rel_err(a, b):.3e} max|ref|={b.abs().max().item():.3e} finite={fin}")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n


def main():
    dev = torch.device("cuda:0")
    print(torch.cuda.get_device_name(0))
    assert refload.ref_dgr() is not None, "oracle/_ref not built"
    small = "--small" in sys.argv
    import simple_knn._C as knn_ours
    K = refload.ref_knn()
    for n in (1, 2, 3, 4, 1000, 100000):
        pts = torch.randn(n, 3, device=dev) * torch.tensor([1.0, 0.2, 3.0], device=dev) + 2.0
        a, b = knn_ours.distCUDA2(pts), K.distCUDA2(pts)
        torch.cuda.synchronize()
        fin = torch.isfinite(b)
        print(f"knn n={n}: bit-mismatch={common.bits_equal(a, b)} max-rel={(((a - b).abs() / b.abs().clamp_min(1e-30))[fin].max().item() if fin.any() else 0):.3e}"
              f" inf ours={int((~torch.isfinite(a)).sum())} ref={int((~fin).sum())}")
    t_ref = timeit(lambda: K.distCUDA2(pts))
    t_our = timeit(lambda: knn_ours.distCUDA2(pts))
    print(f"knn 100k ms: ref={t_ref:.3f} ours={t_our:.3f}")
    run_case("blobs sh3", common.blob_inputs(20000 if small else 300000, 512, 512, dev))
    run_case("blobs big splats", common.blob_inputs(5000, 256, 200, dev, scale_mul=8.0, seed=3))
    run_case("strands rgb", common.strand_inputs(200 if small else 2000, 100, 1024, 1024, dev))
    run_case("strands orientation", common.strand_inputs(200 if small else 2000, 100, 1024, 1024, dev, colors="orientation"))
    if not small:
        d = common.strand_inputs(10000, 100, 1024, 1024, dev)
        run_case("cfg3 strands 1M", d)
        C = refload.ref_dgr()
        t_ref = timeit(lambda: C.rasterize_gaussians(*common.fwd_args(d)))
        t_our = timeit(lambda: ours_C.rasterize_gaussians(*common.fwd_args(d)))
        print(f"forward ms: ref={t_ref:.3f} ours={t_our:.3f}")
        No, co, ro, geo, bino, imgo = ours_C.rasterize_gaussians(*common.fwd_args(d))
        Nr, cr, rr, ger, binr, imgr = C.rasterize_gaussians(*common.fwd_args(d))
        dL = torch.randn_like(cr)
        t_ref = timeit(lambda: C.rasterize_gaussians_backward(*common.bwd_args(d, rr, dL, ger, Nr, binr, imgr)))
        t_our = timeit(lambda: ours_C.rasterize_gaussians_backward(*common.bwd_args(d, ro, dL, geo, No, bino, imgo)))
        print(f"backward ms: ref={t_ref:.3f} ours={t_our:.3f}")


if __name__ == "__main__":
    main()
