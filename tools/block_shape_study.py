"""CPU model of the compositors' cull granularity and candidate-queue policy (no GPU needed).

    python tools/block_shape_study.py [view] [n_tiles]

Runs the C oracle on one cfg3 view (990 k strand Gaussians, 1024x1024), recomputes the alpha >= 1/255 extents of
`alpha_extent` (hair-gs_b200/csrc/preprocess.cu) and counts, for a sample of tiles,
  * the candidates of pixel blocks of different shapes (the cull test of composite_warp.cu), and
  * the trips of the blend / recurrence loops under the queue policies of the 8x4 kernels (one block per warp,
    two candidates per trip) and of the half-warp kernels (two 4x4 blocks per warp with one candidate ring each).
This is the evidence behind composite_fwd_half_kernel / composite_bwd_half_kernel; measured times are in profiles/.
TEST/ANALYSIS INFRASTRUCTURE: imports oracle/, never imported by the product.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hair-gs_b200"), os.path.join(ROOT, "tests")]
import common  # noqa: E402
from oracle import pyoracle  # noqa: E402


def ceil2(v):
    return (v + 1) // 2 * 2


def main():
    view = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n_tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    W = H = 1024
    d = common.strand_inputs(10000, 100, W, H, "cpu", seed=0, view=view, n_views=16)
    f = pyoracle.Forward({k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in d.items()})
    m2, co, pl = f.array("means2D"), f.array("conic_opacity"), f.array("point_list")
    keys, ranges = f.array("point_list_keys"), f.array("ranges")
    ncontrib = f.array("n_contrib").reshape(H, W)
    A, B, C, o = [co[:, i].astype(np.float64) for i in range(4)]
    det = A * C - B * B
    t = 2 * np.log(255 * np.maximum(o, 1 / 255)) * 1.002 + 1e-3
    with np.errstate(all="ignore"):
        hx = np.sqrt(t * C / det) * 1.001 + 1e-3
        hy = np.sqrt(t * A / det) * 1.001 + 1e-3
    print(f"N = {f.N}, median half extent of the alpha >= 1/255 region: {np.median(hx[o > 1 / 255]):.2f} x "
          f"{np.median(hy[o > 1 / 255]):.2f} px")
    tile = (keys >> 32).astype(np.int64)
    gx = W // 16
    tx0, ty0 = (tile % gx) * 16, (tile // gx) * 16
    x, y, ex, ey = m2[pl, 0].astype(np.float64), m2[pl, 1].astype(np.float64), hx[pl], hy[pl]

    print("candidates (instance x block pairs, whole lists) per block shape:")
    for bw, bh in ((8, 4), (4, 8), (16, 2), (4, 4), (8, 8), (16, 16)):
        tot = 0
        for by in range(0, 16, bh):
            for bx in range(0, 16, bw):
                wx0, wy0 = tx0 + bx, ty0 + by
                tot += int((~((x - ex > wx0 + bw - 1) | (x + ex < wx0) | (y - ey > wy0 + bh - 1) | (y + ey < wy0))).sum())
        print(f"  {bw:2d}x{bh:<2d}: {tot / 1e6:6.2f} M pairs, {tot * bw * bh / 1e6:7.1f} M lane evaluations")

    rs = np.random.default_rng(0)
    nonempty = np.nonzero(ranges[:, 1] > ranges[:, 0])[0]
    sel = rs.choice(nonempty, min(n_tiles, len(nonempty)), replace=False)
    tot = dict(fwd_8x4=0, fwd_4x4_per_chunk=0, fwd_4x4_rings=0, bwd_8x4=0, bwd_4x4_rings=0, lower_bound=0,
               cand_8x4=0, cand_4x4=0)
    QN = 40
    for tl in sel:
        r0, r1 = ranges[tl]
        X0, Y0 = (tl % gx) * 16, (tl // gx) * 16
        xs, ys, exs, eys = x[r0:r1], y[r0:r1], ex[r0:r1], ey[r0:r1]
        for w in range(8):
            bx, by = X0 + (w & 1) * 8, Y0 + (w >> 1) * 4
            L = int(ncontrib[by:by + 4, bx:bx + 8].max())  # positions the walks visit (early termination)
            if L == 0:
                continue
            inrow = ~((ys - eys > by + 3) | (ys + eys < by))
            ca = (inrow & ~((xs - exs > bx + 3) | (xs + exs < bx)))[:L]
            cb = (inrow & ~((xs - exs > bx + 7) | (xs + exs < bx + 4)))[:L]
            c8 = ca | cb
            nch = (L + 31) // 32
            pad = nch * 32 - L
            ca_, cb_, c8_ = (np.pad(v, (0, pad)).reshape(nch, 32).sum(1) for v in (ca, cb, c8))
            tot["cand_8x4"] += int(c8_.sum())
            tot["cand_4x4"] += int(ca_.sum() + cb_.sum())
            tot["lower_bound"] += int(max(ca_.sum(), cb_.sum()))
            tot["fwd_8x4"] += int(ceil2(c8_).sum())
            tot["fwd_4x4_per_chunk"] += int(ceil2(np.maximum(ca_, cb_)).sum())
            # forward rings: drain in step while both halves hold a pair, one-sided only to make room
            qa = qb = trips = 0
            for i in range(nch):
                qa += ca_[i]
                qb += cb_[i]
                n = min(qa, qb) // 2 * 2
                trips += n
                qa -= n
                qb -= n
                hi = max(qa, qb)
                if hi > QN - 32:
                    over = ceil2(hi - (QN - 32))
                    trips += over
                    qa -= min(qa, over)
                    qb -= min(qb, over)
            tot["fwd_4x4_rings"] += int(trips + ceil2(max(qa, qb)))
            # backward: groups of 8 (per half), two candidates per trip
            n8 = int(c8_.sum())
            tot["bwd_8x4"] += (n8 // 8) * 8 + ceil2(n8 % 8)
            qa = qb = trips = 0
            for i in range(nch - 1, -1, -1):
                qa += ca_[i]
                qb += cb_[i]
                while (qa >= 8 and qb >= 8) or qa > QN - 32 or qb > QN - 32:
                    na, nb = min(8, qa), min(8, qb)
                    trips += ceil2(max(na, nb))
                    qa -= na
                    qb -= nb
            while qa > 0 or qb > 0:
                na, nb = min(8, qa), min(8, qb)
                trips += ceil2(max(na, nb))
                qa -= na
                qb -= nb
            tot["bwd_4x4_rings"] += int(trips)
    print(f"trips of the blend / recurrence loops over {len(sel)} sampled tiles (candidate slots, 2 per trip):")
    for k, v in tot.items():
        print(f"  {k:20s} {v}")
    f.close()


if __name__ == "__main__":
    main()
