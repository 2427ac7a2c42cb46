"""Work statistics of the forward compositor per (tile, warp) — debug aid for load-balance analysis."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "hair-gs_b200"), os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)
import torch
import bench
from hairgs_b200 import _lib as L
import diff_gaussian_rasterization._C as C
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
cfg = bench.WORKLOADS[name]
dev = torch.device("cuda:0")
h = bench.Harness(cfg, dev, C, 1, 0)
lib = L.load()
T = ((cfg["W"] + 15) // 16) * ((cfg["H"] + 15) // 16)
stats = torch.zeros(T * 8, 4, dtype=torch.int32, device=dev)
lib.hgs_debug_set_stats.argtypes = [L.c_void_p]
lib.hgs_debug_set_stats(stats.data_ptr())
h.step_resident(0)
torch.cuda.synchronize()
lib.hgs_debug_set_stats(None)
s = stats.cpu().long()
chunks, cand, blend, misc = s[:, 0], s[:, 1], s[:, 2], s[:, 3]
ndone, nch = misc & 0xff, misc >> 8
print("warps", s.shape[0], "active", int((nch > 0).sum()))
print("list chunks/tile: mean %.1f max %d" % (nch[nch > 0].float().mean(), nch.max()))
print("chunks walked: total %d (of %d available) max/warp %d" % (chunks.sum(), nch.sum(), chunks.max()))
print("candidates: total %d  per walked chunk %.2f  max/warp %d" % (cand.sum(), cand.sum() / max(chunks.sum(), 1), cand.max()))
print("blends (lane-level): total %d  per candidate %.2f" % (blend.sum(), blend.sum() / max(cand.sum(), 1)))
cost = chunks * 30 + cand * 45
print("est. cost: total %.1fM  max/warp %.0f  mean/active warp %.0f" % (cost.sum() / 1e6, cost.max(), cost[nch > 0].float().mean()))
top = torch.argsort(cost, descending=True)[:12]
for i in top:
    print(" tile %5d warp %d chunks %4d/%4d cand %5d blends %6d done %2d" % (i // 8, i % 8, chunks[i], nch[i], cand[i], blend[i], ndone[i]))
tile_cost = cost.view(-1, 8).max(1).values
print("sum over tiles of max-warp cost %.1fM vs sum of all warp cost/8 %.1fM" % (tile_cost.sum() / 1e6, cost.sum() / 8e6))
hist = torch.histc(cost[nch > 0].float(), bins=10, min=0, max=float(cost.max()))
print("cost histogram", hist.long().tolist())
