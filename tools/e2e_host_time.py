"""Is the e2e step host-bound?  Host time to ENQUEUE one step_e2e (no sync) vs device time per step."""
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402

cfg = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]
dev = torch.device("cuda:0")
from diff_gaussian_rasterization import _C  # noqa: E402
h = bench.Harness(cfg, dev, _C, 1, 0)
h.setup_e2e(fused=True)
for it in range(10):
    h.step_e2e(it)
torch.cuda.synchronize()
n = 200
t0 = time.perf_counter()
for it in range(10, 10 + n):
    h.step_e2e(it)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / n:.3f} ms/step; wall incl. drain {1e3 * (t2 - t0) / n:.3f} ms/step")
import cProfile
import pstats
pr = cProfile.Profile()
pr.enable()
for it in range(n):
    h.step_e2e(it)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
