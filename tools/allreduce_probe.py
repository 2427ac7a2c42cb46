"""Where does the multi-GPU step time go?  Run under torchrun with N ranks:
   (1) NCCL all_reduce of the fused gradient bucket alone, (2) the fused resident step with and without the collective,
   (3) per-view step times on each rank (imbalance)."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = bench.WORKLOADS["cfg3"]
from diff_gaussian_rasterization import _C  # noqa: E402
h = bench.Harness(cfg, dev, _C, world, rank)
h.setup_fused()


def timed(fn, n, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


flat = h.fbucket.flat
t_ar = timed(lambda i: dist.all_reduce(flat), 50)
t_step = timed(h.step_resident_fused, 48)
saved = h.fbucket.all_reduce
h.fbucket.all_reduce = lambda *a, **k: None
t_noar = timed(h.step_resident_fused, 48)
per_view = []
for v in range(len(h.my_views)):
    per_view.append(timed(lambda i, v=v: h.step_resident_fused(v), 10, warm=2))
h.fbucket.all_reduce = saved
if rank == 0:
    print(f"world {world}: bucket {flat.numel() * 4 / 1e6:.1f} MB; all_reduce alone {t_ar:.3f} ms; step {t_step:.3f} ms; "
          f"step without collective {t_noar:.3f} ms")
    print("per-view max-over-ranks step ms (no collective):", [round(x, 3) for x in per_view])
dist.destroy_process_group()
