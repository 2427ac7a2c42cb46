"""Time hgs_adam_step (FlatAdam) against torch.optim.Adam on the cfg3 parameter set."""
import os
import sys
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hair-gs_b200"))
from hairgs_b200 import optim  # noqa: E402

dev = torch.device("cuda:0")
shapes = {"_endpoints": (1000000, 3), "_width": (990000, 1), "_opacity": (990000, 1), "_mask": (990000, 1),
          "_features_dc": (990000, 1, 3)}


def make():
    return {k: torch.nn.Parameter(torch.randn(*s, device=dev)) for k, s in shapes.items()}


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


p1 = make()
fa = optim.FlatAdam([{"params": [p], "lr": 1e-3, "name": k} for k, p in p1.items()])
fa.grads.flat.normal_()
n = fa.flat_param.numel()
t = timeit(lambda: fa.step(zero_grad=False))
print(f"FlatAdam: {t * 1e3:.1f} us for {n} elements -> {28 * n / t / 1e6:.0f} GB/s")
t = timeit(lambda: fa.step(zero_grad=True))
print(f"FlatAdam + grad clear: {t * 1e3:.1f} us -> {32 * n / t / 1e6:.0f} GB/s")
p2 = make()
ta = torch.optim.Adam([{"params": [p], "lr": 1e-3, "name": k} for k, p in p2.items()], lr=0.0, eps=1e-15)
for p in p2.values():
    p.grad = torch.randn_like(p)
t = timeit(ta.step)
print(f"torch.optim.Adam (default foreach): {t * 1e3:.1f} us")
tf = torch.optim.Adam([{"params": [p], "lr": 1e-3, "name": k} for k, p in p2.items()], lr=0.0, eps=1e-15, fused=True)
t = timeit(tf.step)
print(f"torch.optim.Adam (fused=True): {t * 1e3:.1f} us")
