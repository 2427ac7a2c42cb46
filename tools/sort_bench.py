"""Stand-alone throughput of hgs_sort_pairs (u64 key, u32 value) vs torch.sort, for the roofline table."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hair-gs_b200"))
import torch
from hairgs_b200 import _lib as L
lib = L.load()
dev = torch.device("cuda:0")
sizes = [int(a) for a in sys.argv[1:]] or [1 << 20, 1730000, 1 << 23, 1 << 25]
for n in sizes:
    for end_bit in (45,):
        g = torch.Generator(device="cuda").manual_seed(1)
        keys = (torch.randint(0, 4096, (n,), generator=g, device=dev, dtype=torch.int64) << 32) | \
            torch.randint(0x3e000000, 0x3f800000, (n,), generator=g, device=dev, dtype=torch.int64)
        vals = torch.arange(n, device=dev, dtype=torch.int32)
        ws = torch.empty(lib.hgs_sort_bytes(n), dtype=torch.uint8, device=dev)
        ko, vo = torch.empty_like(keys), torch.empty_like(vals)
        def run():
            ki, vi = keys.clone(), vals.clone()
            L.check(lib.hgs_sort_pairs(n, end_bit, ki.data_ptr(), vi.data_ptr(), ko.data_ptr(), vo.data_ptr(), ws.data_ptr(), L.stream_ptr(dev)))
        def clone_only():
            ki, vi = keys.clone(), vals.clone()
        def tsort():
            torch.sort(keys, stable=True)
        res = {}
        for name, fn in (("ours+clone", run), ("clone", clone_only), ("torch.sort(int64)+idx", tsort)):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / 10
        t = res["ours+clone"] - res["clone"]
        passes = (end_bit + 7) // 8
        gb = n * (8 + 24 * passes) / t / 1e6
        print(f"n={n:9d} end_bit={end_bit} ours={t*1000:8.1f} us ({gb:7.1f} GB/s algorithmic, {passes} passes)  torch.sort={res['torch.sort(int64)+idx']*1000:8.1f} us")
        ref_k, order = torch.sort(keys, stable=True)
        assert torch.equal(ko, ref_k) and torch.equal(vo.long(), order)
