"""A/B of the (tile | depth) pair sort on the rasterizer's REAL keys: hgs_sort_pairs (hand-written onesweep) against the
reference's call, cub::DeviceRadixSort::SortPairs(u64, u32, 0, 32 + getHigherMsb(tiles)) (rasterizer_impl.cu:300-308), both
on the same device buffers of the same box (tools/cub_sort.cu -> tools/build/libcubsort.so).

    python tools/sort_bench.py [cfg3 cfg5 ...]      # prints one JSON line per workload

Keys: one forward pass of the workload's view 0 gives the sorted keys / point list; ordering that list by Gaussian id
(stable) restores the emission order of duplicateWithKeys, i.e. what the sort really receives.  Timed with CUDA events over
20 runs after 5 warm-ups; each run re-copies the inputs (the copy is timed alone and subtracted)."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "hair-gs_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import common  # noqa: E402
from hairgs_b200 import _lib as L  # noqa: E402

CASES = {"cfg2": lambda d: common.blob_inputs(300000, 512, 512, d, seed=0),
         "cfg3": lambda d: common.strand_inputs(10000, 100, 1024, 1024, d, seed=0),
         "cfg5": lambda d: common.strand_inputs(40000, 101, 2048, 2048, d, seed=0)}


def timeit(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000.0   # us


def main():
    lib = L.load()
    cub = ctypes.CDLL(os.path.join(ROOT, "tools", "build", "libcubsort.so"))
    cub.cub_sort_temp_bytes.restype = ctypes.c_size_t
    cub.cub_sort_temp_bytes.argtypes = [ctypes.c_int64, ctypes.c_int]
    cub.cub_sort_pairs.argtypes = [ctypes.c_int64, ctypes.c_int] + [ctypes.c_void_p] * 7
    dev = torch.device("cuda:0")
    peak = 6549.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for name in (sys.argv[1:] or ["cfg3", "cfg5"]):
        d = CASES[name](dev)
        N, _, _, _, v = common.ours_forward(d)
        order = torch.sort(v["point_list"], stable=True).indices
        keys, vals = v["point_list_keys"][order].contiguous(), v["point_list"][order].contiguous()
        tiles = ((d["image_width"] + 15) // 16) * ((d["image_height"] + 15) // 16)
        end_bit = 32 + tiles.bit_length()
        del v
        ko, vo = torch.empty_like(keys), torch.empty_like(vals)
        ki, vi = torch.empty_like(keys), torch.empty_like(vals)
        ws = torch.empty(lib.hgs_sort_bytes(N), dtype=torch.uint8, device=dev)
        tb = cub.cub_sort_temp_bytes(N, end_bit)
        tmp = torch.empty(tb, dtype=torch.uint8, device=dev)
        s = L.stream_ptr(dev)

        def copy_in():
            ki.copy_(keys)
            vi.copy_(vals)

        def ours():
            copy_in()
            L.check(lib.hgs_sort_pairs(N, end_bit, ki.data_ptr(), vi.data_ptr(), ko.data_ptr(), vo.data_ptr(), ws.data_ptr(), s))

        def cub_ref():
            copy_in()
            assert cub.cub_sort_pairs(N, end_bit, ki.data_ptr(), vi.data_ptr(), ko.data_ptr(), vo.data_ptr(), tmp.data_ptr(), tb, s) == 0

        t_copy = timeit(copy_in)
        t_ours = timeit(ours) - t_copy
        ok_ours = bool((ko[1:] >= ko[:-1]).all())
        ours_k, ours_v = ko.clone(), vo.clone()
        t_cub = timeit(cub_ref) - t_copy
        same = bool(torch.equal(ko, ours_k) and torch.equal(vo, ours_v))
        passes = (end_bit + 7) // 8
        alg = N * (8 + 24 * passes)       # histogram read + passes x (read + write of key and value)
        print(json.dumps({"workload": name, "pairs": int(N), "end_bit": end_bit, "passes": passes,
                          "ours_us": round(t_ours, 1), "cub_us": round(t_cub, 1), "ours_over_cub": round(t_cub / t_ours, 3),
                          "ours_GBps": round(alg / t_ours / 1e3, 1), "cub_GBps": round(alg / t_cub / 1e3, 1),
                          "ours_frac_of_hbm_peak": round(alg / t_ours / 1e3 / peak, 4),
                          "cub_frac_of_hbm_peak": round(alg / t_cub / 1e3 / peak, 4), "hbm_peak_GBps": peak,
                          "sorted": ok_ours, "bit_identical_to_cub": same,
                          "note": "full reference key width (no depth-range compaction); in the rasterizer the compaction "
                                  "drops one of these passes"}), flush=True)


if __name__ == "__main__":
    main()
