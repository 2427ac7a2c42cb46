// tools/cub_sort.cu - A/B harness: the reference's sort call, cub::DeviceRadixSort::SortPairs(keys u64, values u32,
// begin_bit 0, end_bit 32 + getHigherMsb(tiles)) exactly as rasterizer_impl.cu:300-308 issues it, behind a C entry point
// so tools/sort_bench.py can time it on the SAME device buffers as hgs_sort_pairs.  CUDA 12.9's CUB dispatches to its
// sm_100-tuned onesweep policy (cub/device/dispatch/tuning/tuning_radix_sort.cuh, Policy1000).  Measurement tool only:
// not linked into libhairgs_rast.so.   Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC
#include <cub/cub.cuh>
#include <cstdint>

extern "C" size_t cub_sort_temp_bytes(int64_t n, int end_bit) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, n, 0, end_bit);
    return bytes;
}

extern "C" int cub_sort_pairs(int64_t n, int end_bit, const uint64_t* keys_in, const uint32_t* vals_in, uint64_t* keys_out,
                              uint32_t* vals_out, void* temp, size_t temp_bytes, void* stream) {
    return (int)cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, n, 0, end_bit,
                                                (cudaStream_t)stream);
}
