// knn.cu — distCUDA2: mean squared distance to the 3 nearest neighbours of every point.
// Replaces SimpleKNN::knn (simple_knn.cu:186-222: cub reduce x2 + 2 blocking copies, Morton codes,
// cub sort, 1024-point boxes, boxMeanDist).  Same exact-search contract (self excluded by index,
// duplicates count as distance 0, mean of the three smallest squared distances, FLT_MAX when fewer
// than three neighbours exist); different machinery: device-resident bounds (no host sync), our own
// onesweep sort, and a two-level (32-point / 1024-point) box hierarchy so the per-point cost is
// O(P/1024 + hits*32) box tests instead of O(P/1024) tests + 1024-point brute force per hit.
#include <cfloat>
#include "hgs_common.cuh"

namespace hgs {

int launch_sort_pairs(int64_t, const uint32_t*, int, uint64_t* [2], uint32_t* [2], void*, int*, cudaStream_t, GeomHeader*, int, int);

static constexpr int kBox = 32;

struct KnnLayout {
    uint32_t* bounds;     // [6] order-preserving encodings of min xyz, max xyz
    uint64_t* keys[2];
    uint32_t* vals[2];
    float4* sorted;       // [P] xyz + original index bits
    float4* box_lo;       // [nbox]
    float4* box_hi;
    float4* sbox_lo;      // [nsbox]
    float4* sbox_hi;
    void* sort_ws;
    size_t bytes;
    int nbox, nsbox;
};

static KnnLayout carve_knn(void* base, int P) {
    KnnLayout k;
    char* p = (char*)base;
    size_t off = 0;
    size_t Pz = (size_t)(P > 0 ? P : 0);
    k.nbox = (int)((Pz + kBox - 1) / kBox);
    k.nsbox = (k.nbox + kBox - 1) / kBox;
    k.bounds = (uint32_t*)(p + off); off = align_up(off + 6 * 4);
    k.keys[0] = (uint64_t*)(p + off); off = align_up(off + Pz * 8);
    k.keys[1] = (uint64_t*)(p + off); off = align_up(off + Pz * 8);
    k.vals[0] = (uint32_t*)(p + off); off = align_up(off + Pz * 4);
    k.vals[1] = (uint32_t*)(p + off); off = align_up(off + Pz * 4);
    k.sorted = (float4*)(p + off); off = align_up(off + Pz * 16);
    k.box_lo = (float4*)(p + off); off = align_up(off + (size_t)k.nbox * 16);
    k.box_hi = (float4*)(p + off); off = align_up(off + (size_t)k.nbox * 16);
    k.sbox_lo = (float4*)(p + off); off = align_up(off + (size_t)k.nsbox * 16);
    k.sbox_hi = (float4*)(p + off); off = align_up(off + (size_t)k.nsbox * 16);
    k.sort_ws = (void*)(p + off); off = align_up(off + carve_sort(nullptr, P).bytes);
    k.bytes = off;
    return k;
}

size_t knn_bytes(int P) { return carve_knn(nullptr, P).bytes; }

__device__ __forceinline__ uint32_t enc_f(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f(uint32_t e) {
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

__global__ void knn_init_bounds(uint32_t* bounds) {
    if (threadIdx.x < 3) bounds[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) bounds[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) knn_bounds_kernel(int P, const float* __restrict__ pts, uint32_t* bounds) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = pts[3 * (size_t)i + c];
            lo[c] = fminf(lo[c], v);
            hi[c] = fmaxf(hi[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            atomicMin(&bounds[c], enc_f(lo[c]));
            atomicMax(&bounds[3 + c], enc_f(hi[c]));
        }
    }
}

__device__ __forceinline__ uint32_t spread10(uint32_t x) {  // 10 bits -> every third bit
    x &= 0x3ffu;
    x = (x | (x << 16)) & 0x030000FFu;
    x = (x | (x << 8)) & 0x0300F00Fu;
    x = (x | (x << 4)) & 0x030C30C3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

__global__ void knn_morton_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ bounds,
                                  uint64_t* keys, uint32_t* vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    uint32_t code = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float lo = dec_f(bounds[c]), hi = dec_f(bounds[3 + c]);
        const float ext = hi - lo;
        float u = ext > 0.f ? (pts[3 * (size_t)i + c] - lo) / ext : 0.f;
        u = fminf(fmaxf(u, 0.f), 1.f);
        code |= spread10((uint32_t)(u * 1023.f)) << c;
    }
    keys[i] = code;
    vals[i] = (uint32_t)i;
}

__global__ void knn_gather_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ order, float4* sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint32_t id = order[i];
    sorted[i] = make_float4(pts[3 * (size_t)id], pts[3 * (size_t)id + 1], pts[3 * (size_t)id + 2], __uint_as_float(id));
}

// one warp per group of 32 consecutive items; level 0 reads points, level 1 reads boxes
__global__ void __launch_bounds__(256) knn_box_kernel(int n_in, const float4* __restrict__ in_lo,
                                                      const float4* __restrict__ in_hi, float4* out_lo, float4* out_hi) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int n_out = (n_in + kBox - 1) / kBox;
    if (warp >= n_out) return;
    const int i = warp * kBox + lane;
    float3 lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX), hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    if (i < n_in) {
        const float4 a = in_lo[i], b = in_hi[i];
        lo = make_float3(a.x, a.y, a.z);
        hi = make_float3(b.x, b.y, b.z);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o));
        lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o));
        hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o));
        hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
    }
    if (lane == 0) {
        out_lo[warp] = make_float4(lo.x, lo.y, lo.z, 0.f);
        out_hi[warp] = make_float4(hi.x, hi.y, hi.z, 0.f);
    }
}

__device__ __forceinline__ float dist_box(const float4 lo, const float4 hi, const float3 p) {
    float dx = 0.f, dy = 0.f, dz = 0.f;
    if (p.x < lo.x || p.x > hi.x) dx = fminf(fabsf(p.x - lo.x), fabsf(p.x - hi.x));
    if (p.y < lo.y || p.y > hi.y) dy = fminf(fabsf(p.y - lo.y), fabsf(p.y - hi.y));
    if (p.z < lo.z || p.z > hi.z) dz = fminf(fabsf(p.z - lo.z), fabsf(p.z - hi.z));
    return dx * dx + dy * dy + dz * dz;
}

// insertion into the ascending triple (simple_knn.cu:133-146 semantics)
__device__ __forceinline__ void update3(const float3 ref, const float4 q, float* best) {
    const float3 d = make_float3(q.x - ref.x, q.y - ref.y, q.z - ref.z);
    float dist = d.x * d.x + d.y * d.y + d.z * d.z;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (best[j] > dist) {
            const float t = best[j];
            best[j] = dist;
            dist = t;
        }
    }
}

__global__ void __launch_bounds__(128) knn_search_kernel(int P, const float4* __restrict__ sorted,
                                                         const float4* __restrict__ box_lo,
                                                         const float4* __restrict__ box_hi, int nbox,
                                                         const float4* __restrict__ sbox_lo,
                                                         const float4* __restrict__ sbox_hi, int nsbox,
                                                         float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float4 me = sorted[i];
    const float3 p = make_float3(me.x, me.y, me.z);
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    // seed a rejection radius from the Morton neighbourhood (upper bound on the true 3rd distance)
    for (int j = max(0, i - 3); j <= min(P - 1, i + 3); ++j) {
        if (j == i) continue;
        update3(p, sorted[j], best);
    }
    const float reject = best[2];
    best[0] = FLT_MAX; best[1] = FLT_MAX; best[2] = FLT_MAX;
    const float slack = 1.0f - 4e-6f;  // never prune a box the exact float search would have opened
    for (int sb = 0; sb < nsbox; ++sb) {
        const float ds = dist_box(sbox_lo[sb], sbox_hi[sb], p) * slack;
        if (ds > reject || ds > best[2]) continue;
        const int b1 = min(nbox, (sb + 1) * kBox);
        for (int b = sb * kBox; b < b1; ++b) {
            const float db = dist_box(box_lo[b], box_hi[b], p) * slack;
            if (db > reject || db > best[2]) continue;
            const int e1 = min(P, (b + 1) * kBox);
            for (int j = b * kBox; j < e1; ++j) {
                if (j == i) continue;
                update3(p, sorted[j], best);
            }
        }
    }
    out[__float_as_uint(me.w)] = (best[0] + best[1] + best[2]) / 3.0f;
}

int launch_knn(int P, const float* points, float* out, void* ws, cudaStream_t s) {
    KnnLayout k = carve_knn(ws, P);
    StageScope prof(HGS_STAGE_KNN, s);
    knn_init_bounds<<<1, 32, 0, s>>>(k.bounds);
    const int nb = (P + 255) / 256;
    knn_bounds_kernel<<<min(nb, 148 * 4), 256, 0, s>>>(P, points, k.bounds);
    knn_morton_kernel<<<nb, 256, 0, s>>>(P, points, k.bounds, k.keys[0], k.vals[0]);
    if (int e = check_cuda(cudaGetLastError(), "knn morton launch")) return e;
    int res = 0;
    if (int e = launch_sort_pairs(P, nullptr, 30, k.keys, k.vals, k.sort_ws, &res, s, nullptr, 32, 0)) return e;
    knn_gather_kernel<<<nb, 256, 0, s>>>(P, points, k.vals[res], k.sorted);
    knn_box_kernel<<<(k.nbox * 32 + 255) / 256, 256, 0, s>>>(P, k.sorted, k.sorted, k.box_lo, k.box_hi);
    knn_box_kernel<<<(k.nsbox * 32 + 255) / 256, 256, 0, s>>>(k.nbox, k.box_lo, k.box_hi, k.sbox_lo, k.sbox_hi);
    knn_search_kernel<<<(P + 127) / 128, 128, 0, s>>>(P, k.sorted, k.box_lo, k.box_hi, k.nbox, k.sbox_lo, k.sbox_hi,
                                                       k.nsbox, out);
    return check_cuda(cudaGetLastError(), "knn search launch");
}

}  // namespace hgs
