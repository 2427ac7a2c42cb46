// hgs_common.cuh — shared device helpers, workspace layouts and launch utilities of libhairgs_rast.
// sm_100a only.  Nothing here is derived from the reference's sources; where an arithmetic order is
// part of the numerical contract the reference file:line is cited.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>
#include "../../include/hairgs_rast.h"

#define HGS_TILE_PIX (HGS_TILE * HGS_TILE)

namespace hgs {

// ------------------------------------------------------------------------------------------------
// error plumbing (host)
// ------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int stage_check(const char* stage, int debug, cudaStream_t s);

// Per-device one-time initialisation (function attributes and the like are per device, and a process may drive
// several): true exactly once per (flag word, current device); thread-safe.
inline bool first_call_on_device(std::atomic<unsigned long long>& done) {
    int d = 0;
    cudaGetDevice(&d);
    const unsigned long long bit = 1ull << (d & 63);
    return (done.fetch_or(bit) & bit) == 0;
}

// ------------------------------------------------------------------------------------------------
// Stage profiler: counts every kernel launch of the library and, when enabled through
// hgs_profile_enable(), brackets each launch with CUDA events on the launching stream
// (bench.py's roofline numbers come from here; see hgs_profile_collect in api.cu).
// ------------------------------------------------------------------------------------------------
struct StageScope {
    int stage;
    cudaStream_t stream;
    cudaEvent_t stop;
    StageScope(int stage_id, cudaStream_t s);
    ~StageScope();
};

// ------------------------------------------------------------------------------------------------
// Workspace layouts.  All sub-arrays 256-B aligned; the same carve is repeated by forward, backward
// and the state viewers (the role GeometryState/ImageState/BinningState::fromChunk play in the
// reference, rasterizer_impl.cu:155-194 — but a different, 32-byte-record layout).
// ------------------------------------------------------------------------------------------------
static constexpr size_t kAlign = 256;
__host__ __device__ inline size_t align_up(size_t x, size_t a = kAlign) { return (x + a - 1) / a * a; }

struct GeomHeader {            // first 256 B of the geometry workspace
    uint32_t num_rendered;     // total instance count N (written by the last preprocess block)
    uint32_t block_ticket;     // dynamic block id for the in-kernel chained scan
    uint32_t overflow;         // bit 0: N would exceed 2^31-1; bit 1: the depth range did not fit the sort's depth bits
    uint32_t depth_max;        // max over visible Gaussians of the depth's bit pattern (tile_scan)
    uint32_t depth_min_inv;    // max of ~bits, i.e. ~min (so that the zero fill of the header is the identity)
    uint32_t pad[59];
};

// One 32-byte record per Gaussian: everything the compositors need besides colour, in one sector.
struct __align__(16) GaussRecLo { float x, y, conic_a, conic_b; };   // means2D + conic.xy
struct __align__(16) GaussRecHi { float conic_c, opacity, hx, hy; }; // conic.z + opacity + alpha>=1/255 half extents

static constexpr int kPreprocThreads = 256;

struct GeomLayout {
    GeomHeader* hdr;
    unsigned long long* scan_state;  // [ceil(P/256)] chained-scan status words (directly after hdr: one memset clears both)
    uint32_t* tile_count;     // [tiles * HGS_TILE_SLICES] instances per (tile, depth slice), counted by preprocess_fwd (also
                              // inside the cleared region)
    float4* rec;              // [2P] : rec[2i] = GaussRecLo, rec[2i+1] = GaussRecHi
    float* rgb;               // [P*cstride] colours used by the compositors (SH result or repacked colors_precomp)
    float* depths;            // [P]
    uint32_t* tiles_touched;  // [P]
    uint32_t* offsets;        // [P] inclusive scan of tiles_touched
    uint8_t* clamped;         // [P] bit c set when SH channel c was clamped at 0
    uint2* rects;             // [P] tile rect: .x = xmin | ymin<<16, .y = xmax | ymax<<16 (valid when tiles_touched > 0)
    size_t bytes;
    size_t clear_bytes;       // hdr + scan_state: what must be zeroed before stage A
    int cstride;
};

__host__ __device__ inline int color_stride(int channels) { return channels <= 4 ? 4 : 8; }
// depth slice of an instance (HGS_SORT_TILE): monotone in the depth's bit pattern, so the slices of a tile list are ordered
__host__ __device__ inline uint32_t depth_slice(uint32_t depth_bits, uint32_t base, int shift) {
    if (shift >= 32) return 0u;
    const uint32_t d = depth_bits > base ? depth_bits - base : 0u;
    const uint32_t s = d >> shift;
    return s < (uint32_t)HGS_TILE_SLICES ? s : (uint32_t)HGS_TILE_SLICES - 1u;
}
__host__ __device__ inline size_t tile_count_of(int width, int height) {
    return (size_t)((width + HGS_TILE - 1) / HGS_TILE) * (size_t)((height + HGS_TILE - 1) / HGS_TILE);
}

__host__ __device__ inline GeomLayout carve_geom(void* base, int P, int channels, size_t tiles) {
    GeomLayout g;
    char* p = (char*)base;
    size_t off = 0;
    size_t Pz = (size_t)(P > 0 ? P : 0);
    size_t nblk = (Pz + kPreprocThreads - 1) / kPreprocThreads;
    g.cstride = color_stride(channels);
    g.hdr = (GeomHeader*)(p + off);            off = align_up(off + sizeof(GeomHeader));
    g.scan_state = (unsigned long long*)(p + off); off = align_up(off + nblk * 8);
    g.tile_count = (uint32_t*)(p + off);       off = align_up(off + tiles * HGS_TILE_SLICES * 4);
    g.clear_bytes = off;
    g.rec = (float4*)(p + off);                off = align_up(off + Pz * 32);
    g.rgb = (float*)(p + off);                 off = align_up(off + Pz * 4 * g.cstride);
    g.depths = (float*)(p + off);              off = align_up(off + Pz * 4);
    g.tiles_touched = (uint32_t*)(p + off);    off = align_up(off + Pz * 4);
    g.offsets = (uint32_t*)(p + off);          off = align_up(off + Pz * 4);
    g.clamped = (uint8_t*)(p + off);           off = align_up(off + Pz);
    g.rects = (uint2*)(p + off);               off = align_up(off + Pz * 8);
    g.bytes = off;
    return g;
}

struct ImageLayout {
    float* final_T;        // [H*W]
    uint32_t* n_contrib;   // [H*W]
    uint2* ranges;         // [tiles]  (the reference over-allocates H*W entries, rasterizer_impl.cu:172-178)
    uint32_t* tile_order;  // [tiles]  tile ids, longest list first (CTA i composites tile_order[i])
    // HGS_SORT_TILE (tilesort.cu): per (tile, slice) list start, scatter cursor, and the work lists of the three size classes
    uint32_t* list_start;  // [tiles * S]
    uint32_t* list_cursor; // [tiles * S]  (zeroed per pass together with work_count, which follows it)
    uint32_t* work_count;  // [8]  entries of the three work lists [0..2], claim counters of the sort kernels [4..6]
    uint32_t* work;        // [3][tiles * S]  non-empty lists by size class
    size_t bytes;
};

__host__ __device__ inline ImageLayout carve_image(void* base, int W, int H) {
    ImageLayout im;
    char* p = (char*)base;
    size_t off = 0;
    size_t hw = (size_t)W * H;
    size_t tiles = (size_t)((W + HGS_TILE - 1) / HGS_TILE) * ((H + HGS_TILE - 1) / HGS_TILE);
    im.final_T = (float*)(p + off);      off = align_up(off + hw * 4);
    im.n_contrib = (uint32_t*)(p + off); off = align_up(off + hw * 4);
    im.ranges = (uint2*)(p + off);       off = align_up(off + tiles * 8);
    im.tile_order = (uint32_t*)(p + off); off = align_up(off + tiles * 4);
    im.list_start = (uint32_t*)(p + off);  off = align_up(off + tiles * HGS_TILE_SLICES * 4);
    im.list_cursor = (uint32_t*)(p + off); off += tiles * HGS_TILE_SLICES * 4;
    im.work_count = (uint32_t*)(p + off);  off = align_up(off + 8 * 4);
    im.work = (uint32_t*)(p + off);        off = align_up(off + 3 * tiles * HGS_TILE_SLICES * 4);
    im.bytes = off;
    return im;
}

// Radix sort geometry (binning.cu)
static constexpr int kSortThreads = 256;
static constexpr int kSortItems = 16;                       // keys per thread
static constexpr int kSortTile = kSortThreads * kSortItems; // 4096 keys per block
static constexpr int kRadixBits = 8;
static constexpr int kRadix = 1 << kRadixBits;
static constexpr int kMaxPasses = 8;

struct SortLayout {
    uint32_t* hist;     // [kMaxPasses * 256] global digit histograms -> exclusive offsets
    uint32_t* tickets;  // [kMaxPasses] dynamic tile ids (a tile's id is the order in which its block STARTED)
    uint32_t* status;   // [kMaxPasses * ntiles * 256] per-tile digit counts (flag in bit 31)
    uint32_t* group;    // [kMaxPasses * ngroups * 256] per-group digit counts: arrivals << 24 | sum
    size_t bytes;
    size_t ntiles;
    size_t ngroups;
    int group_shift;    // tiles per group = 1 << group_shift
};

// Two-level look-back of the onesweep passes: tiles are grouped (32 / 64 / 128 per group, ~sqrt(ntiles)); a tile's
// exclusive prefix is the sum of the complete groups before its own plus the tiles before it inside its group.  Every
// word it reads is published by a tile right after its local ranking, so no tile ever waits for another tile's look-back
// (the classic chained scan serialises ~ntiles/32 L2 round trips when the whole grid is one wave, which is what a
// 1-2 M instance sort is on 148 SMs).
__host__ __device__ inline int sort_group_shift(size_t ntiles) { return ntiles <= 1024 ? 5 : (ntiles <= 4096 ? 6 : 7); }

__host__ __device__ inline SortLayout carve_sort(void* base, int64_t n) {
    SortLayout s;
    char* p = (char*)base;
    size_t off = 0;
    s.ntiles = (size_t)((n + kSortTile - 1) / kSortTile);
    if (s.ntiles == 0) s.ntiles = 1;
    s.group_shift = sort_group_shift(s.ntiles);
    s.ngroups = (s.ntiles + ((size_t)1 << s.group_shift) - 1) >> s.group_shift;
    s.hist = (uint32_t*)(p + off);    off = align_up(off + (size_t)kMaxPasses * kRadix * 4);
    s.tickets = (uint32_t*)(p + off); off = align_up(off + (size_t)kMaxPasses * 4);
    s.group = (uint32_t*)(p + off);   off = align_up(off + (size_t)kMaxPasses * s.ngroups * kRadix * 4);
    s.status = (uint32_t*)(p + off);  off = align_up(off + (size_t)kMaxPasses * s.ntiles * kRadix * 4);
    s.bytes = off;
    return s;
}

struct BinningLayout {
    uint64_t* keys[2];   // ping-pong key buffers; [0] receives the emitted (unsorted) keys
    uint32_t* vals[2];   // ping-pong Gaussian ids
    float4* pk_lo;       // [N] sorted-order copies of the Gaussian records (x, y, conic.x, conic.y)
    float4* pk_hi;       // [N] (conic.z, opacity, hx, hy)
    float4* pk_col;      // [N * cstride/4] colours in sorted order
    uint16_t* pk_mask;   // [N] which of the tile's sixteen 4x4 pixel blocks the instance's alpha extent reaches (block_mask16)
    void* sort_ws;
    size_t bytes;
};

__host__ __device__ inline BinningLayout carve_binning(void* base, int64_t n, int channels) {
    BinningLayout b;
    char* p = (char*)base;
    size_t off = 0;
    size_t nz = (size_t)(n > 0 ? n : 0);
    b.keys[0] = (uint64_t*)(p + off); off = align_up(off + nz * 8);
    b.keys[1] = (uint64_t*)(p + off); off = align_up(off + nz * 8);
    b.vals[0] = (uint32_t*)(p + off); off = align_up(off + nz * 4);
    b.vals[1] = (uint32_t*)(p + off); off = align_up(off + nz * 4);
    b.pk_lo = (float4*)(p + off);     off = align_up(off + nz * 16);
    b.pk_hi = (float4*)(p + off);     off = align_up(off + nz * 16);
    b.pk_col = (float4*)(p + off);    off = align_up(off + nz * 4 * color_stride(channels));
    b.pk_mask = (uint16_t*)(p + off); off = align_up(off + nz * 2);
    b.sort_ws = (void*)(p + off);
    SortLayout s = carve_sort(b.sort_ws, n);
    off = align_up(off + s.bytes);
    b.bytes = off;
    return b;
}

#ifdef __CUDACC__
// MUFU.RCP (<= 1 ulp, flush-to-zero) for gradient-side reciprocals of well-scaled arguments, where the IEEE division sequence
// (8 instructions with its range check and slow-path call) buys nothing; never used where bits are compared with the reference.
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Which of the sixteen 4x4 pixel blocks of the 16x16 tile at (X0, Y0) the instance's alpha >= 1/255 extent (preprocess.cu
// alpha_extent, an axis-aligned box) reaches: bit 4 * block_row + block_column.  The same comparisons, on the same float
// values, as the per-chunk test the compositors made on their own two blocks (a block is skipped only when a comparison
// is TRUE, so NaNs keep the instance) - evaluated ONCE per instance here instead of once per warp and pass there, and
// read back as 2 bytes instead of the 68-byte record by the 7 of 8 warps whose blocks the instance does not reach.
__device__ __forceinline__ uint32_t block_mask16(const float4 lo, const float4 hi, uint32_t X0, uint32_t Y0) {
    const float xl = lo.x - hi.z, xr = lo.x + hi.z;
    const float yt = lo.y - hi.w, yb = lo.y + hi.w;
    uint32_t cols = 0, rows = 0;
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
        if (!(xl > (float)(X0 + 4 * j + 3) || xr < (float)(X0 + 4 * j))) cols |= 1u << j;
        if (!(yt > (float)(Y0 + 4 * j + 3) || yb < (float)(Y0 + 4 * j))) rows |= 1u << (4 * j);
    }
    return cols * rows;  // rows has one bit per nibble: the product copies cols into every reached block row
}
#endif

// Number of 8-bit passes for keys whose significant bits are [0, end_bit), and which ping-pong
// buffer ends up holding the sorted data.
__host__ __device__ inline int sort_passes(int end_bit) { return (end_bit + kRadixBits - 1) / kRadixBits; }

// Bits needed for tile ids: same rule as the reference's getHigherMsb (rasterizer_impl.cu:35-50),
// restated as "smallest b with (n >> b) == 0" — identical results for every n >= 1.
__host__ inline int tile_id_bits(uint32_t n) {
    int b = 0;
    while (b < 32 && (n >> b)) ++b;
    return b;
}

// ------------------------------------------------------------------------------------------------
// device math helpers
// ------------------------------------------------------------------------------------------------
// Column-major 3x3 (m[c][r]) with the textbook product order sum_k A[k][r]*B[c][k], k ascending.
// The evaluation order is part of the bit-exactness contract with the reference build
// (SURVEY.md §7.3 item 1): products are accumulated left to right so the compiler's FMA
// contraction sees the same expression trees.
struct M3 {
    float m[3][3];
};

__device__ __forceinline__ M3 m3_mul(const M3& a, const M3& b) {
    M3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int q = 0; q < 3; ++q)
            r.m[c][q] = a.m[0][q] * b.m[c][0] + a.m[1][q] * b.m[c][1] + a.m[2][q] * b.m[c][2];
    return r;
}

__device__ __forceinline__ M3 m3_transpose(const M3& a) {
    M3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int q = 0; q < 3; ++q) r.m[c][q] = a.m[q][c];
    return r;
}

// x' = m0 x + m4 y + m8 z + m12 ... (auxiliary.h:58-77 convention: column-major 4x4)
__device__ __forceinline__ float3 xform_point_4x3(const float3 p, const float* __restrict__ m) {
    float3 t;
    t.x = m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12];
    t.y = m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13];
    t.z = m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14];
    return t;
}
__device__ __forceinline__ float4 xform_point_4x4(const float3 p, const float* __restrict__ m) {
    float4 t;
    t.x = m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12];
    t.y = m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13];
    t.z = m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14];
    t.w = m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15];
    return t;
}

// Pixel coordinate of an NDC coordinate, evaluated in double like the reference (auxiliary.h:41-44).
__device__ __forceinline__ float ndc_to_pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }

// Tile rectangle of a splat (auxiliary.h:46-56): C truncation, clamp to the grid.
__device__ __forceinline__ void tile_rect(const float2 p, int max_radius, uint2& rmin, uint2& rmax,
                                          const uint32_t gx, const uint32_t gy) {
    rmin.x = min(gx, (uint32_t)max((int)0, (int)((p.x - max_radius) / HGS_TILE)));
    rmin.y = min(gy, (uint32_t)max((int)0, (int)((p.y - max_radius) / HGS_TILE)));
    rmax.x = min(gx, (uint32_t)max((int)0, (int)((p.x + max_radius + HGS_TILE - 1) / HGS_TILE)));
    rmax.y = min(gy, (uint32_t)max((int)0, (int)((p.y + max_radius + HGS_TILE - 1) / HGS_TILE)));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- bulk asynchronous copies (TMA engine, 1-D: cp.async.bulk) completed through an mbarrier -----------------------------
// One elected lane arms the barrier with the byte count and issues the copies; every consumer waits on the barrier's
// phase parity.  Source, destination and size must be multiples of 16 bytes.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// non-blocking prefetch of the cache line(s) holding [p, p+bytes): lets a thread start every input stream it will
// need before the first data-dependent early-out (otherwise each cull test exposes one full DRAM latency)
__device__ __forceinline__ void prefetch_l1(const void* p, int bytes = 1) {
    const char* c = (const char*)p;
    for (int o = 0; o < bytes; o += 128) asm volatile("prefetch.global.L1 [%0];\n" ::"l"(c + o));
    if (bytes > 1) asm volatile("prefetch.global.L1 [%0];\n" ::"l"(c + bytes - 1));
}

// streaming 128-bit loads (read-once data: keep it out of L1)
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];\n"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// arguments of the strand endpoint merge search (merge.cu)
struct MergeArgs {
    int K;
    const float* points;        // [K,3] positions of the strand ends considered
    const float* dirs;          // [K,3] unit direction end -> its neighbouring joint
    const int* global_id;       // [K]   endpoint id of each strand end
    const int* other_end;       // [K]   endpoint id of the other end of the same strand (root <-> tip)
    double r2;                  // ball radius squared (cKDTree compares squared distances in double)
    double dir_th;              // cos(angle threshold); the float32 dot product is compared in double like numpy
    int bidirectional;
    int max_num_nn;             // <= 0: unlimited
};

}  // namespace hgs
