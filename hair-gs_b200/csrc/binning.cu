// binning.cu — instance generation, (tile|depth) sort and tile ranges:
//   emit_keys      : load-balanced key/value emission (replaces duplicateWithKeys rasterizer_impl.cu:70-111:
//                    there one thread serially writes its whole rect; here a block's instances are dealt
//                    round-robin to its threads so every store is coalesced)
//   radix_histogram + onesweep_pass : hand-written stable LSD radix sort, 8-bit digits, one global
//                    histogram pass + one chained-scan (decoupled look-back) scatter pass per digit, with
//                    match.any warp-aggregated ranking (replaces cub::DeviceRadixSort::SortPairs
//                    rasterizer_impl.cu:303-308)
//   (tile ranges are produced by finalize_sorted in composite_warp.cu)
#include "hgs_common.cuh"

#include <cstdlib>
#include <cstring>

namespace hgs {

// ------------------------------------------------------------------------------------------------
// key emission
// ------------------------------------------------------------------------------------------------
// key = (tile_id << 32) | float_bits(depth), value = Gaussian id; a Gaussian's instances occupy
// slots [offsets[i]-touched, offsets[i]) with tiles enumerated y-outer / x-inner
// (rasterizer_impl.cu:88-108).
__global__ void __launch_bounds__(256) emit_keys_kernel(int P, const uint2* __restrict__ rects,
                                                        const float* __restrict__ depths,
                                                        const uint32_t* __restrict__ offsets,
                                                        const uint32_t* __restrict__ touched_arr,
                                                        uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                        uint32_t grid_x, uint32_t capacity) {
    __shared__ uint32_t s_start[256];  // block-local exclusive start of each Gaussian's run
    __shared__ uint32_t s_xy[256];     // xmin | ymin << 16
    __shared__ uint32_t s_w[256];      // rect width in tiles
    __shared__ uint32_t s_depth[256];
    __shared__ uint32_t s_base, s_total;

    const int tid = threadIdx.x;
    const int g0 = blockIdx.x * 256;
    const int g = g0 + tid;
    uint32_t touched = 0, incl = 0;
    if (g < P) {
        touched = touched_arr[g];
        incl = offsets[g];
    }
    if (tid == 0) s_base = incl - touched;  // == offsets[g0-1]
    __syncthreads();
    const uint32_t base = s_base;
    s_start[tid] = (g < P) ? (incl - touched - base) : 0xffffffffu;
    if (touched > 0) {
        const uint2 r = rects[g];
        s_xy[tid] = r.x;
        s_w[tid] = (r.y & 0xffffu) - (r.x & 0xffffu);
        s_depth[tid] = __float_as_uint(depths[g]);
    }
    const int last = min(P - g0, 256) - 1;
    if (tid == last) s_total = incl - base;
    __syncthreads();
    const uint32_t total = s_total;
    for (uint32_t s = tid; s < total; s += 256) {
        // largest j with s_start[j] <= s
        int lo = 0, hi = last;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_start[mid] <= s) lo = mid; else hi = mid - 1;
        }
        const uint32_t k = s - s_start[lo];
        const uint32_t w = s_w[lo];
        const uint32_t xy = s_xy[lo];
        const uint32_t ky = k / w;
        const uint32_t kx = k - ky * w;
        const uint32_t x = (xy & 0xffffu) + kx;
        const uint32_t y = (xy >> 16) + ky;
        const uint64_t key = ((uint64_t)(y * grid_x + x) << 32) | (uint64_t)s_depth[lo];
        const uint32_t slot = base + s;
        if (slot < capacity) {
            keys[slot] = key;
            vals[slot] = (uint32_t)(g0 + lo);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// radix sort
// ------------------------------------------------------------------------------------------------
static constexpr uint32_t kStFlagAgg = 1u << 31;       // a tile's digit count has been published
static constexpr uint32_t kStValMask = (1u << 24) - 1;  // counts are < 2^24 (<= 128 tiles x 4096 keys per group)

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// One read of the keys, all digit histograms at once.
// The element count lives on the device (n_ptr, clamped to the buffer capacity) so that the host never has
// to wait for it before launching; n_ptr == nullptr means "exactly cap elements".
// Depth-range compaction of the rasterizer's keys (tile << 32 | depth bits).  All visible depths lie in
// [dmin, dmax] (GeomHeader, reduced by tile_scan); sorting (tile << rb | depth - dmin) orders the pairs exactly like
// the full key as long as dmax - dmin < 2^rb, and needs ceil((tile_bits + rb) / 8) passes instead of
// ceil((tile_bits + 32) / 8).  The keys in memory stay the reference's; only the digit extraction sees the compacted
// value.  rb comes from the host (a hint from earlier views); if the range does not fit, bit 1 of hdr->overflow is
// raised and the caller repeats the stage with rb = 32.  rb == 32 with dmin == 0 is the identity (hgs_sort_pairs).
struct KeyXform {
    uint32_t dmin, dmask;
    int rb;
    __device__ __forceinline__ uint64_t operator()(uint64_t k) const {
        const uint32_t d = min((uint32_t)k - dmin, dmask);
        return ((k >> 32) << rb) | d;
    }
};

__device__ __forceinline__ KeyXform load_key_xform(const GeomHeader* range_hdr, int rb) {
    KeyXform x;
    x.rb = rb;
    x.dmask = rb >= 32 ? 0xffffffffu : ((1u << rb) - 1u);
    x.dmin = (range_hdr != nullptr && rb < 32) ? ~range_hdr->depth_min_inv : 0u;
    return x;
}

__global__ void __launch_bounds__(256) radix_histogram_kernel(const uint64_t* __restrict__ keys, uint32_t cap,
                                                              const uint32_t* __restrict__ n_ptr, int passes,
                                                              uint32_t* __restrict__ ghist, GeomHeader* range_hdr, int rb) {
    const uint32_t n = n_ptr ? min(*n_ptr, cap) : cap;
    const KeyXform xf = load_key_xform(range_hdr, rb);
    if (range_hdr != nullptr && rb < 32 && blockIdx.x == 0 && threadIdx.x == 0 && n > 0) {
        if (((range_hdr->depth_max - xf.dmin) >> rb) != 0) atomicOr(&range_hdr->overflow, 2u);
    }
    __shared__ uint32_t s_hist[kMaxPasses * kRadix];
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t k = xf(keys[i]);
        for (int p = 0; p < passes; ++p) atomicAdd(&s_hist[p * kRadix + (uint32_t)((k >> (8 * p)) & 0xffu)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) {
        const uint32_t v = s_hist[i];
        if (v) atomicAdd(&ghist[i], v);
    }
}

struct OnesweepSmem {
    uint64_t keys[kSortTile];
    uint32_t vals[kSortTile];
    uint32_t warp_hist[kSortThreads / 32][kRadix];
    uint32_t excl[kRadix];   // block-local exclusive digit offsets
    uint32_t gbase[kRadix];  // global position of local sorted index 0 of each digit (minus excl)
    uint32_t scan_tmp[kSortThreads / 32];
    uint32_t tile;
};

// block-wide exclusive scan of one value per thread (256 threads); returns exclusive prefix
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* tmp /*[8]*/) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += n;
    }
    if (lane == 31) tmp[warp] = incl;
    __syncthreads();
    uint32_t wex = 0;
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w)
        if ((uint32_t)w < warp) wex += tmp[w];
    __syncthreads();
    return wex + incl - v;
}

// kBallot: how a key finds the lanes of its warp that hold the same digit.  false: one match.any per key (~70 cycles of
// the SM's single ADU pipe per warp instruction: the pass is ADU-bound, 75 % busy at 8 M pairs, profiles/r1_sort_8m.md).
// true: nbits warp votes (one per digit bit, intersected), CUB's MatchAny.  Measured on B200 (profiles/r2_sort.md): the
// votes go through the same pipe - 8 of them cost what one match.any costs, so the two variants time within 5 %.
template <bool kBallot>
__global__ void __launch_bounds__(kSortThreads, 3) onesweep_pass_kernel(
    const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, uint32_t cap, const uint32_t* __restrict__ n_ptr, int shift,
    const uint32_t* __restrict__ ghist /*[256]*/,
    uint32_t* __restrict__ status /*[ntiles*256]*/, uint32_t* __restrict__ group /*[ngroups*256]*/, int group_shift,
    uint32_t* __restrict__ ticket, const GeomHeader* range_hdr, int rb, int nbits) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OnesweepSmem& sm = *reinterpret_cast<OnesweepSmem*>(smem_raw);

    const uint32_t n = n_ptr ? min(*n_ptr, cap) : cap;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // A tile's id is the order in which its block STARTED (one atomic per block), not blockIdx.x: the look-back below
    // waits on words published by tiles with smaller ids, and CUDA gives no guarantee that blocks are dispatched in
    // index order (MPS, preemption, concurrent graph branches) - with tickets every predecessor is running or done.
    if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w) sm.warp_hist[w][tid] = 0;
    __syncthreads();
    const uint32_t tile = sm.tile;
    if (tile * (uint32_t)kSortTile >= n) return;  // tile beyond the live range (grid is sized by capacity)
    const KeyXform xf = load_key_xform(range_hdr, rb);
    const uint32_t base = tile * kSortTile;
    const uint32_t nvalid = min((uint32_t)kSortTile, n - base);

    // ---- load keys, warp-striped: item r of this lane is element warp*512 + r*32 + lane ----------
    uint64_t key[kSortItems];
    const uint32_t wbase = warp * (kSortItems * 32) + lane;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t li = wbase + r * 32;
        key[r] = (li < nvalid) ? keys_in[base + li] : ~0ull;
    }

    // ---- per-warp ranking with match.any ---------------------------------------------------------
    // One shared-memory atomic per distinct digit per round hands every peer group its running count: no
    // __syncwarp between rounds, so the 16 rounds pipeline instead of serialising on LDS->STS round trips.
    // (Atomics of one warp to one address are performed in program order, which is what stability needs.)
    uint32_t rank[kSortItems];
    // padding keys of the last, partial tile carry digit 0xff: all 8 bits are compared there so that they never pair up
    // with a real top digit whose low nbits happen to be all ones
    const int nb = (nvalid < (uint32_t)kSortTile) ? kRadixBits : nbits;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t* wh = sm.warp_hist[warp];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t d = (uint32_t)(xf(key[r]) >> shift) & 0xffu;
        uint32_t peers;
        if (kBallot) {
            peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < kRadixBits; ++b) {
                if (b < nb) {  // the top digit of a key may have fewer than 8 significant bits
                    const bool bit = (d >> b) & 1u;
                    const uint32_t m = __ballot_sync(0xffffffffu, bit);
                    peers &= bit ? m : ~m;
                }
            }
        } else {
            peers = __match_any_sync(0xffffffffu, d);
        }
        const uint32_t lrank = __popc(peers & lt_mask);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lrank == 0) old = atomicAdd(&wh[d], (uint32_t)__popc(peers));
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = old + lrank;
    }
    __syncthreads();

    // ---- digit `tid`: prefix over warps, block total ----------------------------------------------
    uint32_t block_count = 0;
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w) {
        const uint32_t c = sm.warp_hist[w][tid];
        sm.warp_hist[w][tid] = block_count;
        block_count += c;
    }
    // publish as early as possible: the tile's own word and its contribution to its group's word (one arrival each)
    st_relaxed_u32(status + (size_t)tile * kRadix + tid, kStFlagAgg | block_count);
    const uint32_t my_group = tile >> group_shift, in_group = tile & ((1u << group_shift) - 1u);
    atomicAdd(group + (size_t)my_group * kRadix + tid, (1u << 24) | block_count);

    // values: issue the loads now, they are consumed after the key scatter
    uint32_t val[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t li = wbase + r * 32;
        val[r] = (li < nvalid) ? vals_in[base + li] : 0u;
    }

    const uint32_t local_excl = block_excl_scan_256(block_count, sm.scan_tmp);
    const uint32_t gexcl = block_excl_scan_256(ghist[tid], sm.scan_tmp);
    sm.excl[tid] = local_excl;
    __syncthreads();

    // ---- scatter keys AND values into block-sorted order in shared memory: needs only block-local offsets, so it
    //      overlaps the predecessors' chain latency instead of waiting behind the look-back ------------------------
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t d = (uint32_t)(xf(key[r]) >> shift) & 0xffu;
        const uint32_t pos = sm.excl[d] + wh[d] + rank[r];
        sm.keys[pos] = key[r];
        sm.vals[pos] = val[r];
    }

    // ---- two-level look-back (see sort_group_shift): complete groups before mine + tiles before me in my group.
    //      kLook words per L2 round trip; a word that is not there yet is simply read again -------------------------------
    uint32_t prev_sum = 0;
    {
        constexpr int kLook = 16;
        const uint32_t full = 1u << group_shift;  // arrivals of a complete group (every predecessor group is complete)
        for (uint32_t k0 = 0; k0 < my_group; k0 += kLook) {
            uint32_t w[kLook];
#pragma unroll
            for (int k = 0; k < kLook; ++k)
                w[k] = (k0 + k < my_group) ? ld_relaxed_u32(group + (size_t)(k0 + k) * kRadix + tid) : (full << 24);
#pragma unroll
            for (int k = 0; k < kLook; ++k) {
                while ((w[k] >> 24) != full) w[k] = ld_relaxed_u32(group + (size_t)(k0 + k) * kRadix + tid);
                prev_sum += w[k] & kStValMask;
            }
        }
        const uint32_t first = tile - in_group;
        for (uint32_t k0 = 0; k0 < in_group; k0 += kLook) {
            uint32_t w[kLook];
#pragma unroll
            for (int k = 0; k < kLook; ++k)
                w[k] = (k0 + k < in_group) ? ld_relaxed_u32(status + (size_t)(first + k0 + k) * kRadix + tid) : kStFlagAgg;
#pragma unroll
            for (int k = 0; k < kLook; ++k) {
                while (!(w[k] & kStFlagAgg)) w[k] = ld_relaxed_u32(status + (size_t)(first + k0 + k) * kRadix + tid);
                prev_sum += w[k] & kStValMask;
            }
        }
    }
    sm.gbase[tid] = gexcl + prev_sum - local_excl;
    __syncthreads();

    // ---- coalesced runs out to global -------------------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint32_t li = tid + k * kSortThreads;
        if (li < nvalid) {
            const uint64_t kk = sm.keys[li];
            const uint32_t d = (uint32_t)(xf(kk) >> shift) & 0xffu;
            const uint32_t gi = sm.gbase[d] + li;
            keys_out[gi] = kk;
            vals_out[gi] = sm.vals[li];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
int launch_emit_keys(int P, const GeomLayout& g, const uint2* rects, uint64_t* keys, uint32_t* vals, uint32_t grid_x,
                     uint32_t capacity, cudaStream_t s) {
    if (P <= 0) return HGS_OK;
    StageScope prof(HGS_STAGE_EMIT_KEYS, s);
    emit_keys_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, rects, g.depths, g.offsets, g.tiles_touched, keys, vals,
                                                      grid_x, capacity);
    return check_cuda(cudaGetLastError(), "emit_keys launch");
}

// Sorts n pairs on key bits [0, end_bit).  keys[0]/vals[0] hold the input; returns in *result_buf
// which ping-pong buffer (0/1) holds the sorted output.
// range_hdr / depth_bits: depth-range compaction for the rasterizer's keys (see KeyXform); nullptr / 32 sorts the keys as
// they are.  end_bit counts bits of the COMPACTED key.  start_buf: which ping-pong buffer holds the input.
int launch_sort_pairs(int64_t n, const uint32_t* n_ptr, int end_bit, uint64_t* keys[2], uint32_t* vals[2], void* sort_ws,
                      int* result_buf, cudaStream_t s, GeomHeader* range_hdr, int depth_bits, int start_buf) {
    // the sorted data always ends in buffer (passes & 1), also for the trivial sizes
    const int passes = sort_passes(end_bit);
    if (passes > kMaxPasses) { set_error("end_bit %d too large", end_bit); return HGS_ERR_INVALID; }
    *result_buf = (start_buf ^ passes) & 1;
    if (n <= 0) return HGS_OK;
    if ((n == 1 && n_ptr == nullptr) || end_bit <= 0) {
        if (passes & 1) {
            const int a = start_buf & 1, b = a ^ 1;
            if (int e = check_cuda(cudaMemcpyAsync(keys[b], keys[a], (size_t)n * 8, cudaMemcpyDeviceToDevice, s), "copy keys")) return e;
            if (int e = check_cuda(cudaMemcpyAsync(vals[b], vals[a], (size_t)n * 4, cudaMemcpyDeviceToDevice, s), "copy vals")) return e;
        }
        return HGS_OK;
    }
    if (n >= (1ll << 30)) {
        set_error("radix sort supports < 2^30 instances, got %lld", (long long)n);
        return HGS_ERR_OVERFLOW;
    }
    SortLayout L = carve_sort(sort_ws, n);
    // hist, tickets and the group words of all passes are contiguous, then the tile words of the passes that run
    const size_t clear = (size_t)((char*)L.status - (char*)L.hist) + (size_t)passes * L.ntiles * kRadix * 4;
    if (int e = check_cuda(cudaMemsetAsync(L.hist, 0, clear, s), "memset sort ws")) return e;
    static std::atomic<unsigned long long> attr_done{0};  // function attributes are per device
    if (first_call_on_device(attr_done)) {
        if (int e = check_cuda(cudaFuncSetAttribute(onesweep_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    (int)sizeof(OnesweepSmem)), "onesweep smem attr")) return e;
        if (int e = check_cuda(cudaFuncSetAttribute(onesweep_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    (int)sizeof(OnesweepSmem)), "onesweep smem attr")) return e;
    }
    // HGS_SORT_RANK=ballot selects the warp-vote ranking (A/B measurements, profiles/r2_sort.md: both variants are bound
    // by the same pipe on B200 and time within 5 % of each other; match.any issues fewer instructions)
    static const bool use_ballot = [] { const char* e = getenv("HGS_SORT_RANK"); return e != nullptr && strcmp(e, "ballot") == 0; }();
    const uint32_t nn = (uint32_t)n;
    // two blocks per SM: every block ends with up to passes x 256 global atomics on the SAME 1280 words, and same-address
    // atomics serialise in L2 - with 8 blocks per SM that flush, not the key traffic, was most of the kernel
    int64_t hb = (n + 256 * 8 - 1) / (256 * 8);
    const int hblocks = (int)(hb < 148 * 2 ? hb : 148 * 2);
    {
        StageScope prof(HGS_STAGE_SORT_HISTOGRAM, s);
        radix_histogram_kernel<<<hblocks, 256, 0, s>>>(keys[start_buf & 1], nn, n_ptr, passes, L.hist, range_hdr, depth_bits);
    }
    if (int e = check_cuda(cudaGetLastError(), "radix_histogram launch")) return e;
    int cur = start_buf & 1;
    for (int p = 0; p < passes; ++p) {
        StageScope prof(HGS_STAGE_SORT_ONESWEEP, s);
        const int nbits = min(kRadixBits, end_bit - 8 * p);
        auto kern = use_ballot ? onesweep_pass_kernel<true> : onesweep_pass_kernel<false>;
        kern<<<(unsigned)L.ntiles, kSortThreads, sizeof(OnesweepSmem), s>>>(
            keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], nn, n_ptr, 8 * p, L.hist + p * kRadix,
            L.status + (size_t)p * L.ntiles * kRadix, L.group + (size_t)p * L.ngroups * kRadix, L.group_shift, L.tickets + p,
            range_hdr, depth_bits, nbits);
        if (int e = check_cuda(cudaGetLastError(), "onesweep launch")) return e;
        cur ^= 1;
    }
    *result_buf = cur;
    return HGS_OK;
}

}  // namespace hgs
