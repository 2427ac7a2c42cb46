// tilesort.cu — HGS_SORT_TILE: partition the instances by tile, then sort every tile's list in shared memory.
//
// The reference (and HGS_SORT_GLOBAL) sorts ALL N (tile | depth) keys with a 64-bit radix sort: 5-6 passes over the
// instances, each a one-wave latency-bound kernel at cfg3 and bound by the warp-vote/match pipe at any size
// (profiles/r2_sort.md), plus a histogram pass and the record-packing pass.  But the tile id - the upper half of the key - is
// known when an instance is emitted.  So:
//   (preprocess_fwd counts the instances of every LIST = (tile, depth slice): HGS_TILE_SLICES slices per tile, slice =
//    clamp((depth bits - base) >> shift); any base / shift is correct since the slice is monotone in the depth, good ones -
//    hints from an earlier view's depth range - cut the ~9 k-entry lists of the hair tiles into pieces of a few hundred)
//   tile_offsets  one block: exclusive scan of the counters in (tile, slice) order -> list starts and tile ranges
//                 (identifyTileRanges' output falls out of the counts), longest-first tile order for the compositors, and
//                 the work lists of the three size classes of the sort
//   tile_scatter  instances written to their list's range at an atomic cursor (warp-aggregated) as
//                 (depth bits << 32 | Gaussian id), in arbitrary order                                (writes 8 B / instance)
//   tile_sort_pack one block per list (claimed from the class's work list): sorted ascending by that 64-bit word - i.e. by
//                 (depth, id), exactly the order the reference's STABLE radix sort produces, since it emits instances in
//                 id order - with a bitonic network in shared memory; then the sorted keys / point list and the packed
//                 records the compositors stream are written                (reads 8 B, writes 12 + 48..64 B / instance)
// Two passes over the instances instead of seven, no 64-bit keys in flight, no ranking votes.  Outputs are bit-identical to
// the global sort (keys, point list, ranges, records).  Lists longer than HGS_TILE_SORT_MAX do not fit the block's shared
// memory: bit 2 of the overflow word is raised and the caller repeats stage B with HGS_SORT_GLOBAL.
// Replaces duplicateWithKeys / cub::DeviceRadixSort::SortPairs / identifyTileRanges, rasterizer_impl.cu:70-138,300-314.
#include "hgs_common.cuh"

namespace hgs {

// ------------------------------------------------------------------------------------------------
// list starts, tile ranges, processing order, work lists (one block)
// ------------------------------------------------------------------------------------------------
static constexpr int kOffThreads = 1024;
static constexpr int S = HGS_TILE_SLICES;
static constexpr uint32_t kSmallMax = 512, kMediumMax = 4096;   // size classes of the in-tile sort: (0,512], (512,4096], (4096,MAX]

__device__ __forceinline__ int size_class(uint32_t n) { return n <= kSmallMax ? 2 : (n <= kMediumMax ? 1 : 0); }

__global__ void __launch_bounds__(kOffThreads) tile_offsets_kernel(uint32_t tiles, const uint32_t* __restrict__ list_count,
                                                                   uint32_t capacity, uint2* __restrict__ ranges,
                                                                   uint32_t* __restrict__ order, uint32_t* __restrict__ list_start,
                                                                   uint32_t* __restrict__ work, uint32_t* __restrict__ work_count) {
    __shared__ uint32_t s_warp[kOffThreads / 32];
    __shared__ uint32_t s_count[33];
    __shared__ uint32_t s_offset[33];
    __shared__ uint32_t s_carry;
    __shared__ uint32_t s_work[3];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 33) s_count[tid] = 0;
    if (tid < 3) s_work[tid] = 0;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    const size_t wstride = (size_t)tiles * S;
    // blocked scan over the tiles, kOffThreads per round; a thread owns one tile = S consecutive list counters
    for (uint32_t base = 0; base < tiles; base += kOffThreads) {
        const uint32_t t = base + tid;
        uint32_t cnt[S];
        uint32_t c = 0;
        if (t < tiles) {
#pragma unroll
            for (int q = 0; q < S / 4; ++q) {
                const uint4 v = reinterpret_cast<const uint4*>(list_count + (size_t)t * S)[q];
                cnt[4 * q] = v.x; cnt[4 * q + 1] = v.y; cnt[4 * q + 2] = v.z; cnt[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int q = 0; q < S; ++q) c += cnt[q];
        }
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += n;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t wex = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kOffThreads / 32; ++w) {
            const uint32_t x = s_warp[w];
            if ((uint32_t)w < warp) wex += x;
            total += x;
        }
        const uint32_t start = s_carry + wex + incl - c;
        if (t < tiles) {
            // the reference clears the ranges and writes the non-empty tiles only (rasterizer_impl.cu:116-138,310); a
            // binning workspace that is too small (sync-free guess) truncates the lists instead of overrunning it
            const uint2 r = c ? make_uint2(min(start, capacity), min(start + c, capacity)) : make_uint2(0u, 0u);
            ranges[t] = r;
            atomicAdd(&s_count[r.y > r.x ? __clz(r.y - r.x) : 32], 1u);   // same (clamped) length as the ordering pass below
            uint32_t run = start;
#pragma unroll
            for (int q = 0; q < S; ++q) {
                list_start[(size_t)t * S + q] = run;
                if (cnt[q]) {
                    const int cls = size_class(cnt[q]);
                    work[cls * wstride + atomicAdd(&s_work[cls], 1u)] = t * S + q;
                }
                run += cnt[q];
            }
        }
        __syncthreads();
        if (tid == 0) s_carry += total;
        __syncthreads();
    }
    if (tid < 3) work_count[tid] = s_work[tid];
    // longest-list-first order of the TILES for the compositors, bucketed by floor(log2(length))
    if (tid == 0) {
        uint32_t run = 0;
        for (int b = 0; b < 33; ++b) {
            s_offset[b] = run;
            run += s_count[b];
        }
    }
    __syncthreads();
    for (uint32_t t = tid; t < tiles; t += kOffThreads) {
        const uint2 r = ranges[t];
        const uint32_t c = r.y - r.x;
        order[atomicAdd(&s_offset[c ? __clz(c) : 32], 1u)] = t;
    }
}

// ------------------------------------------------------------------------------------------------
// scatter: load-balanced like emit_keys (a block's instances dealt round-robin to its threads)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tile_scatter_kernel(int P, const uint2* __restrict__ rects,
                                                           const float* __restrict__ depths,
                                                           const uint32_t* __restrict__ offsets,
                                                           const uint32_t* __restrict__ touched_arr,
                                                           const uint32_t* __restrict__ list_start,
                                                           uint32_t* __restrict__ cursor, uint64_t* __restrict__ bucket,
                                                           uint32_t grid_x, uint32_t capacity, uint32_t slice_base, int slice_shift) {
    __shared__ uint32_t s_start[256];  // block-local exclusive start of each Gaussian's run
    __shared__ uint32_t s_xy[256];     // xmin | ymin << 16
    __shared__ uint32_t s_w[256];      // rect width in tiles
    __shared__ uint32_t s_depth[256];
    __shared__ uint32_t s_base, s_total;

    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31;
    const int g0 = blockIdx.x * 256;
    const int g = g0 + tid;
    uint32_t touched = 0, incl = 0;
    if (g < P) {
        touched = touched_arr[g];
        incl = offsets[g];
    }
    if (tid == 0) s_base = incl - touched;
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
    // whole warps stay in the loop (the cursor atomics are warp-aggregated: one per distinct list and trip)
    for (uint32_t s0 = tid & ~31u; s0 < total; s0 += 256) {
        const uint32_t s = s0 + lane;
        const bool has = s < total;
        uint32_t list = 0xffffffffu;
        uint64_t word = 0;
        if (has) {
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
            const uint32_t tile = ((xy >> 16) + ky) * grid_x + (xy & 0xffffu) + kx;
            const uint32_t d = s_depth[lo];
            list = tile * (uint32_t)S + depth_slice(d, slice_base, slice_shift);
            word = ((uint64_t)d << 32) | (uint32_t)(g0 + lo);
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, list);
        const int leader = __ffs(peers) - 1;
        uint32_t first = 0;
        if (has && (int)lane == leader) first = list_start[list] + atomicAdd(&cursor[list], (uint32_t)__popc(peers));
        first = __shfl_sync(0xffffffffu, first, leader);
        if (has) {
            const uint32_t slot = first + __popc(peers & ((1u << lane) - 1u));
            if (slot < capacity) bucket[slot] = word;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// per-list sort + write-out + record packing
// ------------------------------------------------------------------------------------------------
// Persistent blocks claim lists from their size class's work list (no empty blocks: most (tile, slice) lists are empty).
// Ascending bitonic network over the list padded with ~0 to a power of two, in shared memory.
template <int kThreads, int CS>
__global__ void __launch_bounds__(kThreads) tile_sort_pack_kernel(
    const uint32_t* __restrict__ work, const uint32_t* __restrict__ work_n, uint32_t* __restrict__ claim,
    const uint32_t* __restrict__ list_count, const uint32_t* __restrict__ list_start, const uint64_t* __restrict__ bucket,
    uint32_t capacity, uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, const float4* __restrict__ rec,
    const float* __restrict__ rgb, uint32_t tiles_x, float4* __restrict__ pk_lo, float4* __restrict__ pk_hi,
    float4* __restrict__ pk_col, uint16_t* __restrict__ pk_mask) {
    extern __shared__ __align__(16) unsigned char tsp_smem[];
    uint64_t* s = reinterpret_cast<uint64_t*>(tsp_smem);
    __shared__ uint32_t s_item;
    const uint32_t tid = threadIdx.x;
    const uint32_t n_work = *work_n;
    while (true) {
        __syncthreads();   // the previous list's shared memory and s_item have been consumed by every thread
        if (tid == 0) s_item = atomicAdd(claim, 1u);
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= n_work) return;
        const uint32_t list = work[item];
        const uint32_t start = list_start[list];
        if (start >= capacity) continue;
        // A binning workspace that is too small (sync-free guess) or a list that is too long: the caller repeats stage B, but
        // what is inside the workspace must still be valid records - the first slots of the list are all filled (the cursor
        // hands out every slot once), so the part that fits is sorted and packed.
        const uint32_t n = min(min(list_count[list], capacity - start), (uint32_t)HGS_TILE_SORT_MAX);
        uint32_t npad = 2;
        while (npad < n) npad <<= 1;
        for (uint32_t i = tid; i < npad; i += kThreads) s[i] = i < n ? bucket[start + i] : ~0ull;
        __syncthreads();
        for (uint32_t k = 2; k <= npad; k <<= 1) {
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t p = tid; p < (npad >> 1); p += kThreads) {
                    const uint32_t i = ((p & ~(j - 1u)) << 1) | (p & (j - 1u));
                    const uint32_t l = i + j;
                    const uint64_t a = s[i], b = s[l];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        s[i] = b;
                        s[l] = a;
                    }
                }
                __syncthreads();
            }
        }
        const uint32_t tile = list / (uint32_t)HGS_TILE_SLICES;
        const uint64_t tile_hi = (uint64_t)tile << 32;
        const uint32_t X0 = (tile % tiles_x) * HGS_TILE, Y0 = (tile / tiles_x) * HGS_TILE;
        for (uint32_t i = tid; i < n; i += kThreads) {
            const uint64_t w = s[i];
            const uint32_t id = (uint32_t)w;
            const size_t o = (size_t)start + i;
            keys_out[o] = tile_hi | (w >> 32);
            vals_out[o] = id;
            const float4 lo = rec[2 * (size_t)id];
            const float4 hi = rec[2 * (size_t)id + 1];
            const float4 c0 = *reinterpret_cast<const float4*>(rgb + (size_t)id * CS);
            pk_lo[o] = lo;
            pk_hi[o] = hi;
            pk_mask[o] = (uint16_t)block_mask16(lo, hi, X0, Y0);
            pk_col[o * (CS / 4)] = c0;
            if (CS > 4) pk_col[o * (CS / 4) + 1] = *reinterpret_cast<const float4*>(rgb + (size_t)id * CS + 4);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
template <int CS>
static int launch_sort_pack_cs(const ImageLayout& im, const GeomLayout& g, const BinningLayout& b, uint64_t* bucket,
                               uint32_t capacity, size_t lists, uint32_t tiles_x, cudaStream_t s) {
    static std::atomic<unsigned long long> attr_done{0};
    if (first_call_on_device(attr_done)) {
        if (int e = check_cuda(cudaFuncSetAttribute(tile_sort_pack_kernel<1024, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    HGS_TILE_SORT_MAX * 8), "tile_sort_pack smem attr")) return e;
    }
    // persistent grids: one 128 KB block per SM for the long lists, 4 x 32 KB and 16 x 4 KB blocks per SM for the others
    {
        StageScope prof(HGS_STAGE_TILE_SORT_PACK, s);
        tile_sort_pack_kernel<1024, CS><<<148, 1024, HGS_TILE_SORT_MAX * 8, s>>>(
            im.work, im.work_count, im.work_count + 4, g.tile_count, im.list_start, bucket, capacity, b.keys[0], b.vals[0], g.rec,
            g.rgb, tiles_x, b.pk_lo, b.pk_hi, b.pk_col, b.pk_mask);
    }
    {
        StageScope prof(HGS_STAGE_TILE_SORT_PACK, s);
        tile_sort_pack_kernel<512, CS><<<148 * 4, 512, kMediumMax * 8, s>>>(
            im.work + lists, im.work_count + 1, im.work_count + 5, g.tile_count, im.list_start, bucket, capacity, b.keys[0],
            b.vals[0], g.rec, g.rgb, tiles_x, b.pk_lo, b.pk_hi, b.pk_col, b.pk_mask);
    }
    {
        StageScope prof(HGS_STAGE_TILE_SORT_PACK, s);
        tile_sort_pack_kernel<128, CS><<<148 * 16, 128, kSmallMax * 8, s>>>(
            im.work + 2 * lists, im.work_count + 2, im.work_count + 6, g.tile_count, im.list_start, bucket, capacity, b.keys[0],
            b.vals[0], g.rec, g.rgb, tiles_x, b.pk_lo, b.pk_hi, b.pk_col, b.pk_mask);
    }
    return check_cuda(cudaGetLastError(), "tile_sort_pack launch");
}

// Everything between stage A and the compositor for HGS_SORT_TILE.  N = capacity of the binning workspace.
int launch_tile_binning(int P, int channels, int64_t N, const GeomLayout& g, const BinningLayout& b, const ImageLayout& im,
                        uint32_t grid_x, uint32_t grid_y, uint32_t slice_base, int slice_shift, cudaStream_t s) {
    const unsigned tiles = grid_x * grid_y;
    const size_t lists = (size_t)tiles * S;
    // cursors and the work / claim counters behind them
    if (int e = check_cuda(cudaMemsetAsync(im.list_cursor, 0, lists * 4 + 8 * 4, s), "memset list cursors")) return e;
    {
        StageScope prof(HGS_STAGE_TILE_OFFSETS, s);
        tile_offsets_kernel<<<1, kOffThreads, 0, s>>>(tiles, g.tile_count, (uint32_t)N, im.ranges, im.tile_order, im.list_start,
                                                      im.work, im.work_count);
        if (int e = check_cuda(cudaGetLastError(), "tile_offsets launch")) return e;
    }
    if (P <= 0 || N <= 0) return HGS_OK;
    {
        StageScope prof(HGS_STAGE_TILE_SCATTER, s);
        tile_scatter_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, g.rects, g.depths, g.offsets, g.tiles_touched, im.list_start,
                                                            im.list_cursor, b.keys[1], grid_x, (uint32_t)N, slice_base, slice_shift);
        if (int e = check_cuda(cudaGetLastError(), "tile_scatter launch")) return e;
    }
    if (color_stride(channels) == 4) return launch_sort_pack_cs<4>(im, g, b, b.keys[1], (uint32_t)N, lists, grid_x, s);
    return launch_sort_pack_cs<8>(im, g, b, b.keys[1], (uint32_t)N, lists, grid_x, s);
}

}  // namespace hgs
