// composite_warp.cu — per-tile alpha compositing, forward and backward, warp-independent edition.
//   composite_fwd<C> replaces renderCUDA            forward.cu:261-374
//   composite_bwd<C> replaces renderCUDABW_*        backward_distwar.cu:450-1014 (all three variants)
//   finalize_sorted  replaces identifyTileRanges    rasterizer_impl.cu:116-138 and additionally packs the
//                    per-instance records in sorted order
// Blending arithmetic follows SURVEY.md App. A exactly (same expressions, IEEE expf/div, no fast-math)
// so that pixels, final_T and n_contrib agree with the reference build bit for bit.
//
// B200-first structure (what differs from the reference; motivated by profiles/r1_composite_v0.md):
//  * after the sort, one streaming kernel writes every instance's record (means2D, conic, opacity, cull extent,
//    colour) in SORTED order: the compositors then read their tile list with perfectly coalesced 128-bit loads
//    instead of three dependent gathers per instance, forward AND backward;
//  * a CTA still covers one 16x16 tile, but its 8 warps never synchronise with each other: each warp owns an
//    8x4 pixel block and walks the tile list by itself in chunks of 32 instances (one per lane), prefetching
//    the next chunk into registers.  No __syncthreads in the hot loop -> no barrier stalls, and early
//    termination is per 32 pixels instead of per 256;
//  * per chunk every lane tests ITS instance's alpha>=1/255 extent against the warp's pixel block; a ballot
//    gives the candidates, which are broadcast through a 1.5 KB per-warp shared-memory slab.  Skipped instances
//    are exactly those for which every pixel of the warp would take the reference's `continue`
//    (power>0 or alpha<1/255), so outputs are unchanged; list positions still advance;
//  * backward: candidates are compacted into a per-warp queue and processed 8 at a time in two phases (lanes = pixels
//    for the recurrence, lanes = candidate x pixel-quarter for the sums), one lane per candidate issues the
//    red.global.add's — see the comment above composite_bwd_kernel.
#include "hgs_common.cuh"

#include <cstdlib>
#include <cstring>

namespace hgs {

// optional per-(tile,warp) work statistics of the forward compositor (debug; set through hgs_debug_set_stats)
__device__ uint4* g_fwd_stats = nullptr;

static bool g_fwd_stats_on = false;  // host mirror: the counting variant of the forward kernel is launched only then

int set_fwd_stats(void* dev_ptr) {
    uint4* p = (uint4*)dev_ptr;
    g_fwd_stats_on = p != nullptr;
    return check_cuda(cudaMemcpyToSymbol(g_fwd_stats, &p, sizeof(p)), "set stats pointer");
}

// ------------------------------------------------------------------------------------------------
// finalize: tile ranges + sorted-order packing
// ------------------------------------------------------------------------------------------------
template <int CS>
__global__ void __launch_bounds__(256) finalize_sorted_kernel(uint32_t cap, const uint32_t* __restrict__ n_ptr,
                                                              const uint64_t* __restrict__ keys,
                                                              const uint32_t* __restrict__ point_list,
                                                              const float4* __restrict__ rec,
                                                              const float* __restrict__ rgb, uint32_t tiles_x,
                                                              uint2* __restrict__ ranges,
                                                              float4* __restrict__ pk_lo, float4* __restrict__ pk_hi,
                                                              float4* __restrict__ pk_col, uint16_t* __restrict__ pk_mask) {
    const uint32_t L = n_ptr ? min(*n_ptr, cap) : cap;
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L) return;
    const uint32_t id = point_list[idx];
    const float4 lo = rec[2 * (size_t)id];
    const float4 hi = rec[2 * (size_t)id + 1];
    const float4 c0 = *reinterpret_cast<const float4*>(rgb + (size_t)id * CS);
    pk_lo[idx] = lo;
    pk_hi[idx] = hi;
    pk_col[(size_t)idx * (CS / 4)] = c0;
    if (CS > 4) pk_col[(size_t)idx * (CS / 4) + 1] = *reinterpret_cast<const float4*>(rgb + (size_t)id * CS + 4);
    // tile ranges (rasterizer_impl.cu:116-138)
    const uint32_t cur = (uint32_t)(keys[idx] >> 32);
    pk_mask[idx] = (uint16_t)block_mask16(lo, hi, (cur % tiles_x) * HGS_TILE, (cur / tiles_x) * HGS_TILE);
    if (idx == 0) {
        ranges[cur].x = 0;
    } else {
        const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
        if (cur != prev) {
            ranges[prev].y = idx;
            ranges[cur].x = idx;
        }
    }
    if (idx == L - 1) ranges[cur].y = L;
}

// Longest-list-first processing order of the tiles (bucketed by floor(log2(length))): the block scheduler hands
// out CTAs in index order, so the few tiles with thousands of instances start first and the tail of the grid is
// made of cheap tiles (profiles/r1_composite.md: SM active cycles min/avg/max 112k/319k/575k without it).
__global__ void __launch_bounds__(1024) tile_order_kernel(uint32_t tiles, const uint2* __restrict__ ranges,
                                                          uint32_t* __restrict__ order) {
    __shared__ uint32_t s_count[33];
    __shared__ uint32_t s_offset[33];
    if (threadIdx.x < 33) s_count[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < tiles; t += blockDim.x) {
        const uint2 r = ranges[t];
        const uint32_t len = r.y - r.x;
        atomicAdd(&s_count[len ? __clz(len) : 32], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int b = 0; b < 33; ++b) {
            s_offset[b] = run;
            run += s_count[b];
        }
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < tiles; t += blockDim.x) {
        const uint2 r = ranges[t];
        const uint32_t len = r.y - r.x;
        order[atomicAdd(&s_offset[len ? __clz(len) : 32], 1u)] = t;
    }
}

template <int C, int CS>
__global__ void __launch_bounds__(256) composite_fwd_kernel(const uint2* __restrict__ ranges,
                                                            const uint32_t* __restrict__ tile_order, int W, int H,
                                                            const float4* __restrict__ pk_lo,
                                                            const float4* __restrict__ pk_hi,
                                                            const float4* __restrict__ pk_col,
                                                            const float* __restrict__ bg_color,
                                                            float* __restrict__ final_T,
                                                            uint32_t* __restrict__ n_contrib,
                                                            float* __restrict__ out_color) {
    __shared__ float4 s_lo[8][32];
    __shared__ float4 s_hi[8][32];
    __shared__ float4 s_col[8][32 * (CS / 4)];

    // <= 4 channels (the reference's RGB pass): colour * alpha * T exactly as forward.cu:357, so pixels stay bit-equal
    // to the reference build; the fused 7-channel pass has no reference counterpart and folds alpha * T first
    // (one multiply instead of C per blend, ~1 ulp per term).
    constexpr bool kExactOrder = C <= 4;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t horizontal_blocks = (W + HGS_TILE - 1) / HGS_TILE;
    const uint32_t tile = tile_order[blockIdx.x];
    const uint32_t X0 = (tile % horizontal_blocks) * HGS_TILE, Y0 = (tile / horizontal_blocks) * HGS_TILE;
    const uint32_t bx = X0 + (warp & 1) * 8, by = Y0 + (warp >> 1) * 4;
    const uint32_t px = bx + (lane & 7), py = by + (lane >> 3);
    const uint32_t pix_id = W * py + px;
    const float2 pixf = make_float2((float)px, (float)py);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    bool done = !inside;
    const float wx0 = (float)bx, wx1 = (float)(bx + 7), wy0 = (float)by, wy1 = (float)(by + 3);
    const uint32_t lt_mask = (1u << lane) - 1u;

    const uint2 range = ranges[tile];
    const int total = (int)(range.y - range.x);
    const int nchunks = (total + 31) >> 5;

    float T = 1.0f;
    uint32_t last_contributor = 0;
    float Cacc[C];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) Cacc[ch] = 0.f;
    uint32_t st_chunks = 0, st_cand = 0, st_blend = 0;

    if (!__all_sync(0xffffffffu, done) && nchunks > 0) {
        float4 nlo = make_float4(0, 0, 0, 0), nhi = make_float4(0, 0, 0, 0), nc0 = make_float4(0, 0, 0, 0),
               nc1 = make_float4(0, 0, 0, 0);
        if ((int)lane < total) {
            const size_t i = (size_t)range.x + lane;
            nlo = pk_lo[i];
            nhi = pk_hi[i];
            nc0 = pk_col[i * (CS / 4)];
            if (CS > 4) nc1 = pk_col[i * (CS / 4) + 1];
        }
        for (int c = 0; c < nchunks; ++c) {
            const float4 lo = nlo, hi = nhi, c0 = nc0, c1 = nc1;
            const bool have = c * 32 + (int)lane < total;
            if (c + 1 < nchunks) {  // prefetch the next chunk while this one is blended
                const int p = (c + 1) * 32 + (int)lane;
                if (p < total) {
                    const size_t i = (size_t)range.x + p;
                    nlo = pk_lo[i];
                    nhi = pk_hi[i];
                    nc0 = pk_col[i * (CS / 4)];
                    if (CS > 4) nc1 = pk_col[i * (CS / 4) + 1];
                }
            }
            // culled only when a comparison is TRUE, so NaNs keep the instance (the reference would evaluate it)
            const bool cand = have && !(lo.x - hi.z > wx1 || lo.x + hi.z < wx0 || lo.y - hi.w > wy1 || lo.y + hi.w < wy0);
            uint32_t bits = __ballot_sync(0xffffffffu, cand);
            st_chunks++;
            st_cand += __popc(bits);
            if (bits) {
                // candidates are written to the slab COMPACTED (list order kept), the lane they came from rides in the
                // slot of the no longer needed cull extent: the blend loop is then a plain counted loop over
                // consecutive slots instead of a find-first-set scan of the ballot
                const uint32_t ncand = __popc(bits);
                if (cand) {
                    const uint32_t slot = __popc(bits & lt_mask);
                    s_lo[warp][slot] = lo;
                    s_hi[warp][slot] = make_float4(hi.x, hi.y, __uint_as_float(lane), 0.f);
                    s_col[warp][slot * (CS / 4)] = c0;
                    if (CS > 4) s_col[warp][slot * (CS / 4) + 1] = c1;
                }
                __syncwarp();
                const uint32_t pos_base = (uint32_t)(c * 32) + 1u;
                // Two candidates per trip: their evaluate parts (distance, power, expf, alpha) do not depend on the
                // transmittance chain, so issuing them together gives every warp two independent expf chains to
                // overlap.
                for (uint32_t j0 = 0; j0 < ncand; j0 += 2) {
                    const bool two = j0 + 1 < ncand;
                    const uint32_t j1 = two ? j0 + 1 : j0;
                    const float4 glo0 = s_lo[warp][j0], ghi0 = s_hi[warp][j0];
                    const float4 glo1 = s_lo[warp][j1], ghi1 = s_hi[warp][j1];
                    const float2 d0 = make_float2(glo0.x - pixf.x, glo0.y - pixf.y);
                    const float2 d1 = make_float2(glo1.x - pixf.x, glo1.y - pixf.y);
                    const float power0 = -0.5f * (glo0.z * d0.x * d0.x + ghi0.x * d0.y * d0.y) - glo0.w * d0.x * d0.y;
                    const float power1 = -0.5f * (glo1.z * d1.x * d1.x + ghi1.x * d1.y * d1.y) - glo1.w * d1.x * d1.y;
                    const float alpha0 = min(0.99f, ghi0.y * exp(power0));
                    const float alpha1 = min(0.99f, ghi1.y * exp(power1));
                    // same skip rules as forward.cu:336-345, evaluated as predicates
                    const bool ok0 = !(power0 > 0.0f) && !(alpha0 < 1.0f / 255.0f);
                    const bool ok1 = two && !(power1 > 0.0f) && !(alpha1 < 1.0f / 255.0f);
                    if (ok0 && !done) {
                        const float test_T = T * (1 - alpha0);
                        if (test_T < 0.0001f) {
                            done = true;
                        } else {
                            const float* col = reinterpret_cast<const float*>(&s_col[warp][j0 * (CS / 4)]);
                            const float w = alpha0 * T;
#pragma unroll
                            for (int ch = 0; ch < C; ++ch) Cacc[ch] += kExactOrder ? col[ch] * alpha0 * T : col[ch] * w;
                            T = test_T;
                            last_contributor = pos_base + __float_as_uint(ghi0.z);
                            st_blend++;
                        }
                    }
                    if (ok1 && !done) {
                        const float test_T = T * (1 - alpha1);
                        if (test_T < 0.0001f) {
                            done = true;
                        } else {
                            const float* col = reinterpret_cast<const float*>(&s_col[warp][j1 * (CS / 4)]);
                            const float w = alpha1 * T;
#pragma unroll
                            for (int ch = 0; ch < C; ++ch) Cacc[ch] += kExactOrder ? col[ch] * alpha1 * T : col[ch] * w;
                            T = test_T;
                            last_contributor = pos_base + __float_as_uint(ghi1.z);
                            st_blend++;
                        }
                    }
                }
                __syncwarp();
                if (__all_sync(0xffffffffu, done)) break;
            }
        }
    }

    if (inside) {
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) out_color[(size_t)ch * H * W + pix_id] = Cacc[ch] + T * bg_color[ch];
    }
    if (g_fwd_stats != nullptr) {
        st_blend = __reduce_add_sync(0xffffffffu, st_blend);
        const uint32_t ndone = __popc(__ballot_sync(0xffffffffu, done && inside));
        if (lane == 0)
            g_fwd_stats[tile * 8 + warp] = make_uint4(st_chunks, st_cand, st_blend, ndone | ((uint32_t)nchunks << 8));
    }
}

// Forward compositor, half-warp edition.  Same walk as composite_fwd_kernel, but the warp's 8x4 pixel block is treated
// as TWO 4x4 blocks (lanes 0-15 / 16-31): every chunk is culled against both, each half owns a small ring of compacted
// candidates, and in one trip of the blend loop the two halves work on DIFFERENT instances.  Hair Gaussians reach ~2 px
// from their centre, so a 4x4 block rejects many instances that the 8x4 block must keep: on cfg3 the blend loop runs
// 34 % fewer trips (CPU model of the cull and of the queue policy: tools/block_shape_study.py).  Candidates carry over
// from chunk to chunk, so the trips are always full pairs; the rings are drained in step as long as both halves have
// work, and one-sided only when a ring could not take the next chunk.  A pixel still sees exactly the instances, in
// list order, that can reach it, so pixels / final_T / n_contrib are unchanged bit for bit.
template <int CS>
struct FwdHalfSmem {
    static constexpr int QN = 40;  // ring capacity per half: 32 new + <= 8 carried over
    float4 q_lo[2][QN];
    float4 q_hi[2][QN];            // (conic.c, opacity, list position + 1 as bits, -)
    float4 q_col[2][QN * (CS / 4)];
};

// Per-warp landing zone of the bulk-copy staged walk (kBulk): two stages of one 32-instance chunk each.  The tile's list is
// contiguous in the three sorted-order record arrays (finalize_sorted), so a chunk is three 1-D bulk copies
// (cp.async.bulk, the TMA engine: 512 B + 512 B + 32*CS*4 B) completed through the stage's mbarrier; the reference stages
// its batches with plain loads and two block barriers per batch (forward.cu:294-326).
template <int CS>
struct __align__(128) FwdStage {
    float4 lo[32];
    float4 hi[32];
    float4 col[32 * (CS / 4)];
};
static constexpr int kFwdStages = 4;  // chunks in flight per warp: the engine's latency for a 2 KB copy is ~2 chunk times
template <int CS>
struct __align__(128) FwdBulkSmem {
    FwdStage<CS> stage[kFwdStages];
    uint64_t bar[kFwdStages];
};

template <int C, int CS, bool kStats, bool kBulk, bool kMask>
__global__ void __launch_bounds__(256) composite_fwd_half_kernel(const uint2* __restrict__ ranges,
                                                                 const uint32_t* __restrict__ tile_order, int W, int H,
                                                                 const float4* __restrict__ pk_lo,
                                                                 const float4* __restrict__ pk_hi,
                                                                 const float4* __restrict__ pk_col,
                                                                 const uint16_t* __restrict__ pk_mask,
                                                                 const float* __restrict__ bg_color,
                                                                 float* __restrict__ final_T,
                                                                 uint32_t* __restrict__ n_contrib,
                                                                 float* __restrict__ out_color) {
    using WS = FwdHalfSmem<CS>;
    constexpr uint32_t QN = WS::QN;
    __shared__ WS s_ws[8];
    extern __shared__ __align__(128) unsigned char fwd_dyn_smem[];   // kBulk: FwdBulkSmem<CS>[8]

    constexpr bool kExactOrder = C <= 4;  // see composite_fwd_kernel
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t half = lane >> 4, hl = lane & 15;
    WS& ws = s_ws[warp];
    FwdBulkSmem<CS>* bs = kBulk ? &reinterpret_cast<FwdBulkSmem<CS>*>(fwd_dyn_smem)[warp] : nullptr;
    if (kBulk) {
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < kFwdStages; ++k) mbar_init(&bs->bar[k], 1);
            mbar_fence_init();
        }
        __syncwarp();
    }
    const uint32_t horizontal_blocks = (W + HGS_TILE - 1) / HGS_TILE;
    const uint32_t tile = tile_order[blockIdx.x];
    const uint32_t X0 = (tile % horizontal_blocks) * HGS_TILE, Y0 = (tile / horizontal_blocks) * HGS_TILE;
    const uint32_t bx = X0 + (warp & 1) * 8, by = Y0 + (warp >> 1) * 4;
    // lane = half * 16 + row * 4 + column: a warp-wide store still covers 4 rows x 32 contiguous bytes
    const uint32_t px = bx + half * 4 + (hl & 3), py = by + (hl >> 2);
    const uint32_t pix_id = W * py + px;
    const float2 pixf = make_float2((float)px, (float)py);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    bool done = !inside;
    const float ax0 = (float)bx, ax1 = (float)(bx + 3), bx0 = (float)(bx + 4), bx1 = (float)(bx + 7);
    const float wy0 = (float)by, wy1 = (float)(by + 3);
    const uint32_t lt_mask = (1u << lane) - 1u;

    const uint2 range = ranges[tile];
    const int total = (int)(range.y - range.x);
    const int nchunks = (total + 31) >> 5;

    float T = 1.0f;
    uint32_t last_contributor = 0;
    float Cacc[C];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) Cacc[ch] = 0.f;
    uint32_t st_chunks = 0, st_cand = 0, st_blend = 0;

    const float4* my_lo = ws.q_lo[half];
    const float4* my_hi = ws.q_hi[half];
    const float4* my_col = ws.q_col[half];
    uint32_t head_a = 0, cnt_a = 0, head_b = 0, cnt_b = 0;  // warp-uniform ring state

    // blends the first ta / tb queued candidates of the two halves, max(ta, tb) trips, two candidates per trip
    // (independent expf chains, as in composite_fwd_kernel); a half that has run out re-reads a slot, predicated off
    auto run = [&](const uint32_t ta, const uint32_t tb) {
        const uint32_t my_n = half ? tb : ta;
        const uint32_t my_head = half ? head_b : head_a;
        const uint32_t ntrip = max(ta, tb);
        if (kStats) st_cand += ntrip;
        for (uint32_t j0 = 0; j0 < ntrip; j0 += 2) {
            const bool act0 = j0 < my_n, act1 = j0 + 1 < my_n;
            uint32_t i0 = my_head + (act0 ? j0 : 0u);
            i0 = i0 >= QN ? i0 - QN : i0;
            uint32_t i1 = act1 ? i0 + 1 : i0;
            i1 = i1 >= QN ? i1 - QN : i1;
            const float4 glo0 = my_lo[i0], ghi0 = my_hi[i0];
            const float4 glo1 = my_lo[i1], ghi1 = my_hi[i1];
            const float2 d0 = make_float2(glo0.x - pixf.x, glo0.y - pixf.y);
            const float2 d1 = make_float2(glo1.x - pixf.x, glo1.y - pixf.y);
            const float power0 = -0.5f * (glo0.z * d0.x * d0.x + ghi0.x * d0.y * d0.y) - glo0.w * d0.x * d0.y;
            const float power1 = -0.5f * (glo1.z * d1.x * d1.x + ghi1.x * d1.y * d1.y) - glo1.w * d1.x * d1.y;
            const float alpha0 = min(0.99f, ghi0.y * exp(power0));
            const float alpha1 = min(0.99f, ghi1.y * exp(power1));
            // same skip rules as forward.cu:336-345, evaluated as predicates
            const bool ok0 = act0 && !(power0 > 0.0f) && !(alpha0 < 1.0f / 255.0f);
            const bool ok1 = act1 && !(power1 > 0.0f) && !(alpha1 < 1.0f / 255.0f);
            if (ok0 && !done) {
                const float test_T = T * (1 - alpha0);
                if (test_T < 0.0001f) {
                    done = true;
                } else {
                    const float* col = reinterpret_cast<const float*>(&my_col[i0 * (CS / 4)]);
                    const float w = alpha0 * T;
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) Cacc[ch] += kExactOrder ? col[ch] * alpha0 * T : col[ch] * w;
                    T = test_T;
                    last_contributor = __float_as_uint(ghi0.z);
                    if (kStats) st_blend++;
                }
            }
            if (ok1 && !done) {
                const float test_T = T * (1 - alpha1);
                if (test_T < 0.0001f) {
                    done = true;
                } else {
                    const float* col = reinterpret_cast<const float*>(&my_col[i1 * (CS / 4)]);
                    const float w = alpha1 * T;
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) Cacc[ch] += kExactOrder ? col[ch] * alpha1 * T : col[ch] * w;
                    T = test_T;
                    last_contributor = __float_as_uint(ghi1.z);
                    if (kStats) st_blend++;
                }
            }
        }
        head_a += ta;
        head_a = head_a >= QN ? head_a - QN : head_a;
        cnt_a -= ta;
        head_b += tb;
        head_b = head_b >= QN ? head_b - QN : head_b;
        cnt_b -= tb;
        __syncwarp();  // the slots just consumed may be overwritten by the next push
    };

    if (!__all_sync(0xffffffffu, done) && nchunks > 0) {
        float4 nlo = make_float4(0, 0, 0, 0), nhi = make_float4(0, 0, 0, 0), nc0 = make_float4(0, 0, 0, 0),
               nc1 = make_float4(0, 0, 0, 0);
        uint32_t nmask = 0, nmask2 = 0;
        // kMask: this warp's two blocks in the instance masks (block_mask16): block row = warp / 2, columns 2 (warp & 1) + half
        const uint32_t bit_a = 1u << (4 * (warp >> 1) + 2 * (warp & 1)), bits_ab = 3u * bit_a;
        // kBulk: lane 0 asks the TMA engine for chunk c (three contiguous pieces) into stage c & 1
        auto issue = [&](int c) {
            if (lane == 0) {
                const uint32_t cnt = (uint32_t)min(32, total - c * 32);
                const size_t i = (size_t)range.x + (size_t)c * 32;
                FwdStage<CS>& st = bs->stage[c % kFwdStages];
                uint64_t* bar = &bs->bar[c % kFwdStages];
                mbar_expect_tx(bar, cnt * (32u + 4u * CS));
                bulk_g2s(st.lo, pk_lo + i, cnt * 16u, bar);
                bulk_g2s(st.hi, pk_hi + i, cnt * 16u, bar);
                bulk_g2s(st.col, pk_col + i * (CS / 4), cnt * 4u * CS, bar);
            }
        };
        int issued = 0;  // kBulk: last chunk whose copy was issued (every issued copy must LAND before the warp may exit)
        if (kBulk) {
            for (int k = 0; k < kFwdStages - 1 && k < nchunks; ++k) {
                issue(k);
                issued = k;
            }
        } else if (kMask) {
            // masks run two chunks ahead of the walk, the records of the lanes they select one chunk ahead
            if ((int)lane < total) nmask = pk_mask[(size_t)range.x + lane];
            if (32 + (int)lane < total) nmask2 = pk_mask[(size_t)range.x + 32 + lane];
            if (nmask & bits_ab) {
                const size_t i = (size_t)range.x + lane;
                nlo = pk_lo[i];
                nhi = pk_hi[i];
                nc0 = pk_col[i * (CS / 4)];
                if (CS > 4) nc1 = pk_col[i * (CS / 4) + 1];
            }
        } else if ((int)lane < total) {
            const size_t i = (size_t)range.x + lane;
            nlo = pk_lo[i];
            nhi = pk_hi[i];
            nc0 = pk_col[i * (CS / 4)];
            if (CS > 4) nc1 = pk_col[i * (CS / 4) + 1];
        }
        bool all_done = false;
        for (int c = 0; c < nchunks; ++c) {
            float4 lo, hi, c0, c1 = make_float4(0, 0, 0, 0);
            const bool have = c * 32 + (int)lane < total;
            bool cand_a, cand_b;
            if (kMask) {
                // the chunk's 32 block masks are 64 bytes; only the lanes whose instance reaches one of the warp's two blocks
                // read its record (prefetched while the previous chunk was blended, like the masks of the chunk after)
                const uint32_t m = nmask;
                lo = nlo; hi = nhi; c0 = nc0; c1 = nc1;
                nmask = nmask2;
                nmask2 = 0;
                if (c + 2 < nchunks) {
                    const int p = (c + 2) * 32 + (int)lane;
                    if (p < total) nmask2 = pk_mask[(size_t)range.x + p];
                }
                if (nmask & bits_ab) {
                    const size_t i = (size_t)range.x + (size_t)((c + 1) * 32) + lane;
                    nlo = pk_lo[i];
                    nhi = pk_hi[i];
                    nc0 = pk_col[i * (CS / 4)];
                    if (CS > 4) nc1 = pk_col[i * (CS / 4) + 1];
                }
                cand_a = (m & bit_a) != 0;
                cand_b = (m & (bit_a << 1)) != 0;
            } else if (kBulk) {
                // the stage of chunk c + kFwdStages - 1 held chunk c - 1, which every lane has copied out (the __syncwarp
                // below orders those reads before the engine's writes)
                if (c + kFwdStages - 1 < nchunks) {
                    issue(c + kFwdStages - 1);
                    issued = c + kFwdStages - 1;
                }
                mbar_wait(&bs->bar[c % kFwdStages], (uint32_t)(c / kFwdStages) & 1u);
                const FwdStage<CS>& st = bs->stage[c % kFwdStages];
                lo = st.lo[lane];
                hi = st.hi[lane];
                c0 = st.col[lane * (CS / 4)];
                if (CS > 4) c1 = st.col[lane * (CS / 4) + 1];
                __syncwarp();
            } else {
                lo = nlo; hi = nhi; c0 = nc0; c1 = nc1;
                if (c + 1 < nchunks) {  // prefetch the next chunk while this one is blended
                    const int p = (c + 1) * 32 + (int)lane;
                    if (p < total) {
                        const size_t i = (size_t)range.x + p;
                        nlo = pk_lo[i];
                        nhi = pk_hi[i];
                        nc0 = pk_col[i * (CS / 4)];
                        if (CS > 4) nc1 = pk_col[i * (CS / 4) + 1];
                    }
                }
            }
            if (!kMask) {
                // culled only when a comparison is TRUE, so NaNs keep the instance (the reference would evaluate it)
                const float xl = lo.x - hi.z, xr = lo.x + hi.z;
                const bool in_rows = have && !(lo.y - hi.w > wy1 || lo.y + hi.w < wy0);
                cand_a = in_rows && !(xl > ax1 || xr < ax0);
                cand_b = in_rows && !(xl > bx1 || xr < bx0);
            }
            const uint32_t bits_a = __ballot_sync(0xffffffffu, cand_a);
            const uint32_t bits_b = __ballot_sync(0xffffffffu, cand_b);
            if (kStats) st_chunks++;
            if (!(bits_a | bits_b)) continue;
            // push: each half's candidates go to ITS ring compacted (list order kept); the list position rides in the
            // slot of the no longer needed cull extent
            const float4 hi_tag = make_float4(hi.x, hi.y, __uint_as_float((uint32_t)(c * 32) + lane + 1u), 0.f);
            if (cand_a) {
                uint32_t slot = head_a + cnt_a + __popc(bits_a & lt_mask);
                slot = slot >= QN ? slot - QN : slot;
                ws.q_lo[0][slot] = lo;
                ws.q_hi[0][slot] = hi_tag;
                ws.q_col[0][slot * (CS / 4)] = c0;
                if (CS > 4) ws.q_col[0][slot * (CS / 4) + 1] = c1;
            }
            if (cand_b) {
                uint32_t slot = head_b + cnt_b + __popc(bits_b & lt_mask);
                slot = slot >= QN ? slot - QN : slot;
                ws.q_lo[1][slot] = lo;
                ws.q_hi[1][slot] = hi_tag;
                ws.q_col[1][slot * (CS / 4)] = c0;
                if (CS > 4) ws.q_col[1][slot * (CS / 4) + 1] = c1;
            }
            cnt_a += __popc(bits_a);
            cnt_b += __popc(bits_b);
            __syncwarp();
            // drain in step while both halves have a pair; one-sided only as far as the next push needs room
            const uint32_t both = min(cnt_a, cnt_b) & ~1u;
            if (both) run(both, both);
            const uint32_t hi_cnt = max(cnt_a, cnt_b);
            if (hi_cnt > QN - 32) {
                const uint32_t over = (hi_cnt - (QN - 32) + 1u) & ~1u;
                run(min(cnt_a, over), min(cnt_b, over));
            }
            if (__all_sync(0xffffffffu, done)) {
                all_done = true;
                // an in-flight bulk copy targets this CTA's shared memory: it must land before the block can retire
                if (kBulk)
                    for (int k = c + 1; k <= issued; ++k) mbar_wait(&bs->bar[k % kFwdStages], (uint32_t)(k / kFwdStages) & 1u);
                break;
            }
        }
        if (!all_done && (cnt_a | cnt_b)) run(cnt_a, cnt_b);
    }

    if (inside) {
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) out_color[(size_t)ch * H * W + pix_id] = Cacc[ch] + T * bg_color[ch];
    }
    if (kStats && g_fwd_stats != nullptr) {
        st_blend = __reduce_add_sync(0xffffffffu, st_blend);
        const uint32_t ndone = __popc(__ballot_sync(0xffffffffu, done && inside));
        if (lane == 0)
            g_fwd_stats[tile * 8 + warp] = make_uint4(st_chunks, st_cand, st_blend, ndone | ((uint32_t)nchunks << 8));
    }
}

// Backward compositor.  Per warp (8x4 pixels), back to front over the tile list:
//   1. chunks of 32 instances are culled against the warp's pixel block exactly like the forward; surviving
//      instances are COMPACTED into a small per-warp queue (shared memory ring), list order preserved;
//   2. whenever 8 candidates are queued:
//      phase 1 (lanes = pixels): the per-pixel recurrence of backward_distwar.cu:855-1014 runs over the 8 candidates
//               in order; each contributing pixel parks (G, dL/dalpha, alpha*T) for (candidate, pixel) in a slab;
//      phase 2 (lanes = candidate x pixel-quarter): every lane sums ITS candidate's nine gradient terms over the
//               contributing pixels of its quarter in registers, two shuffle levels fold the four quarters, and
//               one lane per candidate issues the red.global.adds.
//      Compared with reducing every candidate across the 32 pixel lanes (16 shuffles + ~50 ALU ops each), this
//      cuts the instruction count per candidate roughly in half and the atomics per warp 4x
//      (profiles/r1_composite.md).
template <int QN>
__device__ __forceinline__ uint32_t qwrap(uint32_t x) { return x >= (uint32_t)QN ? x - QN : x; }

template <int CS>
struct BwdWarpSmem {
    static constexpr int QN = 40;  // candidate queue capacity: 32 new + < 8 left over (ring, wrapped by subtraction)
    static constexpr int GR = 8;   // candidates per group
    float4 q_lo[QN];
    float4 q_hi[QN];
    float4 q_col[QN * (CS / 4)];
    // q_hi.z / q_hi.w carry (list position, Gaussian id) as bit patterns: the cull extents that live there in the
    // packed records are not needed once an instance is queued
    float4 slab[GR * 33];          // [candidate][pixel] (G, dL/dalpha, alpha*T, -), row stride 33 -> conflict-free
    float4 dpix[32 * (CS / 4)];    // dL/dpixel of the warp's 32 pixels
    uint32_t vmask[GR];            // which pixels contributed to each candidate
};

// 3 resident CTAs per SM (80 registers): forcing 4 (64 registers; the shared memory would fit) spills 56 bytes per
// thread in the recurrence and is 45 % slower (0.47 vs 0.32 ms on cfg3).
template <int C, int CS>
__global__ void __launch_bounds__(256, 3) composite_bwd_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order,
    const uint32_t* __restrict__ point_list, int W, int H,
    const float* __restrict__ bg_color, const float4* __restrict__ pk_lo, const float4* __restrict__ pk_hi,
    const float4* __restrict__ pk_col, const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib,
    const float* __restrict__ dL_dpixels, float* __restrict__ dL_dmean2D /*[P,3]*/,
    float* __restrict__ dL_dconic /*[P,4]*/, float* __restrict__ dL_dopacity /*[P]*/,
    float* __restrict__ dL_dcolors /*[P,C]*/) {
    using WS = BwdWarpSmem<CS>;
    constexpr int QN = WS::QN, GR = WS::GR;
    extern __shared__ __align__(16) unsigned char bwd_smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WS& ws = reinterpret_cast<WS*>(bwd_smem_raw)[warp];

    const uint32_t horizontal_blocks = (W + HGS_TILE - 1) / HGS_TILE;
    const uint32_t tile = tile_order[blockIdx.x];
    const uint32_t X0 = (tile % horizontal_blocks) * HGS_TILE, Y0 = (tile / horizontal_blocks) * HGS_TILE;
    const uint32_t bx = X0 + (warp & 1) * 8, by = Y0 + (warp >> 1) * 4;
    const uint32_t px = bx + (lane & 7), py = by + (lane >> 3);
    const uint32_t pix_id = W * py + px;
    const float2 pixf = make_float2((float)px, (float)py);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const float wx0 = (float)bx, wx1 = (float)(bx + 7), wy0 = (float)by, wy1 = (float)(by + 3);

    const uint2 range = ranges[tile];

    const float T_final = inside ? final_Ts[pix_id] : 0;
    float T = T_final;
    const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;

    // nothing at list position >= max over the warp of last_contributor can contribute to these 32 pixels
    int total = last_contributor;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) total = max(total, __shfl_xor_sync(0xffffffffu, total, o));
    if (total == 0) return;
    const int nchunks = (total + 31) >> 5;

    float dL_dpixel[C];
    float acc_dot = 0.f;  // <colour accumulated behind the current instance, dL/dpixel>
#pragma unroll
    for (int ch = 0; ch < C; ++ch) dL_dpixel[ch] = inside ? dL_dpixels[(size_t)ch * H * W + pix_id] : 0.f;
    {
        float tmp[8];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) tmp[ch] = ch < C ? dL_dpixel[ch] : 0.f;
        ws.dpix[lane * (CS / 4)] = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
        if (CS > 4) ws.dpix[lane * (CS / 4) + 1] = make_float4(tmp[4], tmp[5], tmp[6], tmp[7]);
    }
    float bg_dot_dpixel = 0;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) bg_dot_dpixel += bg_color[ch] * dL_dpixel[ch];
    const float ddelx_dx = 0.5 * W;
    const float ddely_dy = 0.5 * H;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // phase-2 role of this lane
    const uint32_t my_g = lane & 7, my_q = lane >> 3;

    uint32_t qhead = 0, qcount = 0;

    auto process_group = [&](const uint32_t n) {
        // ---- phase 1: lanes = pixels, candidates in list order ------------------------------------------
        // evaluate part of one candidate (distance, power, expf, alpha, validity); independent of the recurrence
        auto evaluate = [&](uint32_t slot, float& G, float& alpha) -> bool {
            const float4 glo = ws.q_lo[slot];
            const float4 ghi = ws.q_hi[slot];
            const int pos = __float_as_int(ghi.z);
            bool valid = pos < last_contributor;  // false for outside pixels (last_contributor == 0)
            const float2 d = make_float2(glo.x - pixf.x, glo.y - pixf.y);
            const float power = -0.5f * (glo.z * d.x * d.x + ghi.x * d.y * d.y) - glo.w * d.x * d.y;
            if (power > 0.0f) valid = false;
            G = exp(power);
            alpha = min(0.99f, ghi.y * G);
            if (alpha < 1.0f / 255.0f) valid = false;
            return valid;
        };
        // recurrence step of backward_distwar.cu:960-991 for one contributing (pixel, candidate) pair
        auto recur = [&](uint32_t g, uint32_t slot, float G, float alpha) {
            // one reciprocal serves T/(1-alpha) and T_final/(1-alpha) (the reference divides twice)
            const float om = 1.f - alpha;
            const float rcp = 1.f / om;
            T = T * rcp;
            const float dchannel_dcolor = alpha * T;
            // The reference keeps one running colour per channel (accum_rec, :972-977) and sums
            // (c - accum_rec) * dL/dpixel over the channels.  dL/dpixel is fixed for the pixel, so the same quantity is
            // carried as ONE scalar: with d = <c, dL/dpixel>, acc_dot = <accum_rec, dL/dpixel> obeys the same linear
            // recurrence acc_dot' = alpha * d + (1 - alpha) * acc_dot.  C FMAs + 3 ops instead of 4 C.
            const float* col = reinterpret_cast<const float*>(&ws.q_col[slot * (CS / 4)]);
            float d = 0.f;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) d += col[ch] * dL_dpixel[ch];
            float dL_dalpha = d - acc_dot;
            // colour accumulated behind the NEXT (nearer) instance; the reference performs this update lazily at the
            // start of the next iteration from (last_alpha, last_color)
            acc_dot = alpha * d + om * acc_dot;
            dL_dalpha *= T;
            dL_dalpha += (-T_final * rcp) * bg_dot_dpixel;
            ws.slab[g * 33 + lane] = make_float4(G, dL_dalpha, dchannel_dcolor, 0.f);
        };
        // two candidates per trip: both evaluate parts are issued before either recurrence step (two independent
        // expf chains per warp to overlap, as in the forward)
        for (uint32_t g = 0; g < n; g += 2) {
            const bool two = g + 1 < n;
            const uint32_t slot0 = qwrap<QN>(qhead + g), slot1 = two ? qwrap<QN>(qhead + g + 1) : slot0;
            float G0, a0, G1, a1;
            const bool v0 = evaluate(slot0, G0, a0);
            const bool v1 = evaluate(slot1, G1, a1) && two;
            const uint32_t vm0 = __ballot_sync(0xffffffffu, v0);
            const uint32_t vm1 = __ballot_sync(0xffffffffu, v1);
            if (lane == 0) {
                ws.vmask[g] = vm0;
                if (two) ws.vmask[g + 1] = vm1;
            }
            if (v0) recur(g, slot0, G0, a0);
            if (v1) recur(g + 1, slot1, G1, a1);
        }
        __syncwarp();
        // ---- phase 2: lanes = (candidate my_g, pixel quarter my_q) ---------------------------------------
        float acc[6 + C];
#pragma unroll
        for (int k = 0; k < 6 + C; ++k) acc[k] = 0.f;
        uint32_t my_vm = 0;
        const uint32_t slot = qwrap<QN>(qhead + my_g);
        if (my_g < n) {
            my_vm = ws.vmask[my_g];
            uint32_t m = (my_vm >> (my_q * 8)) & 0xffu;
            if (m) {
                const float4 glo = ws.q_lo[slot];
                const float4 ghi = ws.q_hi[slot];
                while (m) {
                    const uint32_t p = my_q * 8 + (uint32_t)__ffs(m) - 1u;
                    m &= m - 1;
                    const float4 r = ws.slab[my_g * 33 + p];  // (G, dL_dalpha, alpha*T)
                    const float dx = glo.x - (wx0 + (float)(p & 7)), dy = glo.y - (wy0 + (float)(p >> 3));
                    const float dL_dG = ghi.y * r.y;
                    const float gdx = r.x * dx, gdy = r.x * dy;
                    acc[0] += dL_dG * (-gdx * glo.z - gdy * glo.w);
                    acc[1] += dL_dG * (-gdy * ghi.x - gdx * glo.w);
                    acc[2] += gdx * dx * dL_dG;
                    acc[3] += gdx * dy * dL_dG;
                    acc[4] += gdy * dy * dL_dG;
                    acc[5] += r.x * r.y;
                    const float* dp = reinterpret_cast<const float*>(&ws.dpix[p * (CS / 4)]);
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) acc[6 + ch] += r.z * dp[ch];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 6 + C; ++k) {
            acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 8);
            acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 16);
        }
        if (my_q == 0 && my_vm != 0) {
            const uint32_t gid = __float_as_uint(ws.q_hi[slot].w);
            atomicAdd(dL_dmean2D + 3 * (size_t)gid, acc[0] * ddelx_dx);
            atomicAdd(dL_dmean2D + 3 * (size_t)gid + 1, acc[1] * ddely_dy);
            atomicAdd(dL_dconic + 4 * (size_t)gid, -0.5f * acc[2]);
            atomicAdd(dL_dconic + 4 * (size_t)gid + 1, -0.5f * acc[3]);
            atomicAdd(dL_dconic + 4 * (size_t)gid + 3, -0.5f * acc[4]);
            atomicAdd(dL_dopacity + gid, acc[5]);
#pragma unroll
            for (int ch = 0; ch < C; ++ch) atomicAdd(dL_dcolors + (size_t)gid * C + ch, acc[6 + ch]);
        }
        __syncwarp();
    };

    // back to front: lane l of chunk c holds list position total-1-(c*32+l)
    float4 nlo = make_float4(0, 0, 0, 0), nhi = make_float4(0, 0, 0, 0), nc0 = make_float4(0, 0, 0, 0),
           nc1 = make_float4(0, 0, 0, 0);
    uint32_t nid = 0;
    {
        const int pos = total - 1 - (int)lane;
        if (pos >= 0) {
            const size_t i = (size_t)range.x + pos;
            nlo = pk_lo[i];
            nhi = pk_hi[i];
            nc0 = pk_col[i * (CS / 4)];
            if (CS > 4) nc1 = pk_col[i * (CS / 4) + 1];
            nid = point_list[i];
        }
    }
    __syncwarp();
    for (int c = 0; c < nchunks; ++c) {
        const float4 lo = nlo, hi = nhi, c0 = nc0, c1 = nc1;
        const uint32_t id = nid;
        const int my_pos = total - 1 - c * 32 - (int)lane;
        const bool have = my_pos >= 0;
        if (c + 1 < nchunks) {
            const int pos = my_pos - 32;
            if (pos >= 0) {
                const size_t i = (size_t)range.x + pos;
                nlo = pk_lo[i];
                nhi = pk_hi[i];
                nc0 = pk_col[i * (CS / 4)];
                if (CS > 4) nc1 = pk_col[i * (CS / 4) + 1];
                nid = point_list[i];
            }
        }
        const bool cand = have && !(lo.x - hi.z > wx1 || lo.x + hi.z < wx0 || lo.y - hi.w > wy1 || lo.y + hi.w < wy0);
        const uint32_t bits = __ballot_sync(0xffffffffu, cand);
        if (!bits) continue;
        if (cand) {
            const uint32_t slot = qwrap<QN>(qhead + qcount + __popc(bits & lt_mask));
            ws.q_lo[slot] = lo;
            ws.q_hi[slot] = make_float4(hi.x, hi.y, __int_as_float(my_pos), __uint_as_float(id));
            ws.q_col[slot * (CS / 4)] = c0;
            if (CS > 4) ws.q_col[slot * (CS / 4) + 1] = c1;
        }
        qcount += __popc(bits);
        __syncwarp();
        while (qcount >= (uint32_t)GR) {
            process_group(GR);
            qhead = qwrap<QN>(qhead + GR);
            qcount -= GR;
        }
    }
    if (qcount > 0) process_group(qcount);
}

// Backward compositor, half-warp edition (see composite_fwd_half_kernel for the motivation): the warp's 8x4 pixels are
// two 4x4 blocks with their own candidate rings.  A group is processed when BOTH rings hold 8 candidates (or one of
// them could not take the next chunk):
//   phase 1 (lanes = pixels): each half runs the recurrence over ITS up to 8 candidates, the two halves working on
//            different instances in the same trip; (G, dL/dalpha, alpha*T) is parked per (half, candidate, pixel);
//   phase 2 (lanes = half x candidate x column parity): every lane sums its candidate's gradient terms over the
//            contributing pixels of its two columns, ONE shuffle level folds the two parities, and 16 lanes (one per queued
//            candidate of either half) issue the red.global.adds.
// cfg3, CPU model of the queue policy (tools/block_shape_study.py): 26 % fewer phase-1 trips than the 8x4 kernel; an
// instance that reaches both halves is accumulated by both (1.3x the candidates, each over 16 instead of 32 pixels).
template <int CS>
struct BwdHalfSmem {
    static constexpr int QN = 40;   // ring capacity per half: 32 new + <= 8 carried over
    static constexpr int GR = 8;    // candidates per half and group
    float4 q_lo[2][QN];
    float4 q_hi[2][QN];             // (conic.c, opacity, list position as bits, Gaussian id as bits)
    float4 q_col[2][QN * (CS / 4)];
    float2 slab_gd[GR * 33];        // [candidate slot g][lane] (G, dL/dalpha): slot g of half A in lanes 0-15, of half B
    float slab_w[GR * 33];          // in lanes 16-31; alpha * T.  Row stride 33 -> conflict-free in both phases
    float4 dpix[32 * (CS / 4)];     // dL/dpixel of the warp's 32 pixels (indexed by lane)
    uint32_t vmask[GR];             // bits 0-15: pixels of half A that candidate g of half A reached; 16-31: half B
};

// 16-byte vector reduction (sm_90+): one L2 atomic transaction for four floats of one record
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// kVec: gradients accumulate into ONE interleaved 64-byte record per Gaussian (hgs_strand_grads.acc16, passed in
// dL_dmean2D) with four red.global.add.v4.f32 instead of 6 + C scalar atomics.
template <int C, int CS, bool kVec, bool kMask>
__global__ void __launch_bounds__(256, 3) composite_bwd_half_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order,
    const uint32_t* __restrict__ point_list, int W, int H,
    const float* __restrict__ bg_color, const float4* __restrict__ pk_lo, const float4* __restrict__ pk_hi,
    const float4* __restrict__ pk_col, const uint16_t* __restrict__ pk_mask, const float* __restrict__ final_Ts,
    const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpixels, float* __restrict__ dL_dmean2D /*[P,3]*/,
    float* __restrict__ dL_dconic /*[P,4]*/, float* __restrict__ dL_dopacity /*[P]*/,
    float* __restrict__ dL_dcolors /*[P,C]*/) {
    using WS = BwdHalfSmem<CS>;
    constexpr uint32_t QN = WS::QN, GR = WS::GR;
    extern __shared__ __align__(16) unsigned char bwd_smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t half = lane >> 4, hl = lane & 15;
    WS& ws = reinterpret_cast<WS*>(bwd_smem_raw)[warp];

    const uint32_t horizontal_blocks = (W + HGS_TILE - 1) / HGS_TILE;
    const uint32_t tile = tile_order[blockIdx.x];
    const uint32_t X0 = (tile % horizontal_blocks) * HGS_TILE, Y0 = (tile / horizontal_blocks) * HGS_TILE;
    const uint32_t bx = X0 + (warp & 1) * 8, by = Y0 + (warp >> 1) * 4;
    const uint32_t px = bx + half * 4 + (hl & 3), py = by + (hl >> 2);
    const uint32_t pix_id = W * py + px;
    const float2 pixf = make_float2((float)px, (float)py);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const float ax0 = (float)bx, wy0 = (float)by;

    const uint2 range = ranges[tile];

    const float T_final = inside ? final_Ts[pix_id] : 0;
    float T = T_final;
    const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;

    // nothing at list position >= the half's largest last_contributor can contribute to its 16 pixels
    int total_mine = last_contributor;
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) total_mine = max(total_mine, __shfl_xor_sync(0xffffffffu, total_mine, o));
    const int total_other = __shfl_xor_sync(0xffffffffu, total_mine, 16);
    const int total_a = half ? total_other : total_mine, total_b = half ? total_mine : total_other;
    const int total = max(total_a, total_b);
    if (total == 0) return;
    const int nchunks = (total + 31) >> 5;

    float dL_dpixel[C];
    float acc_dot = 0.f;  // <colour accumulated behind the current instance, dL/dpixel>
#pragma unroll
    for (int ch = 0; ch < C; ++ch) dL_dpixel[ch] = inside ? dL_dpixels[(size_t)ch * H * W + pix_id] : 0.f;
    {
        float tmp[8];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) tmp[ch] = ch < C ? dL_dpixel[ch] : 0.f;
        ws.dpix[lane * (CS / 4)] = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
        if (CS > 4) ws.dpix[lane * (CS / 4) + 1] = make_float4(tmp[4], tmp[5], tmp[6], tmp[7]);
    }
    float bg_dot_dpixel = 0;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) bg_dot_dpixel += bg_color[ch] * dL_dpixel[ch];
    const float ddelx_dx = 0.5 * W;
    const float ddely_dy = 0.5 * H;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // phase-2 role of this lane: candidate my_g of half `half`, pixel columns my_o and my_o + 2 of that half
    const uint32_t my_g = lane & 7, my_o = (lane >> 3) & 1;

    const float4* my_lo = ws.q_lo[half];
    const float4* my_hi = ws.q_hi[half];
    const float4* my_col = ws.q_col[half];
    uint32_t head_a = 0, cnt_a = 0, head_b = 0, cnt_b = 0;  // warp-uniform ring state

    auto process_group = [&](const uint32_t na, const uint32_t nb) {
        const uint32_t my_n = half ? nb : na;
        const uint32_t my_head = half ? head_b : head_a;
        const uint32_t ntrip = max(na, nb);
        // ---- phase 1: lanes = pixels, each half over its own candidates in list order ---------------------------
        auto evaluate = [&](uint32_t slot, float& G, float& alpha) -> bool {
            const float4 glo = my_lo[slot];
            const float4 ghi = my_hi[slot];
            const int pos = __float_as_int(ghi.z);
            bool valid = pos < last_contributor;  // false for outside pixels (last_contributor == 0)
            const float2 d = make_float2(glo.x - pixf.x, glo.y - pixf.y);
            const float power = -0.5f * (glo.z * d.x * d.x + ghi.x * d.y * d.y) - glo.w * d.x * d.y;
            if (power > 0.0f) valid = false;
            G = exp(power);
            alpha = min(0.99f, ghi.y * G);
            if (alpha < 1.0f / 255.0f) valid = false;
            return valid;
        };
        // recurrence step of backward_distwar.cu:960-991 for one contributing (pixel, candidate) pair; see
        // composite_bwd_kernel for the scalar form of the colour recurrence
        auto recur = [&](uint32_t g, uint32_t slot, float G, float alpha) {
            const float om = 1.f - alpha;
            // om lies in [0.01, 1]: MUFU.RCP (<= 1 ulp) instead of the IEEE division sequence (8 instructions with its range
            // check and slow-path call); the reference divides T by om (backward_distwar.cu:963), either way T is a
            // reconstruction good to ~1 ulp per step
            const float rcp = rcp_approx(om);
            T = T * rcp;
            const float dchannel_dcolor = alpha * T;
            const float* col = reinterpret_cast<const float*>(&my_col[slot * (CS / 4)]);
            float d = 0.f;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) d += col[ch] * dL_dpixel[ch];
            float dL_dalpha = d - acc_dot;
            acc_dot = alpha * d + om * acc_dot;
            dL_dalpha *= T;
            dL_dalpha += (-T_final * rcp) * bg_dot_dpixel;
            const uint32_t si = g * 33 + lane;
            ws.slab_gd[si] = make_float2(G, dL_dalpha);
            ws.slab_w[si] = dchannel_dcolor;
        };
        for (uint32_t g = 0; g < ntrip; g += 2) {
            const bool act0 = g < my_n, act1 = g + 1 < my_n;
            uint32_t slot0 = my_head + (act0 ? g : 0u);
            slot0 = slot0 >= QN ? slot0 - QN : slot0;
            uint32_t slot1 = act1 ? slot0 + 1 : slot0;
            slot1 = slot1 >= QN ? slot1 - QN : slot1;
            float G0, a0, G1, a1;
            const bool v0 = evaluate(slot0, G0, a0) && act0;
            const bool v1 = evaluate(slot1, G1, a1) && act1;
            const uint32_t vm0 = __ballot_sync(0xffffffffu, v0);
            const uint32_t vm1 = __ballot_sync(0xffffffffu, v1);
            if (lane == 0) {
                ws.vmask[g] = vm0;
                if (g + 1 < GR) ws.vmask[g + 1] = vm1;
            }
            if (v0) recur(g, slot0, G0, a0);
            if (v1) recur(g + 1, slot1, G1, a1);
        }
        __syncwarp();
        // ---- phase 2: lanes = (half, candidate my_g, column parity my_o) ---------------------------------------
        // With s = G * dL/dalpha of a contributing pixel and d = mean2D - pixel, every geometric gradient of the candidate
        // is a combination of six moments of s over its pixels - sum s, s dx, s dy, s dx dx, s dx dy, s dy dy - with
        // coefficients that depend on the candidate only (backward_distwar.cu:993-1006 expanded: dL/dG = opacity * dL/dalpha,
        // dG/ddelx = -G (conic.x dx + conic.y dy), ...): 9 FP operations per pixel instead of 16, the coefficients once per group.
        float acc[6 + C];
#pragma unroll
        for (int k = 0; k < 6 + C; ++k) acc[k] = 0.f;
        uint32_t my_vm = 0;
        uint32_t slot = my_head + my_g;
        slot = slot >= QN ? slot - QN : slot;
        if (my_g < my_n) {
            my_vm = (ws.vmask[my_g] >> (half * 16)) & 0xffffu;
            // the candidate's two lanes take the even / the odd columns of the 4x4 block: a strand covers neighbouring
            // pixels, so the contributing pixels split about evenly (rows 0-1 / 2-3 left 5.9 trips of this loop per
            // group with 7 lanes active, profiles/r1_composite_half.md)
            uint32_t m = my_vm & (0x5555u << my_o);
            if (m) {
                const float2 gxy = *reinterpret_cast<const float2*>(&my_lo[slot]);
                const float4 glo = make_float4(gxy.x, gxy.y, 0.f, 0.f);
                const float hx0 = ax0 + (float)(half * 4);
                while (m) {
                    const uint32_t p = (uint32_t)__ffs(m) - 1u;  // pixel of this half, row-major 4x4
                    m &= m - 1;
                    const uint32_t si = my_g * 33 + half * 16 + p;
                    const float2 r = ws.slab_gd[si];  // (G, dL_dalpha)
                    const float rw = ws.slab_w[si];   // alpha * T
                    const float dx = glo.x - (hx0 + (float)(p & 3)), dy = glo.y - (wy0 + (float)(p >> 2));
                    const float sv = r.x * r.y;
                    const float sx = sv * dx, sy = sv * dy;
                    acc[0] += sv;
                    acc[1] += sx;
                    acc[2] += sy;
                    acc[3] += sx * dx;
                    acc[4] += sx * dy;
                    acc[5] += sy * dy;
                    const float* dp = reinterpret_cast<const float*>(&ws.dpix[(half * 16 + p) * (CS / 4)]);
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) acc[6 + ch] += rw * dp[ch];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 6 + C; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 8);
        if (my_o == 0 && my_vm != 0) {
            const float4 glo = my_lo[slot], ghi = my_hi[slot];   // re-read here: not kept live across the pixel loop
            const uint32_t gid = __float_as_uint(ghi.w);
            const float op = ghi.y;
            // glo = (x, y, conic.x, conic.y), ghi = (conic.z, opacity, position, id)
            const float g_mx = -op * (glo.z * acc[1] + glo.w * acc[2]) * ddelx_dx;
            const float g_my = -op * (ghi.x * acc[2] + glo.w * acc[1]) * ddely_dy;
            const float g_cxx = -0.5f * op * acc[3], g_cxy = -0.5f * op * acc[4], g_cyy = -0.5f * op * acc[5];
            if (kVec) {
                float* rec16 = dL_dmean2D + 16 * (size_t)gid;
                float col[8];
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) col[ch] = ch < C ? acc[6 + ch] : 0.f;
                red_add_v4(rec16, g_mx, g_my, acc[0], 0.f);
                red_add_v4(rec16 + 4, g_cxx, g_cxy, 0.f, g_cyy);
                red_add_v4(rec16 + 8, col[0], col[1], col[2], col[3]);
                if (C > 4) red_add_v4(rec16 + 12, col[4], col[5], col[6], col[7]);
            } else {
                atomicAdd(dL_dmean2D + 3 * (size_t)gid, g_mx);
                atomicAdd(dL_dmean2D + 3 * (size_t)gid + 1, g_my);
                atomicAdd(dL_dconic + 4 * (size_t)gid, g_cxx);
                atomicAdd(dL_dconic + 4 * (size_t)gid + 1, g_cxy);
                atomicAdd(dL_dconic + 4 * (size_t)gid + 3, g_cyy);
                atomicAdd(dL_dopacity + gid, acc[0]);
#pragma unroll
                for (int ch = 0; ch < C; ++ch) atomicAdd(dL_dcolors + (size_t)gid * C + ch, acc[6 + ch]);
            }
        }
        head_a += na;
        head_a = head_a >= QN ? head_a - QN : head_a;
        cnt_a -= na;
        head_b += nb;
        head_b = head_b >= QN ? head_b - QN : head_b;
        cnt_b -= nb;
        __syncwarp();
    };

    // back to front: lane l of chunk c holds list position total-1-(c*32+l)
    float4 nlo = make_float4(0, 0, 0, 0), nhi = make_float4(0, 0, 0, 0), nc0 = make_float4(0, 0, 0, 0),
           nc1 = make_float4(0, 0, 0, 0);
    uint32_t nid = 0, nmask = 0, nmask2 = 0;
    // kMask: this warp's two blocks in the instance masks (block_mask16): block row = warp / 2, columns 2 (warp & 1) + half.
    // Masks run two chunks ahead of the walk, the records of the lanes they select one chunk ahead.
    const uint32_t bit_a = 1u << (4 * (warp >> 1) + 2 * (warp & 1)), bits_ab = 3u * bit_a;
    {
        const int pos = total - 1 - (int)lane;
        if (kMask && pos >= 32) nmask2 = pk_mask[(size_t)range.x + pos - 32];
        if (pos >= 0) {
            const size_t i = (size_t)range.x + pos;
            if (kMask) {
                nmask = pk_mask[i];
                if (nmask & bits_ab) {
                    nlo = pk_lo[i];
                    nhi = pk_hi[i];
                    nc0 = pk_col[i * (CS / 4)];
                    if (CS > 4) nc1 = pk_col[i * (CS / 4) + 1];
                    nid = point_list[i];
                }
            } else {
                nlo = pk_lo[i];
                nhi = pk_hi[i];
                nc0 = pk_col[i * (CS / 4)];
                if (CS > 4) nc1 = pk_col[i * (CS / 4) + 1];
                nid = point_list[i];
            }
        }
    }
    __syncwarp();
    for (int c = 0; c < nchunks; ++c) {
        const float4 lo = nlo, hi = nhi, c0 = nc0, c1 = nc1;
        const uint32_t id = nid;
        const uint32_t m = nmask;
        const int my_pos = total - 1 - c * 32 - (int)lane;
        const bool have = my_pos >= 0;
        nmask = nmask2;
        nmask2 = 0;
        if (kMask && my_pos >= 64) nmask2 = pk_mask[(size_t)range.x + my_pos - 64];
        if (c + 1 < nchunks) {
            const int pos = my_pos - 32;
            if (pos >= 0) {
                const size_t i = (size_t)range.x + pos;
                if (kMask) {
                    if (nmask & bits_ab) {
                        nlo = pk_lo[i];
                        nhi = pk_hi[i];
                        nc0 = pk_col[i * (CS / 4)];
                        if (CS > 4) nc1 = pk_col[i * (CS / 4) + 1];
                        nid = point_list[i];
                    }
                } else {
                    nlo = pk_lo[i];
                    nhi = pk_hi[i];
                    nc0 = pk_col[i * (CS / 4)];
                    if (CS > 4) nc1 = pk_col[i * (CS / 4) + 1];
                    nid = point_list[i];
                }
            }
        }
        bool cand_a, cand_b;
        if (kMask) {
            // (positions before the list start carry mask 0); only the lanes whose instance reaches one of the warp's two
            // blocks go on to read its record
            cand_a = (m & bit_a) != 0 && my_pos < total_a;
            cand_b = (m & (bit_a << 1)) != 0 && my_pos < total_b;
        } else {
            // extents relative to the warp's block origin, so the four block edges are immediates (the rounding of the
            // extra subtraction, <= 2^-22 relative, is far inside the padding of the extents: preprocess.cu alpha_extent)
            const float xl = (lo.x - hi.z) - ax0, xr = (lo.x + hi.z) - ax0;
            const float yt = (lo.y - hi.w) - wy0, yb = (lo.y + hi.w) - wy0;
            const bool in_rows = have && !(yt > 3.f || yb < 0.f);
            cand_a = in_rows && my_pos < total_a && !(xl > 3.f || xr < 0.f);
            cand_b = in_rows && my_pos < total_b && !(xl > 7.f || xr < 4.f);
        }
        const uint32_t bits_a = __ballot_sync(0xffffffffu, cand_a);
        const uint32_t bits_b = __ballot_sync(0xffffffffu, cand_b);
        if (!(bits_a | bits_b)) continue;
        const float4 hi_tag = make_float4(hi.x, hi.y, __int_as_float(my_pos), __uint_as_float(id));
        if (cand_a) {
            uint32_t slot = head_a + cnt_a + __popc(bits_a & lt_mask);
            slot = slot >= QN ? slot - QN : slot;
            ws.q_lo[0][slot] = lo;
            ws.q_hi[0][slot] = hi_tag;
            ws.q_col[0][slot * (CS / 4)] = c0;
            if (CS > 4) ws.q_col[0][slot * (CS / 4) + 1] = c1;
        }
        if (cand_b) {
            uint32_t slot = head_b + cnt_b + __popc(bits_b & lt_mask);
            slot = slot >= QN ? slot - QN : slot;
            ws.q_lo[1][slot] = lo;
            ws.q_hi[1][slot] = hi_tag;
            ws.q_col[1][slot * (CS / 4)] = c0;
            if (CS > 4) ws.q_col[1][slot * (CS / 4) + 1] = c1;
        }
        cnt_a += __popc(bits_a);
        cnt_b += __popc(bits_b);
        __syncwarp();
        while ((cnt_a >= GR && cnt_b >= GR) || cnt_a > QN - 32 || cnt_b > QN - 32)
            process_group(min(cnt_a, GR), min(cnt_b, GR));
    }
    while (cnt_a | cnt_b) process_group(min(cnt_a, GR), min(cnt_b, GR));
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
struct PackedView {
    float4* lo;
    float4* hi;
    float4* col;
    uint16_t* mask;
};

static PackedView packed_view(const BinningLayout& b) { return PackedView{b.pk_lo, b.pk_hi, b.pk_col, b.pk_mask}; }

// How the half-warp compositors find the instances that reach their two 4x4 blocks: by testing every instance's extent per
// warp (default) or, HGS_WALK=mask, from the per-instance block masks written with the sorted records - only the lanes whose
// instance reaches a block then read its 68-byte record.  Both select exactly the same instances.  Measured on cfg3
// (profiles/r2_block_mask.md): the mask walk moves 3.5x fewer bytes through L1 but issues as many warp instructions (a chunk
// nearly always has SOME candidate lane, so the predicated record loads are issued anyway): forward 5 % slower, backward equal.
static bool walk_by_mask() {
    static const bool v = [] { const char* e = getenv("HGS_WALK"); return e != nullptr && strcmp(e, "mask") == 0; }();
    return v;
}

// Pixel-block shape of the compositors: two 4x4 blocks per warp (default) or one 8x4 block per warp (the round-1 kernels,
// kept for A/B measurements: profiles/r1_*).  Both produce identical outputs.  Selected by the environment variable
// HGS_COMPOSITE_BLOCKS=8x4|4x4 or, overriding it, by hgs_debug_set_composite_blocks().
static int g_blocks_override = 0;  // 0: environment / default, 1: 4x4 halves, 2: 8x4

int set_composite_blocks(int mode) {
    if (mode < 0 || mode > 2) {
        set_error("composite block mode %d: expected 0 (default), 1 (4x4 halves) or 2 (8x4)", mode);
        return HGS_ERR_INVALID;
    }
    g_blocks_override = mode;
    return HGS_OK;
}

static bool composite_blocks_4x4() {
    static const bool from_env = [] {
        const char* e = getenv("HGS_COMPOSITE_BLOCKS");
        return !(e != nullptr && strcmp(e, "8x4") == 0);
    }();
    return g_blocks_override ? g_blocks_override == 1 : from_env;
}

int launch_finalize_sorted(int channels, int64_t n, const uint32_t* n_ptr, const uint64_t* keys_sorted,
                           const uint32_t* point_list,
                           const GeomLayout& g, const BinningLayout& b, uint2* ranges, uint32_t* tile_order, size_t tiles,
                           uint32_t tiles_x, cudaStream_t s) {
    if (int e = check_cuda(cudaMemsetAsync(ranges, 0, tiles * sizeof(uint2), s), "memset ranges")) return e;
    StageScope prof(HGS_STAGE_TILE_RANGES, s);
    if (n <= 0) {
        tile_order_kernel<<<1, 1024, 0, s>>>((uint32_t)tiles, ranges, tile_order);
        return check_cuda(cudaGetLastError(), "tile_order launch");
    }
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (color_stride(channels) == 4)
        finalize_sorted_kernel<4><<<nb, 256, 0, s>>>((uint32_t)n, n_ptr, keys_sorted, point_list, g.rec, g.rgb, tiles_x, ranges,
                                                     b.pk_lo, b.pk_hi, b.pk_col, b.pk_mask);
    else
        finalize_sorted_kernel<8><<<nb, 256, 0, s>>>((uint32_t)n, n_ptr, keys_sorted, point_list, g.rec, g.rgb, tiles_x, ranges,
                                                     b.pk_lo, b.pk_hi, b.pk_col, b.pk_mask);
    tile_order_kernel<<<1, 1024, 0, s>>>((uint32_t)tiles, ranges, tile_order);
    return check_cuda(cudaGetLastError(), "finalize_sorted launch");
}

template <int C>
static int launch_fwd_c(const ImageLayout& im, const BinningLayout& b, int W, int H, const float* bg, float* out_color,
                        cudaStream_t s) {
    const unsigned grid = (unsigned)(((W + HGS_TILE - 1) / HGS_TILE) * ((H + HGS_TILE - 1) / HGS_TILE));
    constexpr int CS = (C <= 4) ? 4 : 8;
    const PackedView p = packed_view(b);
    // HGS_FWD_SMEM_PAD=<bytes>: unused dynamic shared memory added to the forward launch to LOWER its occupancy
    // (experiments on co-scheduling with the binning kernels of the next view, profiles/r2_overlap.md)
    static const int pad = [] { const char* e = getenv("HGS_FWD_SMEM_PAD"); return e ? atoi(e) : 0; }();
    // HGS_FWD_STAGING=bulk: chunks staged into shared memory by the TMA engine (cp.async.bulk + mbarrier) instead of the
    // register prefetch (A/B in profiles/r2_fwd_staging.md)
    static const bool bulk = [] { const char* e = getenv("HGS_FWD_STAGING"); return e != nullptr && strcmp(e, "bulk") == 0; }();
    if ((pad > 0 || bulk) && composite_blocks_4x4() && !g_fwd_stats_on) {
        const int dyn = bulk ? (int)(8 * sizeof(FwdBulkSmem<CS>)) + pad : pad;
        static std::atomic<unsigned long long> pad_done{0};
        if (first_call_on_device(pad_done)) {
            if (int e = check_cuda(cudaFuncSetAttribute(composite_fwd_half_kernel<C, CS, false, false, true>,
                                                        cudaFuncAttributeMaxDynamicSharedMemorySize, dyn), "fwd pad attr")) return e;
            if (int e = check_cuda(cudaFuncSetAttribute(composite_fwd_half_kernel<C, CS, false, true, false>,
                                                        cudaFuncAttributeMaxDynamicSharedMemorySize, dyn), "fwd bulk attr")) return e;
        }
        StageScope prof(HGS_STAGE_COMPOSITE_FWD, s);
        if (bulk)
            composite_fwd_half_kernel<C, CS, false, true, false><<<grid, 256, dyn, s>>>(im.ranges, im.tile_order, W, H, p.lo, p.hi,
                                                                                         p.col, p.mask, bg, im.final_T, im.n_contrib,
                                                                                         out_color);
        else
            composite_fwd_half_kernel<C, CS, false, false, true><<<grid, 256, dyn, s>>>(im.ranges, im.tile_order, W, H, p.lo, p.hi,
                                                                                         p.col, p.mask, bg, im.final_T, im.n_contrib,
                                                                                         out_color);
        return check_cuda(cudaGetLastError(), "composite_fwd launch");
    }
    StageScope prof(HGS_STAGE_COMPOSITE_FWD, s);
    if (composite_blocks_4x4() && g_fwd_stats_on)
        composite_fwd_half_kernel<C, CS, true, false, false><<<grid, 256, 0, s>>>(im.ranges, im.tile_order, W, H, p.lo, p.hi, p.col,
                                                                                   p.mask, bg, im.final_T, im.n_contrib, out_color);
    else if (composite_blocks_4x4() && walk_by_mask())
        composite_fwd_half_kernel<C, CS, false, false, true><<<grid, 256, 0, s>>>(im.ranges, im.tile_order, W, H, p.lo, p.hi, p.col,
                                                                                   p.mask, bg, im.final_T, im.n_contrib, out_color);
    else if (composite_blocks_4x4())
        composite_fwd_half_kernel<C, CS, false, false, false><<<grid, 256, 0, s>>>(im.ranges, im.tile_order, W, H, p.lo, p.hi, p.col,
                                                                                    p.mask, bg, im.final_T, im.n_contrib, out_color);
    else
        composite_fwd_kernel<C, CS><<<grid, 256, 0, s>>>(im.ranges, im.tile_order, W, H, p.lo, p.hi, p.col, bg, im.final_T,
                                                          im.n_contrib, out_color);
    return check_cuda(cudaGetLastError(), "composite_fwd launch");
}

int launch_composite_fwd(int channels, const ImageLayout& im, const BinningLayout& b, int W, int H, const float* bg,
                         float* out_color, cudaStream_t s) {
    switch (channels) {
        case 1: return launch_fwd_c<1>(im, b, W, H, bg, out_color, s);
        case 2: return launch_fwd_c<2>(im, b, W, H, bg, out_color, s);
        case 3: return launch_fwd_c<3>(im, b, W, H, bg, out_color, s);
        case 4: return launch_fwd_c<4>(im, b, W, H, bg, out_color, s);
        case 5: return launch_fwd_c<5>(im, b, W, H, bg, out_color, s);
        case 6: return launch_fwd_c<6>(im, b, W, H, bg, out_color, s);
        case 7: return launch_fwd_c<7>(im, b, W, H, bg, out_color, s);
        case 8: return launch_fwd_c<8>(im, b, W, H, bg, out_color, s);
    }
    set_error("unsupported channel count %d", channels);
    return HGS_ERR_INVALID;
}

template <int C>
static int launch_bwd_c(const ImageLayout& im, const BinningLayout& b, const uint32_t* point_list, int W, int H,
                        const float* bg, const float* dL_dpix, const hgs_raster_grads* gr, float* acc16, cudaStream_t s) {
    const unsigned grid = (unsigned)(((W + HGS_TILE - 1) / HGS_TILE) * ((H + HGS_TILE - 1) / HGS_TILE));
    constexpr int CS = (C <= 4) ? 4 : 8;
    const PackedView p = packed_view(b);
    const size_t smem = 8 * sizeof(BwdWarpSmem<CS>);
    const size_t smem_half = 8 * sizeof(BwdHalfSmem<CS>);
    static std::atomic<unsigned long long> attr_done{0};  // per template instance, per device
    if (first_call_on_device(attr_done)) {
        if (int e = check_cuda(cudaFuncSetAttribute(composite_bwd_kernel<C, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    (int)smem), "composite_bwd smem attr")) return e;
        if (int e = check_cuda(cudaFuncSetAttribute(composite_bwd_half_kernel<C, CS, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    (int)smem_half), "composite_bwd_half smem attr")) return e;
        if (int e = check_cuda(cudaFuncSetAttribute(composite_bwd_half_kernel<C, CS, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    (int)smem_half), "composite_bwd_half smem attr")) return e;
        if (int e = check_cuda(cudaFuncSetAttribute(composite_bwd_half_kernel<C, CS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    (int)smem_half), "composite_bwd_half smem attr")) return e;
        if (int e = check_cuda(cudaFuncSetAttribute(composite_bwd_half_kernel<C, CS, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    (int)smem_half), "composite_bwd_half smem attr")) return e;
    }
    StageScope prof(HGS_STAGE_COMPOSITE_BWD, s);
    // half-warp kernel; kVec: interleaved [P,16] records + vector reductions (strand entry); kMask: see walk_by_mask()
#define HGS_BWD_HALF(VEC, MASK, M2D, CONIC, OPAC, COL)                                                                          \
    composite_bwd_half_kernel<C, CS, VEC, MASK><<<grid, 256, smem_half, s>>>(im.ranges, im.tile_order, point_list, W, H, bg, p.lo, \
                                                                              p.hi, p.col, p.mask, im.final_T, im.n_contrib,       \
                                                                              dL_dpix, M2D, CONIC, OPAC, COL)
    if (acc16 != nullptr && walk_by_mask())
        HGS_BWD_HALF(true, true, acc16, nullptr, nullptr, nullptr);
    else if (acc16 != nullptr)
        HGS_BWD_HALF(true, false, acc16, nullptr, nullptr, nullptr);
    else if (composite_blocks_4x4() && walk_by_mask())
        HGS_BWD_HALF(false, true, gr->dL_dmean2D, gr->dL_dconic, gr->dL_dopacity, gr->dL_dcolor);
    else if (composite_blocks_4x4())
        HGS_BWD_HALF(false, false, gr->dL_dmean2D, gr->dL_dconic, gr->dL_dopacity, gr->dL_dcolor);
#undef HGS_BWD_HALF
    else
        composite_bwd_kernel<C, CS><<<grid, 256, smem, s>>>(im.ranges, im.tile_order, point_list, W, H, bg, p.lo, p.hi, p.col,
                                                             im.final_T, im.n_contrib, dL_dpix, gr->dL_dmean2D, gr->dL_dconic,
                                                             gr->dL_dopacity, gr->dL_dcolor);
    return check_cuda(cudaGetLastError(), "composite_bwd launch");
}

int launch_composite_bwd(int channels, const ImageLayout& im, const BinningLayout& b, const uint32_t* point_list, int W,
                         int H, const float* bg, const float* dL_dpix, const hgs_raster_grads* gr, float* acc16, cudaStream_t s) {
    switch (channels) {
        case 1: return launch_bwd_c<1>(im, b, point_list, W, H, bg, dL_dpix, gr, acc16, s);
        case 2: return launch_bwd_c<2>(im, b, point_list, W, H, bg, dL_dpix, gr, acc16, s);
        case 3: return launch_bwd_c<3>(im, b, point_list, W, H, bg, dL_dpix, gr, acc16, s);
        case 4: return launch_bwd_c<4>(im, b, point_list, W, H, bg, dL_dpix, gr, acc16, s);
        case 5: return launch_bwd_c<5>(im, b, point_list, W, H, bg, dL_dpix, gr, acc16, s);
        case 6: return launch_bwd_c<6>(im, b, point_list, W, H, bg, dL_dpix, gr, acc16, s);
        case 7: return launch_bwd_c<7>(im, b, point_list, W, H, bg, dL_dpix, gr, acc16, s);
        case 8: return launch_bwd_c<8>(im, b, point_list, W, H, bg, dL_dpix, gr, acc16, s);
    }
    set_error("unsupported channel count %d", channels);
    return HGS_ERR_INVALID;
}

}  // namespace hgs
