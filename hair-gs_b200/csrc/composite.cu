// composite.cu — per-tile alpha compositing, forward and backward.
//   composite_fwd<C> replaces renderCUDA            forward.cu:261-374
//   composite_bwd<C> replaces renderCUDABW_*        backward_distwar.cu:450-1014 (all three variants)
// Blending arithmetic follows SURVEY.md App. A exactly (same expressions, IEEE expf/div, no fast-math)
// so that pixels, final_T and n_contrib agree with the reference build.
//
// B200-first structure (what differs from the reference):
//  * one CTA per 16x16 tile, but each WARP owns a compact 8x4 pixel block instead of a 16x2 strip;
//  * 256-instance batches are staged into shared memory with cp.async (LDGSTS), double buffered,
//    one 32-byte record + one 16/32-byte colour per instance — colours are staged too (the reference
//    gathers them from global memory per contributing pixel, forward.cu:355);
//  * per instance an 8-bit mask says which warps' pixel blocks the splat's alpha>=1/255 extent can
//    reach; a warp ballots the masks and visits only its own candidates.  Skipped instances are
//    exactly those for which every pixel of the warp would take the reference's `continue`
//    (power>0 or alpha<1/255), so outputs are unchanged; list positions still advance.
//  * backward: gradients of a (warp, instance) pair are reduced with a transposing shuffle tree
//    (V + 5 shuffles instead of 5 V), then 6+C lanes issue one red.global.add each.
#include "hgs_common.cuh"

namespace hgs {

static constexpr int kBatch = 256;

template <int CS>
struct StageBuf {
    float4 lo[kBatch];
    float4 hi[kBatch];
    float4 col[kBatch * (CS / 4)];
    uint32_t mask[kBatch];
};

// which of the 8 warp pixel-blocks (2 columns x 4 rows of 8x4 pixels) can the splat reach?
__device__ __forceinline__ uint32_t warp_block_mask(const float4 lo, const float4 hi, float X0, float Y0) {
    const float gx0 = lo.x - hi.z, gx1 = lo.x + hi.z;
    const float gy0 = lo.y - hi.w, gy1 = lo.y + hi.w;
    uint32_t xm = 0, ym = 0;
    // culled only when a comparison is TRUE, so NaNs keep the instance (reference would evaluate it)
    if (!(gx0 > X0 + 7.f || gx1 < X0)) xm |= 1u;
    if (!(gx0 > X0 + 15.f || gx1 < X0 + 8.f)) xm |= 2u;
#pragma unroll
    for (int r = 0; r < 4; ++r)
        if (!(gy0 > Y0 + 4.f * r + 3.f || gy1 < Y0 + 4.f * r)) ym |= 1u << r;
    uint32_t m = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w)
        if (((xm >> (w & 1)) & 1u) && ((ym >> (w >> 1)) & 1u)) m |= 1u << w;
    return m;
}

template <int CS>
__device__ __forceinline__ void stage_issue(StageBuf<CS>& sb, int tid, int id, const float4* __restrict__ rec,
                                            const float* __restrict__ rgb) {
    if (id >= 0) {
        cp_async16(&sb.lo[tid], &rec[2 * (size_t)id]);
        cp_async16(&sb.hi[tid], &rec[2 * (size_t)id + 1]);
        cp_async16(&sb.col[tid * (CS / 4)], rgb + (size_t)id * CS);
        if (CS > 4) cp_async16(&sb.col[tid * (CS / 4) + 1], rgb + (size_t)id * CS + 4);
    }
}

template <int C, int CS>
__global__ void __launch_bounds__(256) composite_fwd_kernel(const uint2* __restrict__ ranges,
                                                            const uint32_t* __restrict__ point_list, int W, int H,
                                                            const float4* __restrict__ rec,
                                                            const float* __restrict__ rgb,
                                                            const float* __restrict__ bg_color,
                                                            float* __restrict__ final_T,
                                                            uint32_t* __restrict__ n_contrib,
                                                            float* __restrict__ out_color) {
    __shared__ StageBuf<CS> sbuf[2];

    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const uint32_t horizontal_blocks = (W + HGS_TILE - 1) / HGS_TILE;
    const uint32_t X0 = blockIdx.x * HGS_TILE, Y0 = blockIdx.y * HGS_TILE;
    const uint32_t px = X0 + (warp & 1) * 8 + (lane & 7);
    const uint32_t py = Y0 + (warp >> 1) * 4 + (lane >> 3);
    const uint32_t pix_id = W * py + px;
    const float2 pixf = make_float2((float)px, (float)py);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    bool done = !inside;

    const uint2 range = ranges[blockIdx.y * horizontal_blocks + blockIdx.x];
    const int total = (int)(range.y - range.x);
    const int rounds = (total + kBatch - 1) / kBatch;

    float T = 1.0f;
    uint32_t last_contributor = 0;
    float Cacc[C];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) Cacc[ch] = 0.f;

    // software pipeline: ids two rounds ahead (registers), records one round ahead (cp.async)
    int id_next = -1;
    if (rounds > 0) {
        const int id0 = (tid < total) ? (int)point_list[range.x + tid] : -1;
        stage_issue<CS>(sbuf[0], tid, id0, rec, rgb);
        cp_async_commit();
        if (rounds > 1) id_next = (kBatch + tid < total) ? (int)point_list[range.x + kBatch + tid] : -1;
    }
    for (int i = 0; i < rounds; ++i) {
        StageBuf<CS>& sb = sbuf[i & 1];
        if (i + 1 < rounds) stage_issue<CS>(sbuf[(i + 1) & 1], tid, id_next, rec, rgb);
        cp_async_commit();
        if (i + 2 < rounds) {
            const int p = (i + 2) * kBatch + tid;
            id_next = (p < total) ? (int)point_list[range.x + p] : -1;
        }
        cp_async_wait<1>();
        const int in_round = min(kBatch, total - i * kBatch);
        sb.mask[tid] = (tid < in_round) ? warp_block_mask(sb.lo[tid], sb.hi[tid], (float)X0, (float)Y0) : 0u;
        __syncthreads();

        if (!__all_sync(0xffffffffu, done)) {
            const uint32_t pos_base = (uint32_t)(i * kBatch) + 1u;
            for (int c = 0; c * 32 < in_round; ++c) {
                uint32_t bits = __ballot_sync(0xffffffffu, (sb.mask[c * 32 + lane] >> warp) & 1u);
                while (bits) {
                    const int j = c * 32 + __ffs(bits) - 1;
                    bits &= bits - 1;
                    if (done) continue;
                    const float4 lo = sb.lo[j];
                    const float4 hi = sb.hi[j];
                    const float2 d = make_float2(lo.x - pixf.x, lo.y - pixf.y);
                    const float power = -0.5f * (lo.z * d.x * d.x + hi.x * d.y * d.y) - lo.w * d.x * d.y;
                    if (power > 0.0f) continue;
                    const float alpha = min(0.99f, hi.y * exp(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = T * (1 - alpha);
                    if (test_T < 0.0001f) {
                        done = true;
                        continue;
                    }
                    const float* col = reinterpret_cast<const float*>(&sb.col[j * (CS / 4)]);
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) Cacc[ch] += col[ch] * alpha * T;
                    T = test_T;
                    last_contributor = pos_base + (uint32_t)j;
                }
                if (__all_sync(0xffffffffu, done)) break;
            }
        }
        if (__syncthreads_count(done) == kBatch) break;
    }
    cp_async_wait<0>();

    if (inside) {
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) out_color[(size_t)ch * H * W + pix_id] = Cacc[ch] + T * bg_color[ch];
    }
}

// Sum V per-lane values across the warp with a transposing tree: after the call, lane l holds the
// warp total of value index (l >> 2) & (VP-1) where VP = next pow2 >= V (VP <= 8 -> index l>>2).
template <int VP>
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[VP], uint32_t lane) {
    static_assert(VP == 8 || VP == 16, "VP");
    // level 0 (xor 16): keep lower half of the indices if bit4 clear, upper half otherwise
    if (VP == 16) {
        const bool up = lane & 16;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float send = up ? v[k] : v[k + 8];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
            v[k] = (up ? v[k + 8] : v[k]) + recv;
        }
        // now 8 live values: indices (bit4 ? 8 : 0) + k ; continue with xor 8, 4, 2 on 8 -> 1
        {
            const bool u2 = lane & 8;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float send = u2 ? v[k] : v[k + 4];
                const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
                v[k] = (u2 ? v[k + 4] : v[k]) + recv;
            }
        }
        {
            const bool u3 = lane & 4;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float send = u3 ? v[k] : v[k + 2];
                const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
                v[k] = (u3 ? v[k + 2] : v[k]) + recv;
            }
        }
        {
            const bool u4 = lane & 2;
            const float send = u4 ? v[0] : v[1];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
            v[0] = (u4 ? v[1] : v[0]) + recv;
        }
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
        return v[0];  // lane l holds index: bit4*8 + bit3*4 + bit2*2 + bit1  == (l >> 1) & 15
    } else {
        const bool up = lane & 16;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float send = up ? v[k] : v[k + 4];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
            v[k] = (up ? v[k + 4] : v[k]) + recv;
        }
        {
            const bool u2 = lane & 8;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float send = u2 ? v[k] : v[k + 2];
                const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
                v[k] = (u2 ? v[k + 2] : v[k]) + recv;
            }
        }
        {
            const bool u3 = lane & 4;
            const float send = u3 ? v[0] : v[1];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
            v[0] = (u3 ? v[1] : v[0]) + recv;
        }
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
        return v[0];  // lane l holds index (l >> 2) & 7
    }
}

template <int C, int CS>
__global__ void __launch_bounds__(256) composite_bwd_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
    const float* __restrict__ bg_color, const float4* __restrict__ rec, const float* __restrict__ rgb,
    const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpixels,
    float* __restrict__ dL_dmean2D /*[P,3]*/, float* __restrict__ dL_dconic /*[P,4]*/,
    float* __restrict__ dL_dopacity /*[P]*/, float* __restrict__ dL_dcolors /*[P,C]*/) {
    __shared__ StageBuf<CS> sbuf[2];
    __shared__ int s_ids[2][kBatch];

    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const uint32_t horizontal_blocks = (W + HGS_TILE - 1) / HGS_TILE;
    const uint32_t X0 = blockIdx.x * HGS_TILE, Y0 = blockIdx.y * HGS_TILE;
    const uint32_t px = X0 + (warp & 1) * 8 + (lane & 7);
    const uint32_t py = Y0 + (warp >> 1) * 4 + (lane >> 3);
    const uint32_t pix_id = W * py + px;
    const float2 pixf = make_float2((float)px, (float)py);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;

    const uint2 range = ranges[blockIdx.y * horizontal_blocks + blockIdx.x];
    const int total = (int)(range.y - range.x);
    const int rounds = (total + kBatch - 1) / kBatch;

    const float T_final = inside ? final_Ts[pix_id] : 0;
    float T = T_final;
    const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;

    float accum_rec[C], dL_dpixel[C], last_color[C];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
        accum_rec[ch] = 0.f;
        last_color[ch] = 0.f;
        dL_dpixel[ch] = inside ? dL_dpixels[(size_t)ch * H * W + pix_id] : 0.f;
    }
    float last_alpha = 0.f;
    float bg_dot_dpixel = 0;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) bg_dot_dpixel += bg_color[ch] * dL_dpixel[ch];
    const float ddelx_dx = 0.5 * W;
    const float ddely_dy = 0.5 * H;

    // the warp can skip every instance at list position >= max over its lanes of last_contributor
    int warp_last = last_contributor;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, o));

    // back-to-front: slot t of round i is list position total-1-(i*256+t)
    int id_next = -1;
    if (rounds > 0) {
        const int id0 = (tid < total) ? (int)point_list[range.y - 1 - tid] : -1;
        stage_issue<CS>(sbuf[0], tid, id0, rec, rgb);
        cp_async_commit();
        s_ids[0][tid] = id0;
        if (rounds > 1) id_next = (kBatch + tid < total) ? (int)point_list[range.y - 1 - kBatch - tid] : -1;
    }
    for (int i = 0; i < rounds; ++i) {
        StageBuf<CS>& sb = sbuf[i & 1];
        if (i + 1 < rounds) {
            stage_issue<CS>(sbuf[(i + 1) & 1], tid, id_next, rec, rgb);
            s_ids[(i + 1) & 1][tid] = id_next;
        }
        cp_async_commit();
        if (i + 2 < rounds) {
            const int p = (i + 2) * kBatch + tid;
            id_next = (p < total) ? (int)point_list[range.y - 1 - p] : -1;
        }
        cp_async_wait<1>();
        const int in_round = min(kBatch, total - i * kBatch);
        sb.mask[tid] = (tid < in_round) ? warp_block_mask(sb.lo[tid], sb.hi[tid], (float)X0, (float)Y0) : 0u;
        __syncthreads();

        const int* ids = s_ids[i & 1];
        // list position (0-based, forward order) of slot j: total-1-(i*256+j)
        const int pos_first = total - 1 - i * kBatch;
        for (int c = 0; c * 32 < in_round; ++c) {
            if (pos_first - c * 32 - 31 >= warp_last) continue;  // whole chunk behind every pixel's last contributor
            uint32_t bits = __ballot_sync(0xffffffffu, (sb.mask[c * 32 + lane] >> warp) & 1u);
            while (bits) {
                const int j = c * 32 + __ffs(bits) - 1;
                bits &= bits - 1;
                const int pos = pos_first - j;
                bool valid = pos < last_contributor;  // false for outside pixels (last_contributor == 0)
                const float4 lo = sb.lo[j];
                const float4 hi = sb.hi[j];
                const float2 d = make_float2(lo.x - pixf.x, lo.y - pixf.y);
                const float power = -0.5f * (lo.z * d.x * d.x + hi.x * d.y * d.y) - lo.w * d.x * d.y;
                if (power > 0.0f) valid = false;
                const float G = exp(power);
                const float alpha = min(0.99f, hi.y * G);
                if (alpha < 1.0f / 255.0f) valid = false;
                if (!__any_sync(0xffffffffu, valid)) continue;

                constexpr int V = 6 + C;
                constexpr int VP = (V <= 8) ? 8 : 16;
                float v[VP];
#pragma unroll
                for (int k = 0; k < VP; ++k) v[k] = 0.f;
                if (valid) {
                    T = T / (1.f - alpha);
                    const float dchannel_dcolor = alpha * T;
                    float dL_dalpha = 0.0f;
                    const float* col = reinterpret_cast<const float*>(&sb.col[j * (CS / 4)]);
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) {
                        const float cc = col[ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                        last_color[ch] = cc;
                        const float dL_dchannel = dL_dpixel[ch];
                        dL_dalpha += (cc - accum_rec[ch]) * dL_dchannel;
                        v[6 + ch] = dchannel_dcolor * dL_dchannel;
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;
                    const float dL_dG = hi.y * dL_dalpha;
                    const float gdx = G * d.x;
                    const float gdy = G * d.y;
                    const float dG_ddelx = -gdx * lo.z - gdy * lo.w;
                    const float dG_ddely = -gdy * hi.x - gdx * lo.w;
                    v[0] = dL_dG * dG_ddelx * ddelx_dx;
                    v[1] = dL_dG * dG_ddely * ddely_dy;
                    v[2] = -0.5f * gdx * d.x * dL_dG;
                    v[3] = -0.5f * gdx * d.y * dL_dG;
                    v[4] = -0.5f * gdy * d.y * dL_dG;
                    v[5] = G * dL_dalpha;
                }
                const float red = warp_transpose_reduce<VP>(v, lane);
                const int gid = ids[j];
                if (VP == 8) {
                    const int k = (int)(lane >> 2);
                    if ((lane & 3) == 0) {
                        float* dst;
                        switch (k) {
                            case 0: dst = dL_dmean2D + 3 * (size_t)gid; break;
                            case 1: dst = dL_dmean2D + 3 * (size_t)gid + 1; break;
                            case 2: dst = dL_dconic + 4 * (size_t)gid; break;
                            case 3: dst = dL_dconic + 4 * (size_t)gid + 1; break;
                            case 4: dst = dL_dconic + 4 * (size_t)gid + 3; break;
                            case 5: dst = dL_dopacity + gid; break;
                            default: dst = dL_dcolors + (size_t)gid * C + (k - 6); break;
                        }
                        if (k < V) atomicAdd(dst, red);
                    }
                } else {
                    const int k = (int)(lane >> 1);
                    if ((lane & 1) == 0 && k < V) {
                        float* dst;
                        switch (k) {
                            case 0: dst = dL_dmean2D + 3 * (size_t)gid; break;
                            case 1: dst = dL_dmean2D + 3 * (size_t)gid + 1; break;
                            case 2: dst = dL_dconic + 4 * (size_t)gid; break;
                            case 3: dst = dL_dconic + 4 * (size_t)gid + 1; break;
                            case 4: dst = dL_dconic + 4 * (size_t)gid + 3; break;
                            case 5: dst = dL_dopacity + gid; break;
                            default: dst = dL_dcolors + (size_t)gid * C + (k - 6); break;
                        }
                        atomicAdd(dst, red);
                    }
                }
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
template <int C>
static int launch_fwd_c(const ImageLayout& im, const uint32_t* point_list, int W, int H, const GeomLayout& g,
                        const float* bg, float* out_color, cudaStream_t s) {
    dim3 grid((W + HGS_TILE - 1) / HGS_TILE, (H + HGS_TILE - 1) / HGS_TILE);
    constexpr int CS = (C <= 4) ? 4 : 8;
    StageScope prof(HGS_STAGE_COMPOSITE_FWD, s);
    composite_fwd_kernel<C, CS><<<grid, 256, 0, s>>>(im.ranges, point_list, W, H, g.rec, g.rgb, bg, im.final_T,
                                                      im.n_contrib, out_color);
    return check_cuda(cudaGetLastError(), "composite_fwd launch");
}

int launch_composite_fwd(int channels, const ImageLayout& im, const uint32_t* point_list, int W, int H,
                         const GeomLayout& g, const float* bg, float* out_color, cudaStream_t s) {
    switch (channels) {
        case 1: return launch_fwd_c<1>(im, point_list, W, H, g, bg, out_color, s);
        case 2: return launch_fwd_c<2>(im, point_list, W, H, g, bg, out_color, s);
        case 3: return launch_fwd_c<3>(im, point_list, W, H, g, bg, out_color, s);
        case 4: return launch_fwd_c<4>(im, point_list, W, H, g, bg, out_color, s);
        case 5: return launch_fwd_c<5>(im, point_list, W, H, g, bg, out_color, s);
        case 6: return launch_fwd_c<6>(im, point_list, W, H, g, bg, out_color, s);
        case 7: return launch_fwd_c<7>(im, point_list, W, H, g, bg, out_color, s);
        case 8: return launch_fwd_c<8>(im, point_list, W, H, g, bg, out_color, s);
    }
    set_error("unsupported channel count %d", channels);
    return HGS_ERR_INVALID;
}

template <int C>
static int launch_bwd_c(const ImageLayout& im, const uint32_t* point_list, int W, int H, const GeomLayout& g,
                        const float* bg, const float* dL_dpix, const hgs_raster_grads* gr, cudaStream_t s) {
    dim3 grid((W + HGS_TILE - 1) / HGS_TILE, (H + HGS_TILE - 1) / HGS_TILE);
    constexpr int CS = (C <= 4) ? 4 : 8;
    StageScope prof(HGS_STAGE_COMPOSITE_BWD, s);
    composite_bwd_kernel<C, CS><<<grid, 256, 0, s>>>(im.ranges, point_list, W, H, bg, g.rec, g.rgb, im.final_T,
                                                      im.n_contrib, dL_dpix, gr->dL_dmean2D, gr->dL_dconic,
                                                      gr->dL_dopacity, gr->dL_dcolor);
    return check_cuda(cudaGetLastError(), "composite_bwd launch");
}

int launch_composite_bwd(int channels, const ImageLayout& im, const uint32_t* point_list, int W, int H,
                         const GeomLayout& g, const float* bg, const float* dL_dpix, const hgs_raster_grads* gr,
                         cudaStream_t s) {
    switch (channels) {
        case 1: return launch_bwd_c<1>(im, point_list, W, H, g, bg, dL_dpix, gr, s);
        case 2: return launch_bwd_c<2>(im, point_list, W, H, g, bg, dL_dpix, gr, s);
        case 3: return launch_bwd_c<3>(im, point_list, W, H, g, bg, dL_dpix, gr, s);
        case 4: return launch_bwd_c<4>(im, point_list, W, H, g, bg, dL_dpix, gr, s);
        case 5: return launch_bwd_c<5>(im, point_list, W, H, g, bg, dL_dpix, gr, s);
        case 6: return launch_bwd_c<6>(im, point_list, W, H, g, bg, dL_dpix, gr, s);
        case 7: return launch_bwd_c<7>(im, point_list, W, H, g, bg, dL_dpix, gr, s);
        case 8: return launch_bwd_c<8>(im, point_list, W, H, g, bg, dL_dpix, gr, s);
    }
    set_error("unsupported channel count %d", channels);
    return HGS_ERR_INVALID;
}

}  // namespace hgs
