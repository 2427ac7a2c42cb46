// losses.cu — image-space loss building block (first piece): weighted L1 over a stack of
// image planes, value AND gradient in one pass.  Hair-GS's l1_loss (loss/losses.py:16-17) is
// mean(|render - gt|); one training view evaluates it on the RGB render and, in the one-pass strand entry, on the
// mask and orientation planes too.  In torch that is ~25 small kernels over 4-28 MB each (sub, abs, mean, their
// autograd mirrors, the slice-backward zero fills); here: read image + target once, write dL/dimage once.
#include "hgs_common.cuh"

#include <cuda_fp16.h>

namespace hgs {

// loss += sum_c w[c] * sum_i |img[c][i] - tgt[c][i]| ;  dL[c][i] = w[c] * sign(img - tgt)   (sign(0) = 0 as torch)
__global__ void __launch_bounds__(256) weighted_l1_kernel(int C, long long HW, const float* __restrict__ img,
                                                          const float* __restrict__ tgt, const float* __restrict__ w,
                                                          float* __restrict__ loss, float* __restrict__ dL) {
    __shared__ float s_part[8];
    const int c = blockIdx.y;
    const float wc = w[c];
    const float4* a4 = reinterpret_cast<const float4*>(img + (size_t)c * HW);
    const float4* b4 = reinterpret_cast<const float4*>(tgt + (size_t)c * HW);
    float4* d4 = reinterpret_cast<float4*>(dL + (size_t)c * HW);
    const long long n4 = HW / 4;
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = ldg_stream4(a4 + i), b = ldg_stream4(b4 + i);
        const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
        acc += fabsf(dx) + fabsf(dy) + fabsf(dz) + fabsf(dw);
        float4 g;
        g.x = dx > 0.f ? wc : (dx < 0.f ? -wc : 0.f);
        g.y = dy > 0.f ? wc : (dy < 0.f ? -wc : 0.f);
        g.z = dz > 0.f ? wc : (dz < 0.f ? -wc : 0.f);
        g.w = dw > 0.f ? wc : (dw < 0.f ? -wc : 0.f);
        d4[i] = g;
    }
    if (blockIdx.x == 0) {  // ragged tail (HW not a multiple of 4)
        for (long long i = n4 * 4 + threadIdx.x; i < HW; i += blockDim.x) {
            const float dx = img[(size_t)c * HW + i] - tgt[(size_t)c * HW + i];
            acc += fabsf(dx);
            dL[(size_t)c * HW + i] = dx > 0.f ? wc : (dx < 0.f ? -wc : 0.f);
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 8) {
        float v = s_part[threadIdx.x];
        v += __shfl_xor_sync(0xffu, v, 4);
        v += __shfl_xor_sync(0xffu, v, 2);
        v += __shfl_xor_sync(0xffu, v, 1);
        if (threadIdx.x == 0) atomicAdd(loss, v * wc);
    }
}

int launch_weighted_l1(int C, long long HW, const float* img, const float* tgt, const float* w, float* loss, float* dL,
                       cudaStream_t s) {
    if (int e = check_cuda(cudaMemsetAsync(loss, 0, sizeof(float), s), "memset loss")) return e;
    if (C <= 0 || HW <= 0) return HGS_OK;
    const bool aligned = (HW % 4 == 0) && (((uintptr_t)img | (uintptr_t)tgt | (uintptr_t)dL) % 16 == 0);
    if (!aligned) { set_error("weighted_l1 needs 16-byte aligned planes with H*W %% 4 == 0"); return HGS_ERR_INVALID; }
    long long nb = (HW / 4 + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    if (nb < 1) nb = 1;
    StageScope prof(HGS_STAGE_OTHER, s);
    weighted_l1_kernel<<<dim3((unsigned)nb, (unsigned)C), 256, 0, s>>>(C, HW, img, tgt, w, loss, dL);
    return check_cuda(cudaGetLastError(), "weighted_l1 launch");
}


// =====================================================================================================================
// Hair-GS image-space loss of one training view, value + gradient w.r.t. the seven rendered planes
// (SURVEY.md §8f row N3):
//   L = (1-l_dssim) * l1(rgb, gt) + l_dssim * (1 - ssim(rgb, gt))                       loss/losses.py:336-339, :16-17, :43-84
//     + l_mask   * BCEWithLogits(mask_plane, gt_mask)                                   loss/losses.py:292-316
//     + l_orient * mean_{mask}( bidirectional_angle_diff(theta(orientation), gt_theta) * confidence )   :224-289
// In torch this is ~70 kernels over 1-12 MB tensors per view (five grouped 11x11 conv2d + their autograd, the
// permute/matmul/norm/atan2/where chain of the orientation term, BCE); here it is four launches (three kernels over the image + the one-thread finish):
//   ssim_fwd_kernel      : 32x32 output tile per CTA, 42x42 halo of render + target staged in shared memory, separable
//                          11-tap Gaussian for the five moments, SSIM map -> loss sum and three derivative maps
//   ssim_bwd_kernel      : convolves the three derivative maps back (same tiling) and adds the L1 term's sign(); its
//                          channel-0 blocks also count the pixels of the orientation mask (the mean's denominator is
//                          data dependent)
//   hair_pointwise_kernel: BCE-with-logits and the orientation chain, forward value and analytic gradient per pixel
// =====================================================================================================================
static constexpr int kSsimTile = 32;                        // output tile edge
static constexpr int kSsimHalo = 5;
static constexpr int kSsimIn = kSsimTile + 2 * kSsimHalo;   // 42
static constexpr int kSsimInStride = kSsimIn + 1;           // 43: odd stride, the 4-row x 8-segment warp footprint of
                                                            //     the horizontal pass maps to 32 distinct banks
static constexpr int kSsimHStride = kSsimTile + 1;          // 33
static constexpr int kSsimStrip = 4;                        // outputs per thread along the filtered axis
static constexpr int kSsimTaps = 11;
static constexpr int kSsimWin = kSsimStrip + kSsimTaps - 1; // 14 inputs feed 4 outputs

// normalised 11-tap Gaussian, sigma 1.5 (loss/losses.py:24-40): exp(-(x-5)^2 / 4.5) / sum, evaluated in float32.
// A compile-time initialiser: valid on every device of the process and inside a stream capture (no lazy upload).
__device__ __constant__ float kGauss11[11] = {0.00102838036f, 0.00759875868f, 0.0360007733f, 0.109360702f, 0.213005543f, 0.266011745f, 0.213005543f, 0.109360702f, 0.0360007733f, 0.00759875868f, 0.00102838036f};

using HairLossArgs = hgs_hair_loss;

__device__ __forceinline__ float block_sum_256(float v, float* s_part) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x < 8) {
        r = s_part[threadIdx.x];
        r += __shfl_xor_sync(0xffu, r, 4);
        r += __shfl_xor_sync(0xffu, r, 2);
        r += __shfl_xor_sync(0xffu, r, 1);
    }
    __syncthreads();
    return r;  // valid in thread 0
}

__device__ __forceinline__ bool orient_in_mask(const HairLossArgs& a, long long i, float ox, float oy, float oz) {
    if (a.orient_mask) return a.orient_mask[i] != 0;
    return ox != a.bg_orient[0] || oy != a.bg_orient[1] || oz != a.bg_orient[2];
}

// Stages the 42 x 42 halo tile of one plane (zero outside the image, F.conv2d padding=5) into shared memory: rows
// [y0, y0+42), columns [tx-5, tx+37) with tx = 32 * blockIdx.x.  When rows are 16-byte aligned (W % 4 == 0, aligned plane)
// the 48-column window [tx-8, tx+40) is read as 12 float4 per row - 504 128-bit loads per plane instead of 1764 scalar ones
// with their index arithmetic (the halo loads were the kernels' top stall, profiles/r2_loss_ncu.md).
__device__ __forceinline__ void ssim_stage_tile(float (*dst)[kSsimInStride], const float* __restrict__ src, int W, int H,
                                                int tx, int y0) {
    if (((W & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
        for (int i = threadIdx.x; i < kSsimIn * 12; i += 256) {
            const int r = i / 12, v = i - r * 12;
            const int yy = y0 + r, c0 = tx - 8 + 4 * v;
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (yy >= 0 && yy < H && c0 >= 0 && c0 < W) t = __ldg(reinterpret_cast<const float4*>(src + (long long)yy * W + c0));
            const int q = 4 * v - 3;   // tile column of t.x: the window's first three and last three columns are not the tile's
            if (v == 0) {
                dst[r][0] = t.w;
            } else if (v == 11) {
                dst[r][41] = t.x;
            } else {
                dst[r][q] = t.x;
                dst[r][q + 1] = t.y;
                dst[r][q + 2] = t.z;
                dst[r][q + 3] = t.w;
            }
        }
        return;
    }
    for (int i = threadIdx.x; i < kSsimIn * kSsimIn; i += 256) {
        const int r = i / kSsimIn, q = i - r * kSsimIn;
        const int yy = y0 + r, xx = tx - kSsimHalo + q;
        const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
        dst[r][q] = in ? __ldg(src + (long long)yy * W + xx) : 0.f;
    }
}

// Separable 11-tap filter with register sliding windows: every thread produces kSsimStrip consecutive outputs from
// kSsimWin inputs, so a tap costs one FMA and 1/4 of a shared-memory read instead of one read per FMA.
//   pass 1 (rows):    item = (input row r, 4-column segment)  -> 42 x 8 items
//   pass 2 (columns): item = (output column, 4-row strip)     -> 32 x 8 items = one per thread; a warp owns one strip row
template <int NM>
__device__ __forceinline__ void ssim_vertical(const float (*s_h)[kSsimIn][kSsimHStride], int col, int strip,
                                              float (&out)[NM][kSsimStrip]) {
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        float w[kSsimWin];
#pragma unroll
        for (int i = 0; i < kSsimWin; ++i) w[i] = s_h[m][strip * kSsimStrip + i][col];
#pragma unroll
        for (int j = 0; j < kSsimStrip; ++j) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < kSsimTaps; ++k) acc += kGauss11[k] * w[j + k];
            out[m][j] = acc;
        }
    }
}

// grid (ceil(W/32), ceil(H/32), 3 channels), 256 threads
__global__ void __launch_bounds__(256, 5) ssim_fwd_kernel(const HairLossArgs a) {
    __shared__ float s_x[kSsimIn][kSsimInStride];
    __shared__ float s_y[kSsimIn][kSsimInStride];
    __shared__ float s_h[5][kSsimIn][kSsimHStride];  // row-filtered x, y, xx, yy, xy
    __shared__ float s_part[8];
    const int H = a.height, W = a.width, c = blockIdx.z;
    const long long HW = (long long)H * W;
    const float* X = a.image7 + c * HW;
    const float* Y = a.gt_rgb + c * HW;
    const int y0 = blockIdx.y * kSsimTile - kSsimHalo;
    ssim_stage_tile(s_x, X, W, H, blockIdx.x * kSsimTile, y0);
    ssim_stage_tile(s_y, Y, W, H, blockIdx.x * kSsimTile, y0);
    __syncthreads();
    for (int i = threadIdx.x; i < kSsimIn * (kSsimTile / kSsimStrip); i += 256) {
        const int r = i >> 3, c0 = (i & 7) * kSsimStrip;
        float vx[kSsimWin], vy[kSsimWin];
#pragma unroll
        for (int j = 0; j < kSsimWin; ++j) { vx[j] = s_x[r][c0 + j]; vy[j] = s_y[r][c0 + j]; }
        float acc[5][kSsimStrip];
#pragma unroll
        for (int m = 0; m < 5; ++m)
#pragma unroll
            for (int j = 0; j < kSsimStrip; ++j) acc[m][j] = 0.f;
#pragma unroll
        for (int t = 0; t < kSsimWin; ++t) {
            const float x = vx[t], y = vy[t], xx = x * x, yy = y * y, xy = x * y;
#pragma unroll
            for (int j = 0; j < kSsimStrip; ++j) {
                const int k = t - j;
                if (k >= 0 && k < kSsimTaps) {
                    const float g = kGauss11[k];
                    acc[0][j] += g * x; acc[1][j] += g * y; acc[2][j] += g * xx; acc[3][j] += g * yy; acc[4][j] += g * xy;
                }
            }
        }
#pragma unroll
        for (int m = 0; m < 5; ++m)
#pragma unroll
            for (int j = 0; j < kSsimStrip; ++j) s_h[m][r][c0 + j] = acc[m][j];
    }
    __syncthreads();
    const int col = threadIdx.x & 31, strip = threadIdx.x >> 5;
    float mom[5][kSsimStrip];
    ssim_vertical<5>(s_h, col, strip, mom);
    const int px = blockIdx.x * kSsimTile + col;
    float ssim_sum = 0.f, l1_sum = 0.f;
#pragma unroll
    for (int j = 0; j < kSsimStrip; ++j) {
        const int ly = strip * kSsimStrip + j, py = blockIdx.y * kSsimTile + ly;
        if (px < W && py < H) {

            const float mu1 = mom[0][j], mu2 = mom[1][j], e11 = mom[2][j], e22 = mom[3][j], e12 = mom[4][j];
            const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
            const float s11 = e11 - mu1_sq, s22 = e22 - mu2_sq, s12 = e12 - mu12;
            const float A1 = 2.f * mu12 + C1, A2 = 2.f * s12 + C2, B1 = mu1_sq + mu2_sq + C1, B2 = s11 + s22 + C2;
            // B1, B2 >= C1, C2 > 0: two MUFU.RCP instead of three IEEE divisions
            const float inv_B1 = rcp_approx(B1), inv_B2 = rcp_approx(B2);
            const float inv = inv_B1 * inv_B2;
            const float ssim_val = A1 * A2 * inv;
            // d ssim / d(mu1, E[x^2], E[xy]) with s11 = E[x^2]-mu1^2, s12 = E[xy]-mu1 mu2
            const float d_A1 = A2 * inv, d_A2 = A1 * inv, d_B1 = -ssim_val * inv_B1, d_B2 = -ssim_val * inv_B2;
            const float d_mu1 = d_A1 * 2.f * mu2 + d_B1 * 2.f * mu1 + d_B2 * (-2.f * mu1) + d_A2 * 2.f * (-mu2);
            const long long p = (long long)py * W + px;
            a.scratch[(0 * 3 + c) * HW + p] = d_mu1;
            a.scratch[(1 * 3 + c) * HW + p] = d_B2;         // d/dE[x^2]
            a.scratch[(2 * 3 + c) * HW + p] = 2.f * d_A2;   // d/dE[xy]
            ssim_sum += ssim_val;
            l1_sum += fabsf(s_x[ly + kSsimHalo][col + kSsimHalo] - s_y[ly + kSsimHalo][col + kSsimHalo]);
        }
    }
    const float rs = block_sum_256(ssim_sum, s_part);
    const float rl = block_sum_256(l1_sum, s_part);
    if (threadIdx.x == 0) {
        atomicAdd(a.terms + 2, rs);
        atomicAdd(a.terms + 1, rl);
    }
}

// dL/dx = -(l_dssim / n) * [ conv(d_mu1) + 2 x conv(d_e11) + y conv(d_e12) ] + (l_l1 / n) * sign(x - y)
__global__ void __launch_bounds__(256, 5) ssim_bwd_kernel(const HairLossArgs a) {
    __shared__ float s_m[3][kSsimIn][kSsimInStride];
    __shared__ float s_h[3][kSsimIn][kSsimHStride];
    __shared__ float s_part[8];
    const int H = a.height, W = a.width, c = blockIdx.z;
    const long long HW = (long long)H * W;
    const int y0 = blockIdx.y * kSsimTile - kSsimHalo;
#pragma unroll
    for (int m = 0; m < 3; ++m) ssim_stage_tile(s_m[m], a.scratch + (m * 3 + c) * HW, W, H, blockIdx.x * kSsimTile, y0);
    __syncthreads();
    for (int i = threadIdx.x; i < kSsimIn * (kSsimTile / kSsimStrip); i += 256) {
        const int r = i >> 3, c0 = (i & 7) * kSsimStrip;
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            float w[kSsimWin];
#pragma unroll
            for (int j = 0; j < kSsimWin; ++j) w[j] = s_m[m][r][c0 + j];
#pragma unroll
            for (int j = 0; j < kSsimStrip; ++j) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < kSsimTaps; ++k) acc += kGauss11[k] * w[j + k];
                s_h[m][r][c0 + j] = acc;
            }
        }
    }
    __syncthreads();
    const int col = threadIdx.x & 31, strip = threadIdx.x >> 5;
    float cv[3][kSsimStrip];
    ssim_vertical<3>(s_h, col, strip, cv);
    const int px = blockIdx.x * kSsimTile + col;
    const float n = 3.f * (float)HW;
    float cnt = 0.f;
#pragma unroll
    for (int j = 0; j < kSsimStrip; ++j) {
        const int py = blockIdx.y * kSsimTile + strip * kSsimStrip + j;
        if (px < W && py < H) {
            const long long p = (long long)py * W + px;
            if (c == 0) {
                // the channel-0 blocks also count the pixels of the orientation mask (the denominator of the orientation
                // term's mean is data dependent; hair_pointwise, launched next, needs it before it can write gradients)
                float ox = 0.f, oy = 0.f, oz = 0.f;
                if (!a.orient_mask) { ox = a.image7[4 * HW + p]; oy = a.image7[5 * HW + p]; oz = a.image7[6 * HW + p]; }
                cnt += orient_in_mask(a, p, ox, oy, oz) ? 1.f : 0.f;
            }
            const float x = a.image7[c * HW + p], y = a.gt_rgb[c * HW + p];
            const float dssim = cv[0][j] + 2.f * x * cv[1][j] + y * cv[2][j];
            const float d = x - y;
            const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
            a.dL_dimage[c * HW + p] = -(a.l_dssim / n) * dssim + (a.l_l1 / n) * sgn;
        }
    }
    if (c == 0) {   // block-uniform
        const float rc = block_sum_256(cnt, s_part);
        if (threadIdx.x == 0 && rc != 0.f) atomicAdd(a.terms + 5, rc);
    }
}

__global__ void __launch_bounds__(256) hair_pointwise_kernel(const HairLossArgs a) {
    __shared__ float s_part[8];
    const long long HW = (long long)a.height * a.width;
    const float count = a.terms[5];
    const float inv_count = count > 0.f ? 1.f / count : 0.f;
    const float kPi = 3.14159265358979323846f, kPi2 = 1.57079632679489661923f, eps = 1e-7f;
    // world_view_transform[:3,:2]: by value, or from the camera's matrix on the device (graph replay: one graph, any view)
    float r0 = a.view_rot[0], r1 = a.view_rot[1], r3 = a.view_rot[3], r4 = a.view_rot[4], r6 = a.view_rot[6],
          r7 = a.view_rot[7];
    if (a.view_matrix_dev != nullptr) {
        const float* vm = a.view_matrix_dev;
        r0 = vm[0]; r1 = vm[1]; r3 = vm[4]; r4 = vm[5]; r6 = vm[8]; r7 = vm[9];
    }
    float bce_sum = 0.f, ori_sum = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
        // ---- mask: BCEWithLogits(x, z) = max(x,0) - x z + log(1 + exp(-|x|)), mean over pixels ----------------
        const float x = a.image7[3 * HW + i], z = a.gt_mask[i];
        const float e = expf(-fabsf(x));            // shared by the value and the sigmoid: sigmoid(x) = 1/(1+e) or e/(1+e)
        bce_sum += fmaxf(x, 0.f) - x * z + log1pf(e);
        const float inv_1pe = rcp_approx(1.f + e);
        const float sig = x >= 0.f ? inv_1pe : e * inv_1pe;
        a.dL_dimage[3 * HW + i] = a.l_mask * (sig - z) / (float)HW;
        // ---- orientation (loss/losses.py:244-288) ---------------------------------------------------------------
        const float ox = a.image7[4 * HW + i], oy = a.image7[5 * HW + i], oz = a.image7[6 * HW + i];
        // loaded before the branch so that all of the pixel's reads are in flight together (the kernel waits on memory)
        const float gt_theta = a.gt_theta[i], cf = a.confidence[i];
        float gox = 0.f, goy = 0.f, goz = 0.f;
        if (orient_in_mask(a, i, ox, oy, oz)) {
            // view-space xy: o_world @ wvt[:3,:3]
            const float vx = ox * r0 + oy * r3 + oz * r6;
            const float vy = ox * r1 + oy * r4 + oz * r7;
            const float nrm = sqrtf(vx * vx + vy * vy);
            const float den = nrm + eps;
            const float inv_den = rcp_approx(den);      // den >= 1e-7
            const float pxn = vx * inv_den;
            float pyn = vy * inv_den;
            if (pyn < eps) pyn += eps;
            float theta = atan2f(pxn, pyn);
            if (theta < 0.f) theta += kPi;
            const float dt = theta - gt_theta;
            const float inner = fabsf(dt) - kPi2;
            const float diff = kPi2 - fabsf(inner);
            ori_sum += diff * cf;
            // backward
            const float sgn_dt = dt > 0.f ? 1.f : (dt < 0.f ? -1.f : 0.f);
            const float sgn_in = inner > 0.f ? 1.f : (inner < 0.f ? -1.f : 0.f);
            const float g_theta = a.l_orient * inv_count * cf * (-sgn_in * sgn_dt);
            const float r2 = pxn * pxn + pyn * pyn;
            const float inv_r2 = r2 > 1e-30f ? rcp_approx(r2) : (r2 > 0.f ? 1.f / r2 : 0.f);
            const float g_px = g_theta * (pyn * inv_r2);    // d atan2(x, y)/dx =  y/(x^2+y^2)
            const float g_py = g_theta * (-pxn * inv_r2);   //              /dy = -x/(x^2+y^2)
            // p = v / (|v| + eps); d|v|/dv = v/|v| (0 at the origin, as torch.norm)
            const float gdotv = g_px * vx + g_py * vy;
            const float k = nrm > 0.f ? gdotv / (den * den * nrm) : 0.f;   // IEEE: the divisor can be denormal
            const float g_vx = g_px * inv_den - k * vx, g_vy = g_py * inv_den - k * vy;
            gox = g_vx * r0 + g_vy * r1;
            goy = g_vx * r3 + g_vy * r4;
            goz = g_vx * r6 + g_vy * r7;
        }
        a.dL_dimage[4 * HW + i] = gox;
        a.dL_dimage[5 * HW + i] = goy;
        a.dL_dimage[6 * HW + i] = goz;
    }
    const float rb = block_sum_256(bce_sum, s_part);
    const float ro = block_sum_256(ori_sum, s_part);
    if (threadIdx.x == 0) {
        atomicAdd(a.terms + 3, rb);
        atomicAdd(a.terms + 4, ro);
    }
}

// terms: sums -> means, total
__global__ void hair_loss_finish_kernel(const HairLossArgs a) {
    const float HW = (float)a.height * (float)a.width;
    const float l1 = a.terms[1] / (3.f * HW);
    const float dssim = 1.f - a.terms[2] / (3.f * HW);
    const float mask = a.terms[3] / HW;
    const float cnt = a.terms[5];
    const float ori = cnt > 0.f ? a.terms[4] / cnt : 0.f;
    a.terms[1] = l1; a.terms[2] = dssim; a.terms[3] = mask; a.terms[4] = ori;
    a.terms[0] = a.l_l1 * l1 + a.l_dssim * dssim + a.l_mask * mask + a.l_orient * ori;
}

// ------------------------------------------------------------------------------------------------
// target stacks: storage format (8 B / pixel) -> float planes (24 B / pixel)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unpack_targets_kernel(long long HW, long long total, const uchar4* __restrict__ rgbm,
                                                             const __half2* __restrict__ tc, float* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long v = i / HW, px = i - v * HW;
        const uchar4 c = rgbm[i];
        const float2 t = __half22float2(tc[i]);
        float* o = out + v * 6 * HW + px;
        o[0] = (float)c.x / 255.0f;
        o[HW] = (float)c.y / 255.0f;
        o[2 * HW] = (float)c.z / 255.0f;
        o[3 * HW] = (float)c.w / 255.0f;
        o[4 * HW] = t.x;
        o[5 * HW] = t.y;
    }
}

int launch_unpack_targets(int views, long long HW, const void* rgbm, const void* tc, float* out, cudaStream_t s) {
    const long long total = (long long)views * HW;
    if (total <= 0) return HGS_OK;
    long long nb = (total + 255) / 256;
    if (nb > 148 * 16) nb = 148 * 16;
    StageScope prof(HGS_STAGE_OTHER, s);
    unpack_targets_kernel<<<(unsigned)nb, 256, 0, s>>>(HW, total, (const uchar4*)rgbm, (const __half2*)tc, out);
    return check_cuda(cudaGetLastError(), "unpack_targets launch");
}

int launch_hair_image_loss(const HairLossArgs& a, cudaStream_t s) {
    if (int e = check_cuda(cudaMemsetAsync(a.terms, 0, 8 * sizeof(float), s), "memset loss terms")) return e;
    const long long HW = (long long)a.height * a.width;
    long long nb = (HW + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    StageScope prof(HGS_STAGE_OTHER, s);
    dim3 grid((a.width + kSsimTile - 1) / kSsimTile, (a.height + kSsimTile - 1) / kSsimTile, 3);
    ssim_fwd_kernel<<<grid, 256, 0, s>>>(a);
    ssim_bwd_kernel<<<grid, 256, 0, s>>>(a);
    hair_pointwise_kernel<<<(unsigned)nb, 256, 0, s>>>(a);
    hair_loss_finish_kernel<<<1, 1, 0, s>>>(a);
    return check_cuda(cudaGetLastError(), "hair_image_loss launch");
}

}  // namespace hgs
