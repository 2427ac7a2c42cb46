// losses.cu — image-space loss building block (first piece): weighted L1 over a stack of
// image planes, value AND gradient in one pass.  Hair-GS's l1_loss (loss/losses.py:16-17) is
// mean(|render - gt|); one training view evaluates it on the RGB render and, in the one-pass strand entry, on the
// mask and orientation planes too.  In torch that is ~25 small kernels over 4-28 MB each (sub, abs, mean, their
// autograd mirrors, the slice-backward zero fills); here: read image + target once, write dL/dimage once.
#include "hgs_common.cuh"

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
// permute/matmul/norm/atan2/where chain of the orientation term, BCE); here it is four launches:
//   hair_loss_count      : number of pixels in the orientation mask (the mean's denominator is data dependent)
//   ssim_fwd_kernel      : 16x16 output tile per CTA, 26x26 halo of render + target staged in shared memory, separable
//                          11-tap Gaussian for the five moments, SSIM map -> loss sum and three derivative maps
//   ssim_bwd_kernel      : convolves the three derivative maps back (same tiling) and adds the L1 term's sign()
//   hair_pointwise_kernel: BCE-with-logits and the orientation chain, forward value and analytic gradient per pixel
// =====================================================================================================================
static constexpr int kSsimTile = 16;
static constexpr int kSsimHalo = 5;
static constexpr int kSsimIn = kSsimTile + 2 * kSsimHalo;  // 26

__device__ __constant__ float kGauss11[11];  // normalised 11-tap Gaussian, sigma 1.5 (loss/losses.py:24-40)

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

__global__ void __launch_bounds__(256) hair_loss_count_kernel(const HairLossArgs a) {
    __shared__ float s_part[8];
    const long long HW = (long long)a.height * a.width;
    float cnt = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
        const float ox = a.image7[4 * HW + i], oy = a.image7[5 * HW + i], oz = a.image7[6 * HW + i];
        cnt += orient_in_mask(a, i, ox, oy, oz) ? 1.f : 0.f;
    }
    const float r = block_sum_256(cnt, s_part);
    if (threadIdx.x == 0 && r != 0.f) atomicAdd(a.terms + 5, r);
}

// grid (ceil(W/16), ceil(H/16), 3 channels), 256 threads
__global__ void __launch_bounds__(256) ssim_fwd_kernel(const HairLossArgs a) {
    __shared__ float s_x[kSsimIn][kSsimIn + 1];
    __shared__ float s_y[kSsimIn][kSsimIn + 1];
    __shared__ float s_h[5][kSsimIn][kSsimTile + 1];  // horizontally filtered x, y, xx, yy, xy
    __shared__ float s_part[8];
    const int H = a.height, W = a.width, c = blockIdx.z;
    const long long HW = (long long)H * W;
    const float* X = a.image7 + c * HW;
    const float* Y = a.gt_rgb + c * HW;
    const int x0 = blockIdx.x * kSsimTile - kSsimHalo, y0 = blockIdx.y * kSsimTile - kSsimHalo;
    for (int i = threadIdx.x; i < kSsimIn * kSsimIn; i += 256) {
        const int r = i / kSsimIn, q = i % kSsimIn;
        const int yy = y0 + r, xx = x0 + q;
        const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;  // zero padding (F.conv2d padding=5)
        s_x[r][q] = in ? X[(long long)yy * W + xx] : 0.f;
        s_y[r][q] = in ? Y[(long long)yy * W + xx] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSsimIn * kSsimTile; i += 256) {
        const int r = i / kSsimTile, q = i % kSsimTile;
        float mx = 0, my = 0, mxx = 0, myy = 0, mxy = 0;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float g = kGauss11[k], vx = s_x[r][q + k], vy = s_y[r][q + k];
            mx += g * vx; my += g * vy; mxx += g * vx * vx; myy += g * vy * vy; mxy += g * vx * vy;
        }
        s_h[0][r][q] = mx; s_h[1][r][q] = my; s_h[2][r][q] = mxx; s_h[3][r][q] = myy; s_h[4][r][q] = mxy;
    }
    __syncthreads();
    const int tx = threadIdx.x % kSsimTile, ty = threadIdx.x / kSsimTile;
    const int px = blockIdx.x * kSsimTile + tx, py = blockIdx.y * kSsimTile + ty;
    float ssim_val = 0.f, l1_val = 0.f;
    if (px < W && py < H) {
        float mu1 = 0, mu2 = 0, e11 = 0, e22 = 0, e12 = 0;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float g = kGauss11[k];
            mu1 += g * s_h[0][ty + k][tx]; mu2 += g * s_h[1][ty + k][tx]; e11 += g * s_h[2][ty + k][tx];
            e22 += g * s_h[3][ty + k][tx]; e12 += g * s_h[4][ty + k][tx];
        }
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float s11 = e11 - mu1_sq, s22 = e22 - mu2_sq, s12 = e12 - mu12;
        const float A1 = 2.f * mu12 + C1, A2 = 2.f * s12 + C2, B1 = mu1_sq + mu2_sq + C1, B2 = s11 + s22 + C2;
        const float inv = 1.f / (B1 * B2);
        ssim_val = A1 * A2 * inv;
        // d ssim / d(mu1, E[x^2], E[xy]) with s11 = E[x^2]-mu1^2, s12 = E[xy]-mu1 mu2
        const float d_A1 = A2 * inv, d_A2 = A1 * inv, d_B1 = -ssim_val / B1, d_B2 = -ssim_val / B2;
        const float d_e11 = d_B2;              // via s11
        const float d_e12 = 2.f * d_A2;        // via s12
        const float d_mu1 = d_A1 * 2.f * mu2 + d_B1 * 2.f * mu1 + d_B2 * (-2.f * mu1) + d_A2 * 2.f * (-mu2);
        const long long p = (long long)py * W + px;
        a.scratch[(0 * 3 + c) * HW + p] = d_mu1;
        a.scratch[(1 * 3 + c) * HW + p] = d_e11;
        a.scratch[(2 * 3 + c) * HW + p] = d_e12;
        l1_val = fabsf(s_x[ty + kSsimHalo][tx + kSsimHalo] - s_y[ty + kSsimHalo][tx + kSsimHalo]);
    }
    const float rs = block_sum_256(ssim_val, s_part);
    const float rl = block_sum_256(l1_val, s_part);
    if (threadIdx.x == 0) {
        atomicAdd(a.terms + 2, rs);
        atomicAdd(a.terms + 1, rl);
    }
}

// dL/dx = -(l_dssim / n) * [ conv(d_mu1) + 2 x conv(d_e11) + y conv(d_e12) ] + (l_l1 / n) * sign(x - y)
__global__ void __launch_bounds__(256) ssim_bwd_kernel(const HairLossArgs a) {
    __shared__ float s_m[3][kSsimIn][kSsimIn + 1];
    __shared__ float s_h[3][kSsimIn][kSsimTile + 1];
    const int H = a.height, W = a.width, c = blockIdx.z;
    const long long HW = (long long)H * W;
    const int x0 = blockIdx.x * kSsimTile - kSsimHalo, y0 = blockIdx.y * kSsimTile - kSsimHalo;
    for (int i = threadIdx.x; i < kSsimIn * kSsimIn; i += 256) {
        const int r = i / kSsimIn, q = i % kSsimIn;
        const int yy = y0 + r, xx = x0 + q;
        const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
        const long long p = (long long)yy * W + xx;
#pragma unroll
        for (int m = 0; m < 3; ++m) s_m[m][r][q] = in ? a.scratch[(m * 3 + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSsimIn * kSsimTile; i += 256) {
        const int r = i / kSsimTile, q = i % kSsimTile;
        float v0 = 0, v1 = 0, v2 = 0;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float g = kGauss11[k];
            v0 += g * s_m[0][r][q + k]; v1 += g * s_m[1][r][q + k]; v2 += g * s_m[2][r][q + k];
        }
        s_h[0][r][q] = v0; s_h[1][r][q] = v1; s_h[2][r][q] = v2;
    }
    __syncthreads();
    const int tx = threadIdx.x % kSsimTile, ty = threadIdx.x / kSsimTile;
    const int px = blockIdx.x * kSsimTile + tx, py = blockIdx.y * kSsimTile + ty;
    if (px < W && py < H) {
        float c0 = 0, c1 = 0, c2 = 0;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float g = kGauss11[k];
            c0 += g * s_h[0][ty + k][tx]; c1 += g * s_h[1][ty + k][tx]; c2 += g * s_h[2][ty + k][tx];
        }
        const long long p = (long long)py * W + px;
        const float x = a.image7[c * HW + p], y = a.gt_rgb[c * HW + p];
        const float n = 3.f * (float)HW;
        const float dssim = c0 + 2.f * x * c1 + y * c2;
        const float d = x - y;
        const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        a.dL_dimage[c * HW + p] = -(a.l_dssim / n) * dssim + (a.l_l1 / n) * sgn;
    }
}

__global__ void __launch_bounds__(256) hair_pointwise_kernel(const HairLossArgs a) {
    __shared__ float s_part[8];
    const long long HW = (long long)a.height * a.width;
    const float count = a.terms[5];
    const float inv_count = count > 0.f ? 1.f / count : 0.f;
    const float kPi = 3.14159265358979323846f, kPi2 = 1.57079632679489661923f, eps = 1e-7f;
    float bce_sum = 0.f, ori_sum = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
        // ---- mask: BCEWithLogits(x, z) = max(x,0) - x z + log(1 + exp(-|x|)), mean over pixels ----------------
        const float x = a.image7[3 * HW + i], z = a.gt_mask[i];
        bce_sum += fmaxf(x, 0.f) - x * z + log1pf(expf(-fabsf(x)));
        const float sig = 1.f / (1.f + expf(-x));
        a.dL_dimage[3 * HW + i] = a.l_mask * (sig - z) / (float)HW;
        // ---- orientation (loss/losses.py:244-288) ---------------------------------------------------------------
        const float ox = a.image7[4 * HW + i], oy = a.image7[5 * HW + i], oz = a.image7[6 * HW + i];
        float gox = 0.f, goy = 0.f, goz = 0.f;
        if (orient_in_mask(a, i, ox, oy, oz)) {
            // view-space xy: o_world @ wvt[:3,:3]
            const float vx = ox * a.view_rot[0] + oy * a.view_rot[3] + oz * a.view_rot[6];
            const float vy = ox * a.view_rot[1] + oy * a.view_rot[4] + oz * a.view_rot[7];
            const float nrm = sqrtf(vx * vx + vy * vy);
            const float den = nrm + eps;
            const float pxn = vx / den;
            float pyn = vy / den;
            if (pyn < eps) pyn += eps;
            float theta = atan2f(pxn, pyn);
            if (theta < 0.f) theta += kPi;
            const float dt = theta - a.gt_theta[i];
            const float inner = fabsf(dt) - kPi2;
            const float diff = kPi2 - fabsf(inner);
            const float cf = a.confidence[i];
            ori_sum += diff * cf;
            // backward
            const float sgn_dt = dt > 0.f ? 1.f : (dt < 0.f ? -1.f : 0.f);
            const float sgn_in = inner > 0.f ? 1.f : (inner < 0.f ? -1.f : 0.f);
            const float g_theta = a.l_orient * inv_count * cf * (-sgn_in * sgn_dt);
            const float r2 = pxn * pxn + pyn * pyn;
            const float g_px = r2 > 0.f ? g_theta * (pyn / r2) : 0.f;   // d atan2(x, y)/dx =  y/(x^2+y^2)
            const float g_py = r2 > 0.f ? g_theta * (-pxn / r2) : 0.f;  //              /dy = -x/(x^2+y^2)
            // p = v / (|v| + eps); d|v|/dv = v/|v| (0 at the origin, as torch.norm)
            const float gdotv = g_px * vx + g_py * vy;
            const float k = nrm > 0.f ? gdotv / (den * den * nrm) : 0.f;
            const float g_vx = g_px / den - k * vx, g_vy = g_py / den - k * vy;
            gox = g_vx * a.view_rot[0] + g_vy * a.view_rot[1];
            goy = g_vx * a.view_rot[3] + g_vy * a.view_rot[4];
            goz = g_vx * a.view_rot[6] + g_vy * a.view_rot[7];
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

int launch_hair_image_loss(const HairLossArgs& a, cudaStream_t s) {
    static bool init = false;
    if (!init) {
        float g[11], sum = 0.f;
        for (int x = 0; x < 11; ++x) { g[x] = expf(-((x - 5) * (x - 5)) / (2.f * 1.5f * 1.5f)); sum += g[x]; }
        for (int x = 0; x < 11; ++x) g[x] /= sum;
        if (int e = check_cuda(cudaMemcpyToSymbol(kGauss11, g, sizeof(g)), "gaussian window")) return e;
        init = true;
    }
    if (int e = check_cuda(cudaMemsetAsync(a.terms, 0, 8 * sizeof(float), s), "memset loss terms")) return e;
    const long long HW = (long long)a.height * a.width;
    long long nb = (HW + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    StageScope prof(HGS_STAGE_OTHER, s);
    hair_loss_count_kernel<<<(unsigned)nb, 256, 0, s>>>(a);
    dim3 grid((a.width + kSsimTile - 1) / kSsimTile, (a.height + kSsimTile - 1) / kSsimTile, 3);
    ssim_fwd_kernel<<<grid, 256, 0, s>>>(a);
    ssim_bwd_kernel<<<grid, 256, 0, s>>>(a);
    hair_pointwise_kernel<<<(unsigned)nb, 256, 0, s>>>(a);
    hair_loss_finish_kernel<<<1, 1, 0, s>>>(a);
    return check_cuda(cudaGetLastError(), "hair_image_loss launch");
}

}  // namespace hgs
