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

}  // namespace hgs
