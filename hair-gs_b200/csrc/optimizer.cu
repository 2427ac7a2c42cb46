// optimizer.cu — SURVEY.md §8f row N4: the per-iteration optimiser work of Hair-GS on the flat parameter bucket.
//   adam_flat_kernel     : torch.optim.Adam(lr per group, betas (0.9, 0.999), eps 1e-15, no weight decay / amsgrad), the
//                          optimiser scene/gaussian_model.py:250 builds, over ALL parameter groups in one launch.  The
//                          reference steps 7 groups through torch's foreach path (several launches per group) and calls
//                          zero_grad(set_to_none) afterwards; here the gradient is cleared by the same kernel.
//   densify_stats_kernel : update_densification_stats (scene/gaussian_model.py:675-682): max_radii2D, the accumulated
//                          screen-space gradient norm and the visibility count, one pass, no boolean-index temporaries.
// HBM-bound streaming kernels: 28 B (+4 B for the gradient clear) per parameter element.
#include "hgs_common.cuh"

namespace hgs {

struct AdamGroups {
    int n;
    long long end[HGS_ADAM_MAX_GROUPS];  // exclusive end offset of group g in the flat bucket
    float step_size[HGS_ADAM_MAX_GROUPS];  // lr_g / (1 - beta1^t)
};

__device__ __forceinline__ float adam_one(float p, float g, float& m, float& v, float beta1, float beta2, float step_size,
                                          float inv_bc2_sqrt, float eps) {
    // exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
    m = m + (g - m) * (1.f - beta1);
    v = v * beta2 + (1.f - beta2) * g * g;
    const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
    return p - step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adam_flat_kernel(long long n, float* __restrict__ param, float* __restrict__ grad,
                                                        float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                        const AdamGroups groups, float beta1, float beta2,
                                                        float inv_bc2_sqrt, float eps, float grad_scale, int zero_grad) {
    const long long n4 = n / 4;
    float4* p4 = reinterpret_cast<float4*>(param);
    float4* g4 = reinterpret_cast<float4*>(grad);
    float4* m4 = reinterpret_cast<float4*>(exp_avg);
    float4* v4 = reinterpret_cast<float4*>(exp_avg_sq);
    auto step_of = [&](long long i) {
        int g = 0;
        while (g + 1 < groups.n && i >= groups.end[g]) ++g;
        return groups.step_size[g];
    };
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 p = p4[i], g = g4[i], m = m4[i], v = v4[i];
        const long long e = i * 4;
        const float s0 = step_of(e), s3 = step_of(e + 3);
        const float s1 = s0 == s3 ? s0 : step_of(e + 1), s2 = s0 == s3 ? s0 : step_of(e + 2);
        p.x = adam_one(p.x, g.x * grad_scale, m.x, v.x, beta1, beta2, s0, inv_bc2_sqrt, eps);
        p.y = adam_one(p.y, g.y * grad_scale, m.y, v.y, beta1, beta2, s1, inv_bc2_sqrt, eps);
        p.z = adam_one(p.z, g.z * grad_scale, m.z, v.z, beta1, beta2, s2, inv_bc2_sqrt, eps);
        p.w = adam_one(p.w, g.w * grad_scale, m.w, v.w, beta1, beta2, s3, inv_bc2_sqrt, eps);
        p4[i] = p; m4[i] = m; v4[i] = v;
        if (zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (blockIdx.x == 0) {
        for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
            float m = exp_avg[i], v = exp_avg_sq[i];
            param[i] = adam_one(param[i], grad[i] * grad_scale, m, v, beta1, beta2, step_of(i), inv_bc2_sqrt, eps);
            exp_avg[i] = m; exp_avg_sq[i] = v;
            if (zero_grad) grad[i] = 0.f;
        }
    }
}

int launch_adam_flat(long long n, float* param, float* grad, float* m, float* v, int n_groups, const int64_t* group_end,
                     const float* lr, int step, float beta1, float beta2, float eps, float grad_scale, int zero_grad,
                     cudaStream_t s) {
    if (n == 0) return HGS_OK;
    if ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) % 16) != 0) {
        set_error("adam_step needs 16-byte aligned flat buffers"); return HGS_ERR_INVALID;
    }
    AdamGroups g;
    g.n = n_groups;
    // bias corrections in double on the host like torch's python scalars (optim/adam.py _single_tensor_adam)
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    for (int i = 0; i < n_groups; ++i) { g.end[i] = group_end[i]; g.step_size[i] = (float)((double)lr[i] / bc1); }
    for (int i = n_groups; i < HGS_ADAM_MAX_GROUPS; ++i) { g.end[i] = n; g.step_size[i] = 0.f; }
    long long nb = (n / 4 + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    if (nb < 1) nb = 1;
    StageScope prof(HGS_STAGE_OTHER, s);
    adam_flat_kernel<<<(unsigned)nb, 256, 0, s>>>(n, param, grad, m, v, g, beta1, beta2, (float)(1.0 / sqrt(bc2)), eps,
                                                  grad_scale, zero_grad);
    return check_cuda(cudaGetLastError(), "adam_step launch");
}

// visible = radii > 0 (update_filter = visibility_filter, train.py:169-176)
__global__ void __launch_bounds__(256) densify_stats_kernel(int P, const int* __restrict__ radii,
                                                            const float* __restrict__ dL_dmean2D, int grad_stride,
                                                            int* __restrict__ max_radii2D_i, float* __restrict__ max_radii2D_f,
                                                            float* __restrict__ grad_accum, float* __restrict__ denom) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int r = radii[i];
    if (r <= 0) return;
    if (max_radii2D_f) max_radii2D_f[i] = fmaxf(max_radii2D_f[i], (float)r);
    if (max_radii2D_i) max_radii2D_i[i] = max(max_radii2D_i[i], r);
    const float gx = dL_dmean2D[(size_t)i * grad_stride], gy = dL_dmean2D[(size_t)i * grad_stride + 1];
    grad_accum[i] += sqrtf(gx * gx + gy * gy);
    denom[i] += 1.f;
}

int launch_densify_stats(int P, const int* radii, const float* dL_dmean2D, int grad_stride, int* max_i, float* max_f,
                         float* grad_accum, float* denom, cudaStream_t s) {
    if (P == 0) return HGS_OK;
    StageScope prof(HGS_STAGE_OTHER, s);
    densify_stats_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, radii, dL_dmean2D, grad_stride, max_i, max_f, grad_accum, denom);
    return check_cuda(cudaGetLastError(), "densify_stats launch");
}

}  // namespace hgs
