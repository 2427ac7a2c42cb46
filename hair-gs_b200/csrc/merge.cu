// merge.cu — strand endpoint merge search (second half): the candidate search and the greedy
// one-to-one matching of HairGaussianModel.compute_endpoint_pair_to_merge (scene/hair_gaussian_model.py:1205-1362).
// The reference builds a scipy cKDTree on the host, queries a ball per strand end, filters the hits in a Python loop per
// point (:1294-1335) and then walks the distance-sorted pair list in another Python loop over device tensors (:1237-1255).
// Here: the strand ends (2 per strand) are few enough that an exhaustive, shared-memory-tiled all-pairs test is both exact
// and fast (K = 20 000 ends -> 4e8 tests, well under a millisecond); it also yields every point's hits already ordered
// by index, which is the order cKDTree's return_sorted=True gives and max_num_nn truncates in.  count + fill passes, so
// the output order (by p1, then p2) is deterministic; the greedy matching is one sequential device pass.
#include "hgs_common.cuh"

namespace hgs {


constexpr int kMergeTile = 256;

// hit test of (i, j) exactly as the reference filters them: inside the ball (double, <=), not itself, not the other end
// of its own strand, directions opposed within the angle threshold
__device__ __forceinline__ bool merge_hit(const MergeArgs& a, int i, int j, float3 pi, float3 di, int other_i, float3 pj,
                                          float3 dj, int gid_j) {
    if (i == j || gid_j == other_i) return false;
    const double dx = (double)pi.x - (double)pj.x, dy = (double)pi.y - (double)pj.y, dz = (double)pi.z - (double)pj.z;
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    if (!(d2 <= a.r2)) return false;
    // p1_nn_dirs @ (-p1_dir) (hair_gaussian_model.py:1315-1319)
    float dot = dj.x * -di.x;
    dot = fmaf(dj.y, -di.y, dot);
    dot = fmaf(dj.z, -di.z, dot);
    if (a.bidirectional) dot = fabsf(dot);
    return (double)dot >= a.dir_th;
}

// FILL == false: counts[i] = number of accepted hits (after max_num_nn); FILL == true: writes them at offsets[i]
template <bool FILL>
__global__ void __launch_bounds__(kMergeTile) merge_candidates_kernel(const MergeArgs a, int* __restrict__ counts,
                                                                      const long long* __restrict__ offsets,
                                                                      int* __restrict__ p1, int* __restrict__ p2,
                                                                      float* __restrict__ dist) {
    __shared__ float s_p[kMergeTile][3];
    __shared__ float s_d[kMergeTile][3];
    __shared__ int s_g[kMergeTile];
    const int i = blockIdx.x * kMergeTile + threadIdx.x;
    const bool live = i < a.K;
    float3 pi = make_float3(0, 0, 0), di = make_float3(0, 0, 0);
    int other_i = -1, gid_i = -1;
    if (live) {
        pi = make_float3(a.points[3 * i], a.points[3 * i + 1], a.points[3 * i + 2]);
        di = make_float3(a.dirs[3 * i], a.dirs[3 * i + 1], a.dirs[3 * i + 2]);
        other_i = a.other_end[i];
        gid_i = a.global_id[i];
    }
    int n = 0;
    long long out = (FILL && live) ? offsets[i] : 0;
    for (int j0 = 0; j0 < a.K; j0 += kMergeTile) {
        const int j = j0 + threadIdx.x;
        __syncthreads();
        if (j < a.K) {
            s_p[threadIdx.x][0] = a.points[3 * j]; s_p[threadIdx.x][1] = a.points[3 * j + 1]; s_p[threadIdx.x][2] = a.points[3 * j + 2];
            s_d[threadIdx.x][0] = a.dirs[3 * j]; s_d[threadIdx.x][1] = a.dirs[3 * j + 1]; s_d[threadIdx.x][2] = a.dirs[3 * j + 2];
            s_g[threadIdx.x] = a.global_id[j];
        }
        __syncthreads();
        if (!live) continue;
        const int lim = min(kMergeTile, a.K - j0);
        for (int t = 0; t < lim; ++t) {
            if (a.max_num_nn > 0 && n >= a.max_num_nn) break;
            const float3 pj = make_float3(s_p[t][0], s_p[t][1], s_p[t][2]);
            // cheap float reject before the exact double test (1e-3 relative slack keeps it conservative)
            const float fx = pi.x - pj.x, fy = pi.y - pj.y, fz = pi.z - pj.z;
            if ((double)(fx * fx + fy * fy + fz * fz) > a.r2 * 1.001 + 1e-30) continue;
            const float3 dj = make_float3(s_d[t][0], s_d[t][1], s_d[t][2]);
            if (!merge_hit(a, i, j0 + t, pi, di, other_i, pj, dj, s_g[t])) continue;
            if (FILL) {
                p1[out] = gid_i;
                p2[out] = s_g[t];
                // np.linalg.norm of the float32 difference (:1321-1324): sqrt((dx^2 + dy^2) + dz^2) in float32, unfused
                dist[out] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)), __fmul_rn(fz, fz)));
                ++out;
            }
            ++n;
        }
    }
    if (!FILL && live) counts[i] = n;
}

// Sequential pass over the distance-sorted pairs:
//   remove_duplicate_endpoint_rows (:711-726): a row survives only if BOTH its ids occur for the first time in the
//     row-major flattened list (ids are marked seen whether or not the row survives);
//   remove_complementary_rows (:1237-1255): walking the survivors in order, a row is dropped if either id was disabled
//     by an earlier kept row, otherwise it is kept and the other strand end of both ids is disabled.
// flags: one byte per endpoint id, bit 0 = seen, bit 1 = disabled; must be zero on entry.
__global__ void merge_greedy_kernel(long long n, const int* __restrict__ p1, const int* __restrict__ p2,
                                    const int* __restrict__ other_end_of, unsigned char* __restrict__ flags,
                                    unsigned char* __restrict__ keep) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (long long r = 0; r < n; ++r) {
        const int u = p1[r], v = p2[r];
        const unsigned char fu = flags[u];
        flags[u] = fu | 1;
        const unsigned char fv = flags[v];
        flags[v] = fv | 1;
        bool k = !(fu & 1) && !(fv & 1);
        if (k) {
            // re-read: u's or v's disabled bit may have been set by an earlier kept row
            if ((flags[u] & 2) || (flags[v] & 2)) {
                k = false;
            } else {
                const int cu = other_end_of[u], cv = other_end_of[v];
                if (cu >= 0) flags[cu] |= 2;
                if (cv >= 0) flags[cv] |= 2;
            }
        }
        keep[r] = k ? 1 : 0;
    }
}

int launch_merge_candidates(const MergeArgs& a, bool fill, int* counts, const long long* offsets, int* p1, int* p2,
                            float* dist, cudaStream_t s) {
    if (a.K <= 0) return HGS_OK;
    const unsigned nb = (unsigned)((a.K + kMergeTile - 1) / kMergeTile);
    StageScope prof(HGS_STAGE_OTHER, s);
    if (fill)
        merge_candidates_kernel<true><<<nb, kMergeTile, 0, s>>>(a, counts, offsets, p1, p2, dist);
    else
        merge_candidates_kernel<false><<<nb, kMergeTile, 0, s>>>(a, counts, offsets, p1, p2, dist);
    return check_cuda(cudaGetLastError(), "merge_candidates launch");
}

int launch_merge_greedy(long long n, const int* p1, const int* p2, const int* other_end_of, unsigned char* flags,
                        unsigned char* keep, cudaStream_t s) {
    if (n <= 0) return HGS_OK;
    StageScope prof(HGS_STAGE_OTHER, s);
    merge_greedy_kernel<<<1, 32, 0, s>>>(n, p1, p2, other_end_of, flags, keep);
    return check_cuda(cudaGetLastError(), "merge_greedy launch");
}

}  // namespace hgs
