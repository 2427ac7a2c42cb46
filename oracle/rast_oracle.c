/*
 * rast_oracle.c — CPU restatement of the Hair-GS render path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may build,
 * load or call this file.  The product (hair-gs_b200/) never does: it has no CPU path at all.
 *
 * What it restates (file:line in /root/reference/submodules/diff-gaussian-rasterization unless noted):
 *   orc_forward   preprocess   cuda_rasterizer/forward.cu:155-256 with auxiliary.h:41-77,139-164,
 *                              computeCov3D forward.cu:118-152, computeCov2D :74-113, SH :20-71
 *                 binning      rasterizer_impl.cu:70-111 (keys), :277 (inclusive scan), :300-308 (stable
 *                              sort on bits [0,32+getHigherMsb(T))), :116-138 (tile ranges)
 *                 compositor   forward.cu:261-374
 *   orc_backward  compositor   cuda_rasterizer/backward_distwar.cu:855-1014 (original variant: the three
 *                              variants compute the same sums, only the reduction order differs)
 *                 preprocess   backward_distwar.cu:145-275, :347-397, :21-140, :279-342
 *   orc_mark_visible           rasterizer_impl.cu:54-66
 *   orc_knn3      simple-knn/simple_knn.cu:133-184 semantics (exact 3-NN, self excluded by index,
 *                 mean of the three smallest squared distances) by brute force
 *   orc_higher_msb             rasterizer_impl.cu:35-50
 *
 * Numerics.  The reference is CUDA-only; its float results depend on nvcc's FMA contraction.  The forward
 * preprocess below spells every contraction out with fmaf() following the rule nvcc applies to the
 * reference's expression trees ("a*b + c*d" -> fma(a,b,rn(c*d)); sums are left-associated), so that
 * radii, tile rects, depth bits, means2D, conics and cov3D reproduce the B200 build bit for bit — this is
 * pinned against fixtures recorded from the real reference on a B200 (tests/golden/, tests/test_oracle.py).
 * expf() differs from CUDA's (MUFU.EX2 based) by <= 2 ulp, so pixels / final_T are tolerance-checked.
 * Compile WITHOUT fast-math and WITH -ffp-contract=off (see oracle/Makefile).
 *
 * Parity status: pinned by golden vectors generated on the GPU box from oracle/_ref (the reference's own
 * tree holds no tests or fixtures: SURVEY.md §4).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16
#define MAXC 8

typedef struct orc_params {
    int32_t P, D, M, W, H, C;
    float tan_fovx, tan_fovy, scale_modifier;
} orc_params;

typedef struct orc_state {
    orc_params prm;
    int64_t N;
    int32_t* radii;          /* [P] */
    float* depths;           /* [P] */
    float* means2D;          /* [P,2] */
    float* cov3D;            /* [P,6] */
    float* conic_opacity;    /* [P,4] */
    float* rgb;              /* [P,C] */
    uint8_t* clamped;        /* [P,3] */
    uint32_t* tiles_touched; /* [P] */
    uint32_t* point_offsets; /* [P] inclusive */
    uint64_t* keys_unsorted; /* [N] */
    uint32_t* vals_unsorted; /* [N] */
    uint64_t* keys;          /* [N] sorted */
    uint32_t* point_list;    /* [N] */
    uint32_t* ranges;        /* [T,2] */
    float* final_T;          /* [H*W] */
    uint32_t* n_contrib;     /* [H*W] */
} orc_state;

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f,  -0.5900435899266435f};

/* float -> int with CUDA's cvt.rzi.s32.f32 semantics (saturating, NaN -> 0) */
static inline int f2i(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (-2147483647 - 1);
    return (int)f;
}
static inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
/* CUDA fminf/fmaxf: return the non-NaN operand */
static inline float cmin(float a, float b) { return fminf(a, b); }
static inline float cmax(float a, float b) { return fmaxf(a, b); }

/* rasterizer_impl.cu:35-50 */
uint32_t orc_higher_msb(uint32_t n) {
    uint32_t msb = sizeof(n) * 4;
    uint32_t step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

/* m0 x + m4 y + m8 z + m12 as nvcc contracts it: ((m0x + m4y) + m8z) + m12 */
static inline float row_affine(const float* m, int r, float x, float y, float z) {
    return fmaf(m[8 + r], z, fmaf(m[r], x, m[4 + r] * y)) + m[12 + r];
}
/* a0*b0 + a1*b1 + a2*b2, left-associated, contracted */
static inline float dot3c(float a0, float b0, float a1, float b1, float a2, float b2) {
    return fmaf(a2, b2, fmaf(a0, b0, a1 * b1));
}

/* forward.cu:118-152 */
static void cov3d_from_scale_rot(const float* s3, float mod, const float* q4, float* cov3D) {
    const float sx = mod * s3[0], sy = mod * s3[1], sz = mod * s3[2];
    const float r = q4[0], x = q4[1], y = q4[2], z = q4[3];
    float R[3][3]; /* R[c][r], glm column-major */
    /* which product of each pair nvcc keeps as a rounded FMUL (shared between entries) and which it fuses was
     * determined against the B200 build of the reference (tests/golden): yy, zz, xz, rz, rx are rounded. */
    const float yy = y * y, zz = z * z, xz = x * z, rz = r * z, rx = r * x;
    R[0][0] = fmaf(-2.f, yy + zz, 1.f);          R[0][1] = 2.f * fmaf(x, y, -rz);             R[0][2] = 2.f * fmaf(r, y, xz);
    R[1][0] = 2.f * fmaf(x, y, rz);              R[1][1] = fmaf(-2.f, fmaf(x, x, zz), 1.f);   R[1][2] = 2.f * fmaf(y, z, -rx);
    R[2][0] = 2.f * fmaf(-r, y, xz);             R[2][1] = 2.f * fmaf(y, z, rx);              R[2][2] = fmaf(-2.f, fmaf(x, x, yy), 1.f);
    float M[3][3]; /* M = S*R : M[c][r] = s_r * R[c][r] */
    for (int c = 0; c < 3; ++c) { M[c][0] = sx * R[c][0]; M[c][1] = sy * R[c][1]; M[c][2] = sz * R[c][2]; }
    /* Sigma = M^T M : Sigma[c][r] = sum_k M[r][k] M[c][k] */
#define SIG(c, r) dot3c(M[r][0], M[c][0], M[r][1], M[c][1], M[r][2], M[c][2])
    cov3D[0] = SIG(0, 0); cov3D[1] = SIG(0, 1); cov3D[2] = SIG(0, 2);
    cov3D[3] = SIG(1, 1); cov3D[4] = SIG(1, 2); cov3D[5] = SIG(2, 2);
#undef SIG
}

/* forward.cu:74-113; returns cov2D (a,b,c) incl. +0.3, T[2][3] = rows 0/1 of W*J as T[c][r], clamped t */
static void cov2d_ewa(const float* mean, float fx, float fy, float tan_fovx, float tan_fovy, const float* cov3D,
                      const float* view, float* cov_out, float T[3][3], float* t_out, float* txtz_out,
                      float* tytz_out) {
    float t[3] = {row_affine(view, 0, mean[0], mean[1], mean[2]), row_affine(view, 1, mean[0], mean[1], mean[2]),
                  row_affine(view, 2, mean[0], mean[1], mean[2])};
    const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
    const float txtz = t[0] / t[2], tytz = t[1] / t[2];
    t[0] = cmin(limx, cmax(-limx, txtz)) * t[2];
    t[1] = cmin(limy, cmax(-limy, tytz)) * t[2];
    const float J00 = fx / t[2], J02 = -(fx * t[0]) / (t[2] * t[2]);
    const float J11 = fy / t[2], J12 = -(fy * t[1]) / (t[2] * t[2]);
    /* T = W*J, W[k][r] = view[4r+k]; zero entries of J contribute exact zeros */
    for (int r = 0; r < 3; ++r) {
        T[0][r] = fmaf(view[4 * r + 2], J02, view[4 * r] * J00);
        T[1][r] = fmaf(view[4 * r + 2], J12, view[4 * r + 1] * J11);
        T[2][r] = 0.f;
    }
    float V[3][3] = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
    /* A = T^T V^T : A[c][r] = sum_k T[r][k] V[k][c];  cov = A T : cov[c][r] = sum_k A[k][r] T[c][k] */
    float A[3][3];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) A[c][r] = dot3c(T[r][0], V[0][c], T[r][1], V[1][c], T[r][2], V[2][c]);
#define COV(c, r) dot3c(A[0][r], T[c][0], A[1][r], T[c][1], A[2][r], T[c][2])
    cov_out[0] = COV(0, 0) + 0.3f;
    cov_out[1] = COV(0, 1);
    cov_out[2] = COV(1, 1) + 0.3f;
#undef COV
    t_out[0] = t[0]; t_out[1] = t[1]; t_out[2] = t[2];
    *txtz_out = txtz; *tytz_out = tytz;
}

static inline float ndc2pix(float v, int S) { return (float)(fma((double)v + 1.0, (double)S, -1.0) * 0.5); }

/* auxiliary.h:46-56 */
static void get_rect(float px, float py, int r, uint32_t gx, uint32_t gy, uint32_t* rmin, uint32_t* rmax) {
    rmin[0] = (uint32_t)imin((int)gx, imax(0, f2i((px - (float)r) / 16.0f)));
    rmin[1] = (uint32_t)imin((int)gy, imax(0, f2i((py - (float)r) / 16.0f)));
    rmax[0] = (uint32_t)imin((int)gx, imax(0, f2i((px + (float)r + 16.0f - 1.0f) / 16.0f)));
    rmax[1] = (uint32_t)imin((int)gy, imax(0, f2i((py + (float)r + 16.0f - 1.0f) / 16.0f)));
}

/* forward.cu:20-71 */
static void sh_to_rgb(int deg, const float* sh, const float* pos, const float* campos, float* out, uint8_t* clamped) {
    float dx = pos[0] - campos[0], dy = pos[1] - campos[1], dz = pos[2] - campos[2];
    const float len = sqrtf(fmaf(dz, dz, fmaf(dx, dx, dy * dy)));
    const float x = dx / len, y = dy / len, z = dz / len;
    for (int c = 0; c < 3; ++c) {
#define SH(k) sh[(k) * 3 + c]
        float res = SH_C0 * SH(0);
        if (deg > 0) {
            res = fmaf(-(SH_C1 * x), SH(3), fmaf(SH_C1 * z, SH(2), fmaf(-(SH_C1 * y), SH(1), res)));
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                res = fmaf(SH_C2[0] * xy, SH(4), res);
                res = fmaf(SH_C2[1] * yz, SH(5), res);
                res = fmaf(SH_C2[2] * (fmaf(2.0f, zz, -xx) - yy), SH(6), res);
                res = fmaf(SH_C2[3] * xz, SH(7), res);
                res = fmaf(SH_C2[4] * (xx - yy), SH(8), res);
                if (deg > 2) {
                    res = fmaf(SH_C3[0] * y * (fmaf(3.0f, xx, -yy)), SH(9), res);
                    res = fmaf(SH_C3[1] * xy * z, SH(10), res);
                    res = fmaf(SH_C3[2] * y * (fmaf(4.0f, zz, -xx) - yy), SH(11), res);
                    res = fmaf(SH_C3[3] * z * (fmaf(-3.0f, yy, fmaf(2.0f, zz, -(3.0f * xx)))), SH(12), res);
                    res = fmaf(SH_C3[4] * x * (fmaf(4.0f, zz, -xx) - yy), SH(13), res);
                    res = fmaf(SH_C3[5] * z * (xx - yy), SH(14), res);
                    res = fmaf(SH_C3[6] * x * (fmaf(-3.0f, yy, xx)), SH(15), res);
                }
            }
        }
#undef SH
        res += 0.5f;
        clamped[c] = (res < 0);
        out[c] = cmax(res, 0.0f);
    }
}

static int cmp_u64(const void* a, const void* b) { (void)a; (void)b; return 0; }

/* stable LSD radix sort on key bits [0,end_bit) — any stable sort on those bits gives the same permutation
 * as cub::DeviceRadixSort::SortPairs (rasterizer_impl.cu:303-308) */
void orc_sort_pairs(int64_t n, int end_bit, uint64_t* keys, uint32_t* vals) {
    (void)cmp_u64;
    if (n <= 1) return;
    uint64_t* k2 = (uint64_t*)malloc((size_t)n * 8);
    uint32_t* v2 = (uint32_t*)malloc((size_t)n * 4);
    uint64_t *ka = keys, *kb = k2;
    uint32_t *va = vals, *vb = v2;
    for (int shift = 0; shift < end_bit; shift += 8) {
        int bits = end_bit - shift < 8 ? end_bit - shift : 8;
        uint64_t mask = (1ull << bits) - 1;
        size_t cnt[257];
        memset(cnt, 0, sizeof(cnt));
        for (int64_t i = 0; i < n; ++i) cnt[((ka[i] >> shift) & mask) + 1]++;
        for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
        for (int64_t i = 0; i < n; ++i) {
            size_t p = cnt[(ka[i] >> shift) & mask]++;
            kb[p] = ka[i];
            vb[p] = va[i];
        }
        uint64_t* tk = ka; ka = kb; kb = tk;
        uint32_t* tv = va; va = vb; vb = tv;
    }
    if (ka != keys) { memcpy(keys, ka, (size_t)n * 8); memcpy(vals, va, (size_t)n * 4); }
    free(k2); free(v2);
}

void orc_free(orc_state* s) {
    if (!s) return;
    free(s->radii); free(s->depths); free(s->means2D); free(s->cov3D); free(s->conic_opacity); free(s->rgb);
    free(s->clamped); free(s->tiles_touched); free(s->point_offsets); free(s->keys_unsorted); free(s->vals_unsorted);
    free(s->keys); free(s->point_list); free(s->ranges); free(s->final_T); free(s->n_contrib);
    free(s);
}

int64_t orc_num_rendered(const orc_state* s) { return s->N; }
const void* orc_array(const orc_state* s, int what) {
    switch (what) {
        case 0: return s->depths; case 1: return s->means2D; case 2: return s->conic_opacity; case 3: return s->rgb;
        case 4: return s->tiles_touched; case 5: return s->point_offsets; case 6: return s->clamped;
        case 7: return s->keys; case 8: return s->point_list; case 9: return s->ranges; case 10: return s->final_T;
        case 11: return s->n_contrib; case 12: return s->keys_unsorted; case 13: return s->cov3D;
        case 14: return s->radii; case 15: return s->vals_unsorted;
    }
    return NULL;
}

orc_state* orc_forward(const orc_params* prm, const float* bg, const float* means3D, const float* shs,
                       const float* colors_precomp, const float* opacities, const float* scales,
                       const float* rotations, const float* cov3D_precomp, const float* view, const float* proj,
                       const float* campos, float* out_color, int32_t* radii_out) {
    const int P = prm->P, W = prm->W, H = prm->H, C = prm->C;
    const uint32_t gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const size_t T = (size_t)gx * gy, HW = (size_t)W * H;
    const float focal_y = H / (2.0f * prm->tan_fovy), focal_x = W / (2.0f * prm->tan_fovx);
    orc_state* s = (orc_state*)calloc(1, sizeof(orc_state));
    s->prm = *prm;
    size_t Pz = P > 0 ? P : 1;
    s->radii = (int32_t*)calloc(Pz, 4); s->depths = (float*)calloc(Pz, 4); s->means2D = (float*)calloc(Pz, 8);
    s->cov3D = (float*)calloc(Pz, 24); s->conic_opacity = (float*)calloc(Pz, 16); s->rgb = (float*)calloc(Pz * C, 4);
    s->clamped = (uint8_t*)calloc(Pz, 3); s->tiles_touched = (uint32_t*)calloc(Pz, 4);
    s->point_offsets = (uint32_t*)calloc(Pz, 4);
    s->ranges = (uint32_t*)calloc(T * 2, 4); s->final_T = (float*)calloc(HW, 4); s->n_contrib = (uint32_t*)calloc(HW, 4);

    /* ---- preprocess (forward.cu:182-255) ---- */
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; ++idx) {
        const float* p = means3D + 3 * (size_t)idx;
        const float hx = row_affine(proj, 0, p[0], p[1], p[2]), hy = row_affine(proj, 1, p[0], p[1], p[2]);
        const float hw = row_affine(proj, 3, p[0], p[1], p[2]);
        const float p_w = 1.0f / (hw + 0.0000001f);
        const float projx = hx * p_w, projy = hy * p_w;
        const float vz = row_affine(view, 2, p[0], p[1], p[2]);
        if (vz <= 0.2f) continue;
        float* cov3D = s->cov3D + 6 * (size_t)idx;
        if (cov3D_precomp) memcpy(cov3D, cov3D_precomp + 6 * (size_t)idx, 24);
        else cov3d_from_scale_rot(scales + 3 * (size_t)idx, prm->scale_modifier, rotations + 4 * (size_t)idx, cov3D);
        float cov[3], Tm[3][3], t[3], txtz, tytz;
        cov2d_ewa(p, focal_x, focal_y, prm->tan_fovx, prm->tan_fovy, cov3D, view, cov, Tm, t, &txtz, &tytz);
        const float det = fmaf(cov[0], cov[2], -(cov[1] * cov[1]));
        if (det == 0.0f) continue;
        const float det_inv = 1.f / det;
        const float conic[3] = {cov[2] * det_inv, -cov[1] * det_inv, cov[0] * det_inv};
        const float mid = 0.5f * (cov[0] + cov[2]);
        const float disc = sqrtf(cmax(0.1f, fmaf(mid, mid, -det)));
        const float lambda1 = mid + disc, lambda2 = mid - disc;
        const float my_radius = ceilf(3.f * sqrtf(cmax(lambda1, lambda2)));
        const float px = ndc2pix(projx, W), py = ndc2pix(projy, H);
        uint32_t rmin[2], rmax[2];
        get_rect(px, py, f2i(my_radius), gx, gy, rmin, rmax);
        if ((rmax[0] - rmin[0]) * (rmax[1] - rmin[1]) == 0) continue;
        if (!colors_precomp) {
            sh_to_rgb(prm->D, shs + (size_t)idx * prm->M * 3, p, campos, s->rgb + (size_t)idx * C,
                      s->clamped + 3 * (size_t)idx);
        } else {
            for (int c = 0; c < C; ++c) s->rgb[(size_t)idx * C + c] = colors_precomp[(size_t)idx * C + c];
        }
        s->depths[idx] = vz;
        s->radii[idx] = f2i(my_radius);
        s->means2D[2 * (size_t)idx] = px; s->means2D[2 * (size_t)idx + 1] = py;
        float* co = s->conic_opacity + 4 * (size_t)idx;
        co[0] = conic[0]; co[1] = conic[1]; co[2] = conic[2]; co[3] = opacities[idx];
        s->tiles_touched[idx] = (rmax[1] - rmin[1]) * (rmax[0] - rmin[0]);
    }
    if (radii_out) memcpy(radii_out, s->radii, (size_t)P * 4);

    /* ---- inclusive scan, key emission, stable sort, ranges (rasterizer_impl.cu:277-318) ---- */
    uint64_t run = 0;
    for (int i = 0; i < P; ++i) { run += s->tiles_touched[i]; s->point_offsets[i] = (uint32_t)run; }
    const int64_t N = (int64_t)run;
    s->N = N;
    size_t Nz = N > 0 ? (size_t)N : 1;
    s->keys_unsorted = (uint64_t*)malloc(Nz * 8); s->vals_unsorted = (uint32_t*)malloc(Nz * 4);
    s->keys = (uint64_t*)malloc(Nz * 8); s->point_list = (uint32_t*)malloc(Nz * 4);
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; ++idx) {
        if (s->radii[idx] <= 0) continue;
        uint32_t off = idx == 0 ? 0 : s->point_offsets[idx - 1];
        uint32_t rmin[2], rmax[2];
        get_rect(s->means2D[2 * (size_t)idx], s->means2D[2 * (size_t)idx + 1], s->radii[idx], gx, gy, rmin, rmax);
        for (uint32_t y = rmin[1]; y < rmax[1]; ++y)
            for (uint32_t x = rmin[0]; x < rmax[0]; ++x) {
                uint64_t key = (uint64_t)(y * gx + x);
                key <<= 32;
                key |= fbits(s->depths[idx]);
                s->keys_unsorted[off] = key;
                s->vals_unsorted[off] = (uint32_t)idx;
                off++;
            }
    }
    memcpy(s->keys, s->keys_unsorted, (size_t)N * 8);
    memcpy(s->point_list, s->vals_unsorted, (size_t)N * 4);
    orc_sort_pairs(N, 32 + (int)orc_higher_msb((uint32_t)T), s->keys, s->point_list);
    for (int64_t i = 0; i < N; ++i) {
        uint32_t cur = (uint32_t)(s->keys[i] >> 32);
        if (i == 0) s->ranges[2 * cur] = 0;
        else {
            uint32_t prev = (uint32_t)(s->keys[i - 1] >> 32);
            if (cur != prev) { s->ranges[2 * prev + 1] = (uint32_t)i; s->ranges[2 * cur] = (uint32_t)i; }
        }
        if (i == N - 1) s->ranges[2 * cur + 1] = (uint32_t)N;
    }

    /* ---- compositor (forward.cu:275-373) ---- */
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t tile = 0; tile < (int64_t)T; ++tile) {
        const uint32_t ty = (uint32_t)(tile / gx), tx = (uint32_t)(tile % gx);
        const uint32_t r0 = s->ranges[2 * tile], r1 = s->ranges[2 * tile + 1];
        for (uint32_t ly = 0; ly < TILE; ++ly)
            for (uint32_t lx = 0; lx < TILE; ++lx) {
                const uint32_t pxi = tx * TILE + lx, pyi = ty * TILE + ly;
                if (pxi >= (uint32_t)W || pyi >= (uint32_t)H) continue;
                const size_t pix = (size_t)W * pyi + pxi;
                const float pfx = (float)pxi, pfy = (float)pyi;
                float Tr = 1.0f, Cc[MAXC] = {0};
                uint32_t contributor = 0, last = 0;
                for (uint32_t k = r0; k < r1; ++k) {
                    contributor++;
                    const uint32_t id = s->point_list[k];
                    const float* co = s->conic_opacity + 4 * (size_t)id;
                    const float dx = s->means2D[2 * (size_t)id] - pfx, dy = s->means2D[2 * (size_t)id + 1] - pfy;
                    /* -0.5f*(A dx dx + C dy dy) - B dx dy as nvcc contracts it */
                    const float power = fmaf(fmaf(co[0] * dx, dx, (co[2] * dy) * dy), -0.5f, -((co[1] * dx) * dy));
                    if (power > 0.0f) continue;
                    const float alpha = cmin(0.99f, co[3] * expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = Tr * (1 - alpha);
                    if (test_T < 0.0001f) break;
                    for (int ch = 0; ch < C; ++ch) Cc[ch] = fmaf(s->rgb[(size_t)id * C + ch] * alpha, Tr, Cc[ch]);
                    Tr = test_T;
                    last = contributor;
                }
                s->final_T[pix] = Tr;
                s->n_contrib[pix] = last;
                for (int ch = 0; ch < C; ++ch) out_color[(size_t)ch * HW + pix] = fmaf(Tr, bg[ch], Cc[ch]);
            }
    }
    return s;
}

/* ---- backward ------------------------------------------------------------------------------ */
static inline void atomic_addd(double* p, double v) {
#pragma omp atomic
    *p += v;
}

void orc_backward(const orc_state* s, const float* bg, const float* means3D, const float* shs,
                  const float* colors_precomp, const float* scales, const float* rotations,
                  const float* cov3D_precomp, const float* view, const float* proj, const float* campos,
                  const float* dL_dpix, float* dL_dmean2D /*[P,3]*/, float* dL_dconic /*[P,4]*/,
                  float* dL_dopacity /*[P]*/, float* dL_dcolor /*[P,C]*/, float* dL_dmean3D /*[P,3]*/,
                  float* dL_dcov3D /*[P,6]*/, float* dL_dsh /*[P,M,3]*/, float* dL_dscale /*[P,3]*/,
                  float* dL_drot /*[P,4]*/) {
    const orc_params* prm = &s->prm;
    const int P = prm->P, W = prm->W, H = prm->H, C = prm->C, M = prm->M, D = prm->D;
    const uint32_t gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const size_t T = (size_t)gx * gy, HW = (size_t)W * H;
    const float focal_y = H / (2.0f * prm->tan_fovy), focal_x = W / (2.0f * prm->tan_fovx);
    const int V = 6 + C;
    double* acc = (double*)calloc((size_t)(P > 0 ? P : 1) * V, sizeof(double));

    /* compositor backward (backward_distwar.cu:855-1014), sums accumulated in double */
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t tile = 0; tile < (int64_t)T; ++tile) {
        const uint32_t ty = (uint32_t)(tile / gx), tx = (uint32_t)(tile % gx);
        const uint32_t r0 = s->ranges[2 * tile], r1 = s->ranges[2 * tile + 1];
        const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
        for (uint32_t ly = 0; ly < TILE; ++ly)
            for (uint32_t lx = 0; lx < TILE; ++lx) {
                const uint32_t pxi = tx * TILE + lx, pyi = ty * TILE + ly;
                if (pxi >= (uint32_t)W || pyi >= (uint32_t)H) continue;
                const size_t pix = (size_t)W * pyi + pxi;
                const float pfx = (float)pxi, pfy = (float)pyi;
                const float T_final = s->final_T[pix];
                float Tr = T_final;
                const uint32_t last_contributor = s->n_contrib[pix];
                float accum_rec[MAXC] = {0}, last_color[MAXC] = {0}, dpx[MAXC];
                float last_alpha = 0.f, bg_dot = 0.f;
                for (int ch = 0; ch < C; ++ch) dpx[ch] = dL_dpix[(size_t)ch * HW + pix];
                for (int ch = 0; ch < C; ++ch) bg_dot += bg[ch] * dpx[ch];
                for (uint32_t k = r0 + last_contributor; k-- > r0;) {
                    const uint32_t id = s->point_list[k];
                    const float* co = s->conic_opacity + 4 * (size_t)id;
                    const float dx = s->means2D[2 * (size_t)id] - pfx, dy = s->means2D[2 * (size_t)id + 1] - pfy;
                    const float power = fmaf(fmaf(co[0] * dx, dx, (co[2] * dy) * dy), -0.5f, -((co[1] * dx) * dy));
                    if (power > 0.0f) continue;
                    const float G = expf(power);
                    const float alpha = cmin(0.99f, co[3] * G);
                    if (alpha < 1.0f / 255.0f) continue;
                    Tr = Tr / (1.f - alpha);
                    const float dchannel_dcolor = alpha * Tr;
                    float dL_dalpha = 0.0f;
                    double* a = acc + (size_t)id * V;
                    for (int ch = 0; ch < C; ++ch) {
                        const float c = s->rgb[(size_t)id * C + ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                        last_color[ch] = c;
                        dL_dalpha += (c - accum_rec[ch]) * dpx[ch];
                        atomic_addd(a + 6 + ch, (double)(dchannel_dcolor * dpx[ch]));
                    }
                    dL_dalpha *= Tr;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                    const float dL_dG = co[3] * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * co[0] - gdy * co[1];
                    const float dG_ddely = -gdy * co[2] - gdx * co[1];
                    atomic_addd(a + 0, (double)(dL_dG * dG_ddelx * ddelx_dx));
                    atomic_addd(a + 1, (double)(dL_dG * dG_ddely * ddely_dy));
                    atomic_addd(a + 2, (double)(-0.5f * gdx * dx * dL_dG));
                    atomic_addd(a + 3, (double)(-0.5f * gdx * dy * dL_dG));
                    atomic_addd(a + 4, (double)(-0.5f * gdy * dy * dL_dG));
                    atomic_addd(a + 5, (double)(G * dL_dalpha));
                }
            }
    }
    for (int i = 0; i < P; ++i) {
        const double* a = acc + (size_t)i * V;
        dL_dmean2D[3 * (size_t)i] = (float)a[0]; dL_dmean2D[3 * (size_t)i + 1] = (float)a[1]; dL_dmean2D[3 * (size_t)i + 2] = 0.f;
        dL_dconic[4 * (size_t)i] = (float)a[2]; dL_dconic[4 * (size_t)i + 1] = (float)a[3];
        dL_dconic[4 * (size_t)i + 2] = 0.f; dL_dconic[4 * (size_t)i + 3] = (float)a[4];
        dL_dopacity[i] = (float)a[5];
        for (int ch = 0; ch < C; ++ch) dL_dcolor[(size_t)i * C + ch] = (float)a[6 + ch];
    }
    free(acc);

    /* preprocess backward */
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; ++idx) {
        float* gm = dL_dmean3D + 3 * (size_t)idx;
        float* gc = dL_dcov3D + 6 * (size_t)idx;
        float* gs = dL_dscale + 3 * (size_t)idx;
        float* gq = dL_drot + 4 * (size_t)idx;
        gm[0] = gm[1] = gm[2] = 0.f;
        for (int i = 0; i < 6; ++i) gc[i] = 0.f;
        gs[0] = gs[1] = gs[2] = 0.f;
        gq[0] = gq[1] = gq[2] = gq[3] = 0.f;
        if (dL_dsh) for (int i = 0; i < M * 3; ++i) dL_dsh[(size_t)idx * M * 3 + i] = 0.f;
        if (!(s->radii[idx] > 0)) continue;
        const float* mean = means3D + 3 * (size_t)idx;
        const float* cov3D = cov3D_precomp ? cov3D_precomp + 6 * (size_t)idx : s->cov3D + 6 * (size_t)idx;

        /* computeCov2DCUDA (backward_distwar.cu:145-275) */
        const float dcx = dL_dconic[4 * (size_t)idx], dcy = dL_dconic[4 * (size_t)idx + 1], dcz = dL_dconic[4 * (size_t)idx + 3];
        float cov[3], Tm[3][3], t[3], txtz, tytz;
        cov2d_ewa(mean, focal_x, focal_y, prm->tan_fovx, prm->tan_fovy, cov3D, view, cov, Tm, t, &txtz, &tytz);
        const float limx = 1.3f * prm->tan_fovx, limy = 1.3f * prm->tan_fovy;
        const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
        const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        const float a = cov[0], b = cov[1], c = cov[2];
        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float V3[3][3] = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
            dL_dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
            dL_db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
            gc[0] = (Tm[0][0] * Tm[0][0] * dL_da + Tm[0][0] * Tm[1][0] * dL_db + Tm[1][0] * Tm[1][0] * dL_dc);
            gc[3] = (Tm[0][1] * Tm[0][1] * dL_da + Tm[0][1] * Tm[1][1] * dL_db + Tm[1][1] * Tm[1][1] * dL_dc);
            gc[5] = (Tm[0][2] * Tm[0][2] * dL_da + Tm[0][2] * Tm[1][2] * dL_db + Tm[1][2] * Tm[1][2] * dL_dc);
            gc[1] = 2 * Tm[0][0] * Tm[0][1] * dL_da + (Tm[0][0] * Tm[1][1] + Tm[0][1] * Tm[1][0]) * dL_db + 2 * Tm[1][0] * Tm[1][1] * dL_dc;
            gc[2] = 2 * Tm[0][0] * Tm[0][2] * dL_da + (Tm[0][0] * Tm[1][2] + Tm[0][2] * Tm[1][0]) * dL_db + 2 * Tm[1][0] * Tm[1][2] * dL_dc;
            gc[4] = 2 * Tm[0][2] * Tm[0][1] * dL_da + (Tm[0][1] * Tm[1][2] + Tm[0][2] * Tm[1][1]) * dL_db + 2 * Tm[1][1] * Tm[1][2] * dL_dc;
        }
#define TV(r, k) (Tm[r][0] * V3[k][0] + Tm[r][1] * V3[k][1] + Tm[r][2] * V3[k][2])
        const float dL_dT00 = 2 * TV(0, 0) * dL_da + TV(1, 0) * dL_db;
        const float dL_dT01 = 2 * TV(0, 1) * dL_da + TV(1, 1) * dL_db;
        const float dL_dT02 = 2 * TV(0, 2) * dL_da + TV(1, 2) * dL_db;
        const float dL_dT10 = 2 * TV(1, 0) * dL_dc + TV(0, 0) * dL_db;
        const float dL_dT11 = 2 * TV(1, 1) * dL_dc + TV(0, 1) * dL_db;
        const float dL_dT12 = 2 * TV(1, 2) * dL_dc + TV(0, 2) * dL_db;
#undef TV
        /* W[k][r] = view[4r+k] */
        const float dL_dJ00 = view[0] * dL_dT00 + view[4] * dL_dT01 + view[8] * dL_dT02;
        const float dL_dJ02 = view[2] * dL_dT00 + view[6] * dL_dT01 + view[10] * dL_dT02;
        const float dL_dJ11 = view[1] * dL_dT10 + view[5] * dL_dT11 + view[9] * dL_dT12;
        const float dL_dJ12 = view[2] * dL_dT10 + view[6] * dL_dT11 + view[10] * dL_dT12;
        const float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        const float dL_dtx = x_grad_mul * -focal_x * tz2 * dL_dJ02;
        const float dL_dty = y_grad_mul * -focal_y * tz2 * dL_dJ12;
        const float dL_dtz = -focal_x * tz2 * dL_dJ00 - focal_y * tz2 * dL_dJ11 + (2 * focal_x * t[0]) * tz3 * dL_dJ02 +
                             (2 * focal_y * t[1]) * tz3 * dL_dJ12;
        float dmean[3];
        dmean[0] = view[0] * dL_dtx + view[1] * dL_dty + view[2] * dL_dtz;
        dmean[1] = view[4] * dL_dtx + view[5] * dL_dty + view[6] * dL_dtz;
        dmean[2] = view[8] * dL_dtx + view[9] * dL_dty + view[10] * dL_dtz;

        /* preprocessCUDA backward (backward_distwar.cu:371-388) */
        const float m_hw = proj[3] * mean[0] + proj[7] * mean[1] + proj[11] * mean[2] + proj[15];
        const float m_w = 1.0f / (m_hw + 0.0000001f);
        const float g2x = dL_dmean2D[3 * (size_t)idx], g2y = dL_dmean2D[3 * (size_t)idx + 1];
        const float mul1 = (proj[0] * mean[0] + proj[4] * mean[1] + proj[8] * mean[2] + proj[12]) * m_w * m_w;
        const float mul2 = (proj[1] * mean[0] + proj[5] * mean[1] + proj[9] * mean[2] + proj[13]) * m_w * m_w;
        dmean[0] += (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
        dmean[1] += (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
        dmean[2] += (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;

        /* SH backward (backward_distwar.cu:21-140) */
        if (shs && M > 0) {
            const float* sh = shs + (size_t)idx * M * 3;
            float* dsh = dL_dsh + (size_t)idx * M * 3;
            const float dox = mean[0] - campos[0], doy = mean[1] - campos[1], doz = mean[2] - campos[2];
            const float len = sqrtf(dox * dox + doy * doy + doz * doz);
            const float x = dox / len, y = doy / len, z = doz / len;
            float dRGB[3], ddx[3] = {0, 0, 0}, ddy[3] = {0, 0, 0}, ddz[3] = {0, 0, 0};
            for (int ch = 0; ch < 3; ++ch) dRGB[ch] = dL_dcolor[(size_t)idx * C + ch] * (s->clamped[3 * (size_t)idx + ch] ? 0.f : 1.f);
#define SH(k) sh[(k) * 3 + ch]
#define DSH(k) dsh[(k) * 3 + ch]
            for (int ch = 0; ch < 3; ++ch) {
                DSH(0) = SH_C0 * dRGB[ch];
                if (D > 0) {
                    DSH(1) = -SH_C1 * y * dRGB[ch]; DSH(2) = SH_C1 * z * dRGB[ch]; DSH(3) = -SH_C1 * x * dRGB[ch];
                    ddx[ch] = -SH_C1 * SH(3); ddy[ch] = -SH_C1 * SH(1); ddz[ch] = SH_C1 * SH(2);
                    if (D > 1) {
                        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                        DSH(4) = SH_C2[0] * xy * dRGB[ch]; DSH(5) = SH_C2[1] * yz * dRGB[ch];
                        DSH(6) = SH_C2[2] * (2.f * zz - xx - yy) * dRGB[ch]; DSH(7) = SH_C2[3] * xz * dRGB[ch];
                        DSH(8) = SH_C2[4] * (xx - yy) * dRGB[ch];
                        ddx[ch] += SH_C2[0] * y * SH(4) + SH_C2[2] * 2.f * -x * SH(6) + SH_C2[3] * z * SH(7) + SH_C2[4] * 2.f * x * SH(8);
                        ddy[ch] += SH_C2[0] * x * SH(4) + SH_C2[1] * z * SH(5) + SH_C2[2] * 2.f * -y * SH(6) + SH_C2[4] * 2.f * -y * SH(8);
                        ddz[ch] += SH_C2[1] * y * SH(5) + SH_C2[2] * 2.f * 2.f * z * SH(6) + SH_C2[3] * x * SH(7);
                        if (D > 2) {
                            DSH(9) = SH_C3[0] * y * (3.f * xx - yy) * dRGB[ch]; DSH(10) = SH_C3[1] * xy * z * dRGB[ch];
                            DSH(11) = SH_C3[2] * y * (4.f * zz - xx - yy) * dRGB[ch];
                            DSH(12) = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy) * dRGB[ch];
                            DSH(13) = SH_C3[4] * x * (4.f * zz - xx - yy) * dRGB[ch];
                            DSH(14) = SH_C3[5] * z * (xx - yy) * dRGB[ch]; DSH(15) = SH_C3[6] * x * (xx - 3.f * yy) * dRGB[ch];
                            ddx[ch] += (SH_C3[0] * SH(9) * 3.f * 2.f * xy + SH_C3[1] * SH(10) * yz + SH_C3[2] * SH(11) * -2.f * xy +
                                        SH_C3[3] * SH(12) * -3.f * 2.f * xz + SH_C3[4] * SH(13) * (-3.f * xx + 4.f * zz - yy) +
                                        SH_C3[5] * SH(14) * 2.f * xz + SH_C3[6] * SH(15) * 3.f * (xx - yy));
                            ddy[ch] += (SH_C3[0] * SH(9) * 3.f * (xx - yy) + SH_C3[1] * SH(10) * xz +
                                        SH_C3[2] * SH(11) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * SH(12) * -3.f * 2.f * yz +
                                        SH_C3[4] * SH(13) * -2.f * xy + SH_C3[5] * SH(14) * -2.f * yz + SH_C3[6] * SH(15) * -3.f * 2.f * xy);
                            ddz[ch] += (SH_C3[1] * SH(10) * xy + SH_C3[2] * SH(11) * 4.f * 2.f * yz +
                                        SH_C3[3] * SH(12) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * SH(13) * 4.f * 2.f * xz +
                                        SH_C3[5] * SH(14) * (xx - yy));
                        }
                    }
                }
            }
#undef SH
#undef DSH
            const float ddirx = ddx[0] * dRGB[0] + ddx[1] * dRGB[1] + ddx[2] * dRGB[2];
            const float ddiry = ddy[0] * dRGB[0] + ddy[1] * dRGB[1] + ddy[2] * dRGB[2];
            const float ddirz = ddz[0] * dRGB[0] + ddz[1] * dRGB[1] + ddz[2] * dRGB[2];
            /* dnormvdv (auxiliary.h:107-117) */
            const float sum2 = dox * dox + doy * doy + doz * doz;
            const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
            dmean[0] += ((+sum2 - dox * dox) * ddirx - doy * dox * ddiry - doz * dox * ddirz) * invsum32;
            dmean[1] += (-dox * doy * ddirx + (sum2 - doy * doy) * ddiry - doz * doy * ddirz) * invsum32;
            dmean[2] += (-dox * doz * ddirx - doy * doz * ddiry + (sum2 - doz * doz) * ddirz) * invsum32;
        }
        gm[0] = dmean[0]; gm[1] = dmean[1]; gm[2] = dmean[2];

        /* computeCov3D backward (backward_distwar.cu:279-342) */
        if (scales) {
            const float* q4 = rotations + 4 * (size_t)idx;
            const float r = q4[0], x = q4[1], y = q4[2], z = q4[3];
            float R[3][3];
            R[0][0] = 1.f - 2.f * (y * y + z * z); R[0][1] = 2.f * (x * y - r * z); R[0][2] = 2.f * (x * z + r * y);
            R[1][0] = 2.f * (x * y + r * z); R[1][1] = 1.f - 2.f * (x * x + z * z); R[1][2] = 2.f * (y * z - r * x);
            R[2][0] = 2.f * (x * z - r * y); R[2][1] = 2.f * (y * z + r * x); R[2][2] = 1.f - 2.f * (x * x + y * y);
            const float sv[3] = {prm->scale_modifier * scales[3 * (size_t)idx], prm->scale_modifier * scales[3 * (size_t)idx + 1],
                                 prm->scale_modifier * scales[3 * (size_t)idx + 2]};
            float Mx[3][3], dSig[3][3], dM[3][3], dMt[3][3];
            for (int cc = 0; cc < 3; ++cc) for (int rr = 0; rr < 3; ++rr) Mx[cc][rr] = sv[rr] * R[cc][rr];
            dSig[0][0] = gc[0]; dSig[0][1] = 0.5f * gc[1]; dSig[0][2] = 0.5f * gc[2];
            dSig[1][0] = 0.5f * gc[1]; dSig[1][1] = gc[3]; dSig[1][2] = 0.5f * gc[4];
            dSig[2][0] = 0.5f * gc[2]; dSig[2][1] = 0.5f * gc[4]; dSig[2][2] = gc[5];
            for (int cc = 0; cc < 3; ++cc)
                for (int rr = 0; rr < 3; ++rr)
                    dM[cc][rr] = 2.0f * Mx[0][rr] * dSig[cc][0] + 2.0f * Mx[1][rr] * dSig[cc][1] + 2.0f * Mx[2][rr] * dSig[cc][2];
            for (int cc = 0; cc < 3; ++cc) for (int rr = 0; rr < 3; ++rr) dMt[cc][rr] = dM[rr][cc];
            /* Rt[c] . dMt[c] with Rt[c][k] = R[k][c] */
            gs[0] = R[0][0] * dMt[0][0] + R[1][0] * dMt[0][1] + R[2][0] * dMt[0][2];
            gs[1] = R[0][1] * dMt[1][0] + R[1][1] * dMt[1][1] + R[2][1] * dMt[1][2];
            gs[2] = R[0][2] * dMt[2][0] + R[1][2] * dMt[2][1] + R[2][2] * dMt[2][2];
            for (int k = 0; k < 3; ++k) { dMt[0][k] *= sv[0]; dMt[1][k] *= sv[1]; dMt[2][k] *= sv[2]; }
            gq[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
            gq[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
            gq[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
            gq[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
        }
    }
}

/* rasterizer_impl.cu:54-66 */
void orc_mark_visible(int P, const float* means3D, const float* view, uint8_t* present) {
    for (int i = 0; i < P; ++i) {
        const float* p = means3D + 3 * (size_t)i;
        present[i] = !(row_affine(view, 2, p[0], p[1], p[2]) <= 0.2f);
    }
}

/* exact 3-NN mean squared distance by brute force (simple_knn.cu:133-184 semantics) */
void orc_knn3(int P, const float* pts, float* out) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
        const float* a = pts + 3 * (size_t)i;
        for (int j = 0; j < P; ++j) {
            if (j == i) continue;
            const float* b = pts + 3 * (size_t)j;
            const float dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
            float dist = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
            for (int k = 0; k < 3; ++k)
                if (best[k] > dist) { float t = best[k]; best[k] = dist; dist = t; }
        }
        out[i] = (best[0] + best[1] + best[2]) / 3.0f;
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
