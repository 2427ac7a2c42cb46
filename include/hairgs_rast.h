/*
 * hairgs_rast.h — C ABI of the B200-native differentiable Gaussian rasterizer that drops in
 * behind Hair-GS's render path (libhairgs_rast.so, sm_100a only, no CPU fallback).
 *
 * Every entry point takes plain pointers / sizes / a cudaStream_t (as void*), returns an int
 * status (>= 0 ok, < 0 error; text via hgs_last_error()).  No torch types cross this boundary.
 * All pointers are DEVICE pointers unless the name ends in _host.  "Absent" optional inputs are
 * passed as NULL exactly like the reference (rasterize_points.cu:84-87, forward.cu:205,241).
 *
 * Reference interface each entry point replaces (paths relative to
 * submodules/diff-gaussian-rasterization/ and submodules/simple-knn/ of yimin-pan/hair-gs):
 *
 *   hgs_rasterize_forward   <- CudaRasterizer::Rasterizer::forward   cuda_rasterizer/rasterizer.h:33-58
 *                              (impl rasterizer_impl.cu:198-336), bound by RasterizeGaussiansCUDA
 *                              rasterize_points.cu:35-115
 *   hgs_rasterize_backward  <- CudaRasterizer::Rasterizer::backward  cuda_rasterizer/rasterizer.h:60-85
 *                              (impl rasterizer_impl.cu:340-434), bound by
 *                              RasterizeGaussiansBackwardCUDA rasterize_points.cu:117-196
 *   hgs_mark_visible        <- CudaRasterizer::Rasterizer::markVisible cuda_rasterizer/rasterizer.h:24-31
 *                              (impl rasterizer_impl.cu:141-153), bound by markVisible
 *                              rasterize_points.cu:198-217
 *   hgs_dist2_knn3          <- SimpleKNN::knn simple_knn.h:17 (impl simple_knn.cu:186-222), bound by
 *                              distCUDA2 spatial.cu:15-26
 *   hgs_*_bytes / hgs_forward_stage_* / hgs_state_view
 *                           <- GeometryState/ImageState/BinningState::fromChunk + required<T>()
 *                              rasterizer_impl.cu:155-194, rasterizer_impl.h:22-71 (internal ABI of
 *                              the reference; ours is laid out differently)
 *
 * Entry points for the callers either side of the rasterizer (paths relative to the
 * Hair-GS repository root).  The reference implements these in Python/torch, so there is no C
 * interface to cite; each replaces the torch code at the lines given:
 *
 *   hgs_strands_forward_stage_a/_b, hgs_strands_backward
 *                           <- HairGaussianModel getters scene/hair_gaussian_model.py:134-206 +
 *                              the three render() calls of one iteration loss/losses.py:245-248,
 *                              311-312, train.py:146-155
 *   hgs_weighted_l1         <- l1_loss loss/losses.py:16-17
 *   hgs_hair_image_loss     <- loss_function's image terms loss/losses.py:319-346 (ssim :43-84,
 *                              orientation_loss_rast :224-289, mask_loss_rast :292-316)
 *   hgs_adam_step           <- torch.optim.Adam as configured at scene/gaussian_model.py:250,
 *                              stepped at train.py:203-204
 *   hgs_densify_stats       <- update_densification_stats scene/gaussian_model.py:675-682
 *   hgs_merge_count/_fill/_greedy
 *                           <- compute_endpoint_pair_to_merge scene/hair_gaussian_model.py:1205-1362
 *                              (scipy.spatial.cKDTree.query_ball_point + Python loops)
 */
#ifndef HAIRGS_RAST_H_
#define HAIRGS_RAST_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HGS_ABI_VERSION 4
#define HGS_TILE 16            /* BLOCK_X == BLOCK_Y == 16, cuda_rasterizer/config.h:16-17 */
#define HGS_MAX_CHANNELS 8     /* colour channels per pass: 3 (reference) .. 8 (fused RGB+mask+orientation) */

/* status codes */
#define HGS_OK 0
#define HGS_ERR_INVALID (-1)     /* bad argument (mirrors AT_ERROR / std::runtime_error of the reference) */
#define HGS_ERR_CUDA (-2)        /* a CUDA call or (debug mode) a kernel failed */
#define HGS_ERR_ALLOC (-3)       /* an allocator callback returned NULL / too small */
#define HGS_ERR_OVERFLOW (-4)    /* instance count exceeds 2^31-1 (reference: int num_rendered) */

/* Allocator callback: the C form of the reference's std::function<char*(size_t)> resize functors
 * (rasterize_points.cu:27-33).  Must return a device pointer to >= bytes bytes, 256-B aligned. */
typedef void* (*hgs_alloc_fn)(void* user, size_t bytes);

/* Scalar parameters of one rasterization call (rasterizer.h:33-58 scalar arguments). */
typedef struct hgs_raster_params {
    int32_t P;              /* number of Gaussians */
    int32_t D;              /* active SH degree 0..3 */
    int32_t M;              /* SH coefficients per channel in the shs tensor, 0 if absent */
    int32_t width, height;  /* image size in pixels */
    int32_t channels;       /* colour channels; 3 unless colors_precomp carries more (<= HGS_MAX_CHANNELS) */
    float tan_fovx, tan_fovy;
    float scale_modifier;
    int32_t prefiltered;    /* reference traps when a prefiltered point is culled; we report HGS_ERR_INVALID lazily via debug */
    int32_t debug;          /* synchronise + check after every stage (auxiliary.h:166-173) */
    int32_t sort_depth_bits;/* 0 or 32: sort on the full 32 depth bits (reference key).  1..31: the caller asserts that the
                             * bit patterns of all visible depths span less than 2^sort_depth_bits; the sort then drops
                             * the unused high bits (fewer passes).  The range is known after stage A
                             * (hgs_forward_read_num_rendered words 3-4): a caller that enqueued stage B on a hint must
                             * check (depth_max - depth_min) >> sort_depth_bits == 0 and otherwise repeat stage B with 0
                             * (the device also raises bit 1 of the overflow word).  Only read by the stage-B entries. */
    int32_t sort_mode;      /* HGS_SORT_TILE (0): instances are partitioned by (tile, depth slice) first (counts taken
                             * by the preprocess, one scatter) and every list is sorted by (depth, id) inside shared memory by
                             * the kernel that also packs the sorted records - two passes over the instances.  A list longer
                             * than HGS_TILE_SORT_MAX raises bit 2 of the overflow word in stage A: run stage B with HGS_SORT_GLOBAL.
                             * Read by stage A too (it takes the counts): pass the same value to both stages; a stage B in
                             * HGS_SORT_TILE after a stage A in HGS_SORT_GLOBAL is invalid (the reverse is fine).
                             * HGS_SORT_GLOBAL (1): stable radix sort of the 64-bit (tile | depth) keys, the reference's
                             * formulation (rasterizer_impl.cu:300-308); sort_depth_bits applies to this mode only.  This is
                             * what the Python packages select by default (measured faster on B200, profiles/r2_tilesort.md).
                             * Both produce identical keys, point list, ranges and records. */
    int32_t slice_base;     /* HGS_SORT_TILE only, performance hints (any values are CORRECT): every tile list is split into */
    int32_t slice_shift;    /* HGS_TILE_SLICES depth slices, slice = clamp((depth_bits - slice_base) >> slice_shift, 0, S-1),
                             * each sorted on its own; depth order between slices is implied.  Good values spread the visible
                             * depths over the slices: slice_base = bit pattern of the smallest visible depth, slice_shift =
                             * bits(depth_max - depth_min) - log2(S), both known from an earlier view's read-back words 3-4.
                             * slice_shift >= 32 (or 0 / 0 on a first call) puts everything into slice 0. */
} hgs_raster_params;
#define HGS_SORT_TILE 0
#define HGS_SORT_GLOBAL 1
#define HGS_TILE_SLICES 16       /* depth slices per tile list (HGS_SORT_TILE) */
#define HGS_TILE_SORT_MAX 16384  /* longest (tile, slice) list the in-tile sort takes */

/* Device pointers of the per-Gaussian inputs (rasterizer.h:33-58 pointer arguments). */
typedef struct hgs_raster_inputs {
    const float* background;     /* [channels] */
    const float* means3D;        /* [P,3] */
    const float* shs;            /* [P,M,3] or NULL */
    const float* colors_precomp; /* [P,channels] or NULL */
    const float* opacities;      /* [P] */
    const float* scales;         /* [P,3] or NULL */
    const float* rotations;      /* [P,4] (w,x,y,z) or NULL */
    const float* cov3D_precomp;  /* [P,6] or NULL */
    const float* viewmatrix;     /* [16] column-major W2C */
    const float* projmatrix;     /* [16] column-major P*W2C */
    const float* cam_pos;        /* [3] */
} hgs_raster_inputs;

/* Gradient outputs (rasterizer.h:60-85).  Every element of every non-NULL array is written
 * (zeros for culled Gaussians): callers need NOT pre-zero anything. */
typedef struct hgs_raster_grads {
    float* dL_dmean2D;   /* [P,3]  (z always 0) */
    float* dL_dconic;    /* [P,4]  scratch: (xx, xy, unused, yy) as the reference's [P,2,2] */
    float* dL_dopacity;  /* [P] */
    float* dL_dcolor;    /* [P,channels] */
    float* dL_dmean3D;   /* [P,3] */
    float* dL_dcov3D;    /* [P,6] */
    float* dL_dsh;       /* [P,M,3] or NULL when M == 0 */
    float* dL_dscale;    /* [P,3] */
    float* dL_drot;      /* [P,4] */
} hgs_raster_grads;

/* ---- Fused strand-aligned entry (SURVEY.md §8f rows N1 + N2; no single reference counterpart) ----------------
 * One pass renders what Hair-GS obtains from three render() calls per training view (train.py:146-155,
 * loss/losses.py:246-249,311-312,341-346): channels 0-2 SH colour, 3 mask = sigmoid(mask_logit), 4-6 the world-space
 * unit direction of the segment.  The Gaussians are derived inside the preprocess kernel from the segment end points
 * exactly as HairGaussianModel's getters do (scene/hair_gaussian_model.py:134-206): mean = (e0+e1)/2,
 * sigma_x = max(|e1-e0|/2 * 0.51021, 1e-7), sigma_yz = exp(width), x axis rotated onto the segment,
 * opacity = sigmoid(opacity_logit); the backward returns gradients w.r.t. the RAW parameters. prm->channels must be 7. */
typedef struct hgs_strand_inputs {
    const float* background;       /* [7] */
    const float* endpoints;        /* [E,3] joints */
    const int64_t* endpoint_pairs; /* [P,2] */
    const float* width;            /* [P] log sigma_yz */
    const float* opacity_logit;    /* [P] */
    const float* mask_logit;       /* [P] */
    const float* features;         /* [P,M,3] SH coefficients (dc first) */
    const float* viewmatrix;       /* [16] */
    const float* projmatrix;       /* [16] */
    const float* cam_pos;          /* [3] */
    int64_t num_endpoints;         /* E */
} hgs_strand_inputs;

typedef struct hgs_strand_grads {
    float* dL_dmean2D;        /* [P,3] screen-space mean gradients (densification statistics), z = 0 */
    float* dL_dconic;         /* [P,4] scratch */
    float* dL_dopacity;       /* [P]   scratch (w.r.t. activated opacity) */
    float* dL_dcolor;         /* [P,7] scratch */
    float* dL_dendpoints;     /* [E,3] */
    float* dL_dwidth;         /* [P] */
    float* dL_dopacity_logit; /* [P] */
    float* dL_dmask_logit;    /* [P] */
    float* dL_dfeatures;      /* [P,M,3] */
    int32_t accumulate;       /* 0: the five parameter gradients above are overwritten (every element written once);
                               * 1: they are ADDED to (several views sunk into one gradient bucket) */
    float* acc16;             /* ABI v4, may be NULL.  [P,16] scratch, 16-byte aligned: ONE interleaved accumulation record per
                               * Gaussian (mean2D.xy, opacity, - | conic xx, xy, -, yy | colour 0-3 | colour 4-6, -) that the
                               * backward compositor fills with four red.global.add.v4.f32 per (instance, pixel block) instead
                               * of thirteen scalar atomics; dL_dconic / dL_dopacity / dL_dcolor are then unused (may be NULL)
                               * and dL_dmean2D is written from the record by the preprocess backward.  Vector reductions
                               * flush denormal addends to zero (REDG.F32x4.FTZ). */
} hgs_strand_grads;

int hgs_strands_forward_stage_a(const hgs_raster_params* prm, const hgs_strand_inputs* in, void* geom_ws,
                                int32_t* radii, void* stream);
int hgs_strands_forward_stage_b(const hgs_raster_params* prm, const hgs_strand_inputs* in, void* geom_ws,
                                void* binning_ws, void* image_ws, int64_t capacity, float* out_color, void* stream);
int hgs_strands_backward(const hgs_raster_params* prm, const hgs_strand_inputs* in, int64_t capacity,
                         const void* geom_ws, const void* binning_ws, const void* image_ws, const float* dL_dpix,
                         const hgs_strand_grads* grads, void* stream);

/* The backward in its two halves (ABI v4): parts & 1 = clear the per-Gaussian accumulators + backward compositor (touches
 * only this view's scratch), parts & 2 = preprocess backward (writes / adds the PARAMETER gradients).  Views of one
 * optimiser step that run on different streams may overlap their part 1; their part 2 must be ordered (it read-modify-
 * writes the shared gradient tensors).  hgs_strands_backward == parts 3. */
int hgs_strands_backward_parts(const hgs_raster_params* prm, const hgs_strand_inputs* in, int64_t capacity,
                               const void* geom_ws, const void* binning_ws, const void* image_ws, const float* dL_dpix,
                               const hgs_strand_grads* grads, int32_t parts, void* stream);

int hgs_abi_version(void);
const char* hgs_last_error(void);

/* Workspace sizes in bytes (required<T>() of the reference).  binning: N = num_rendered. */
size_t hgs_geom_bytes(int32_t P, int32_t channels, int32_t width, int32_t height);  /* ABI v4: + per-tile counters */
size_t hgs_image_bytes(int32_t width, int32_t height);
size_t hgs_binning_bytes(int64_t num_rendered, int32_t channels);
/* capacity a binning workspace of `bytes` bytes was sized for: num_rendered itself if it matches, else the
 * multiple of 4096 that does (sync-free mode); < 0 if none. */
int64_t hgs_binning_capacity(size_t bytes, int32_t channels, int64_t num_rendered);

/* Whole forward pass, reference call shape: allocates through the three callbacks, blocks once on
 * the instance count exactly like rasterizer_impl.cu:281, returns num_rendered (>= 0) or an error.
 * out_color [channels,H,W] and radii [P] (may be NULL) are fully written. */
int hgs_rasterize_forward(hgs_alloc_fn geom_alloc, void* geom_user,
                          hgs_alloc_fn binning_alloc, void* binning_user,
                          hgs_alloc_fn image_alloc, void* image_user,
                          const hgs_raster_params* prm, const hgs_raster_inputs* in,
                          float* out_color, int32_t* radii, void* stream);

/* The same pass in three stream-ordered stages, for callers that own their workspaces and want to
 * overlap the instance-count read-back (no allocation, no implicit synchronisation):
 *   A  preprocess + tile-count scan; leaves num_rendered in the geometry workspace
 *   (read it with hgs_forward_read_num_rendered: async copy of FIVE uint32 into pinned host memory:
 *    num_rendered, -, overflow flags, depth_max, ~depth_min — the last two are the bit patterns bounding the depths of
 *    the visible Gaussians, see hgs_raster_params.sort_depth_bits)
 *   B  key emission + (tile|depth) radix sort + tile ranges + compositing.  `capacity` is the number of
 *      instances binning_ws was sized for (hgs_binning_bytes(capacity, channels)); the live count is read from
 *      the geometry workspace ON THE DEVICE, so B may be enqueued before the host knows it.  If the count
 *      turns out larger than capacity the outputs are garbage-but-in-bounds: call B again with more room. */
int hgs_forward_stage_a(const hgs_raster_params* prm, const hgs_raster_inputs* in,
                        void* geom_ws, int32_t* radii, void* stream);
int hgs_forward_read_num_rendered(const void* geom_ws, int32_t P, uint32_t* n_pinned_host, void* stream);
int hgs_forward_stage_b(const hgs_raster_params* prm, const hgs_raster_inputs* in,
                        void* geom_ws, void* binning_ws, void* image_ws, int64_t capacity,
                        const int32_t* radii, float* out_color, void* stream);

/* Stage B in its two halves (ABI v4), for callers that overlap the latency-bound binning of one view with the issue-bound
 * compositing of another on a second stream (hairgs_b200/graphs.py: GraphedStrandBatch; the reference runs K4-K7 back to
 * back on one stream, rasterizer_impl.cu:284-333):
 *   binning    key emission + (tile|depth) sort + tile ranges + sorted-order record packing   (K4-K6)
 *   composite  per-tile alpha compositing into out_color / final_T / n_contrib                (K7)
 * Same workspaces and `capacity` as hgs_forward_stage_b, which is exactly binning followed by composite.  Valid for both
 * entries (Gaussians given, or strand end points): only prm and the background are read. */
int hgs_forward_stage_b_binning(const hgs_raster_params* prm, void* geom_ws, void* binning_ws, void* image_ws,
                                int64_t capacity, void* stream);
int hgs_forward_stage_b_composite(const hgs_raster_params* prm, const float* background, void* geom_ws, void* binning_ws,
                                  void* image_ws, int64_t capacity, float* out_color, void* stream);

/* CUDA-graph helpers (ABI v4).  A captured graph runs all of its kernel nodes at the priority of the stream it is launched
 * into unless it is instantiated with cudaGraphInstantiateFlagUseNodePriority; the two-branch batch graph needs the node
 * priorities (its binning branch is captured on a high-priority stream so that its short, latency-bound kernels are
 * dispatched ahead of the thousands of queued compositor blocks of the other branch).  `graph` is a cudaGraph_t obtained
 * from the capture (e.g. torch.cuda.CUDAGraph(keep_graph=True).raw_cuda_graph()); *exec receives a cudaGraphExec_t. */
int hgs_graph_instantiate(void* graph, int32_t use_node_priority, void** exec);
int hgs_graph_launch(void* exec, void* stream);
int hgs_graph_exec_destroy(void* exec);

/* Backward pass.  R = the capacity binning_ws was carved with (num_rendered in exact mode, see
 * hgs_binning_capacity); workspaces are the forward's. */
int hgs_rasterize_backward(const hgs_raster_params* prm, const hgs_raster_inputs* in,
                           int64_t R, const int32_t* radii,
                           const void* geom_ws, const void* binning_ws, const void* image_ws,
                           const float* dL_dpix, const hgs_raster_grads* grads, void* stream);

/* Frustum test only (checkFrustum, rasterizer_impl.cu:54-66): present[i] = view-space z > 0.2. */
int hgs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, uint8_t* present, void* stream);

/* Weighted L1 over a stack of image planes, value and gradient in one pass (first piece of SURVEY §8f N3):
 *   *loss = sum_c weights[c] * sum_i |image[c][i] - target[c][i]|,   dL_dimage[c][i] = weights[c] * sign(image - target).
 * With weights[c] = lambda_c / (planes_in_group * H*W) this is Hair-GS's l1_loss (loss/losses.py:16-17) summed over the
 * RGB / mask / orientation groups.  Planes must be 16-byte aligned and H*W a multiple of 4.  *loss is overwritten. */
int hgs_weighted_l1(int32_t C, int64_t HW, const float* image, const float* target, const float* weights, float* loss,
                    float* dL_dimage, void* stream);

/* Hair-GS's image-space loss of one training view, value and gradient w.r.t. the seven planes the
 * one-pass strand entry renders (rgb | mask logit | world orientation), replacing loss_function's image terms
 * (loss/losses.py:319-346):
 *   total = l_l1 * mean|rgb - gt_rgb|                                   (l1_loss, losses.py:16-17)
 *         + l_dssim * (1 - ssim(rgb, gt_rgb))                           (11x11 Gaussian window, sigma 1.5, losses.py:24-84)
 *         + l_mask * BCEWithLogits(mask, gt_mask)                       (mask_loss_rast, losses.py:292-316)
 *         + l_orient * mean_{orient_mask}(angle_diff(theta, gt_theta) * confidence)   (orientation_loss_rast, :224-289)
 * (the reference sets l_l1 = 1 - lambda_dssim).  theta = atan2 of the view-space xy of the orientation, folded to
 * [0, pi); angle_diff is the bidirectional distance pi/2 - ||theta - gt| - pi/2|.  orient_mask NULL selects the pixels
 * whose orientation differs from bg_orient (losses.py:266-268).
 * terms[8] (device): 0 total, 1 l1, 2 dssim, 3 mask, 4 orientation, 5 orientation pixel count, 6-7 unused.
 * scratch: 9*H*W floats.  dL_dimage: [7,H,W], every element written. */
typedef struct hgs_hair_loss {
    int32_t height, width;
    const float* image7;
    const float* gt_rgb;            /* [3,H,W] */
    const float* gt_mask;           /* [H,W] */
    const float* gt_theta;          /* [H,W] radians in [0, pi) */
    const float* confidence;        /* [H,W] */
    const uint8_t* orient_mask;     /* [H,W] bool, may be NULL */
    float view_rot[9];              /* world_view_transform[:3,:3], row-major (view = world @ view_rot) */
    float bg_orient[3];
    float l_l1, l_dssim, l_mask, l_orient;
    float* terms;
    float* scratch;
    float* dL_dimage;
    const float* view_matrix_dev;   /* ABI v3, may be NULL: DEVICE pointer to camera.world_view_transform ([4,4] floats in
                                     * torch's row-major layout, the tensor the rasterizer takes as `viewmatrix`); when set,
                                     * the kernels read the rotation from it and ignore view_rot, so that one captured CUDA
                                     * graph serves every view */
} hgs_hair_loss;
int hgs_hair_image_loss(const hgs_hair_loss* args, void* stream);

/* Target stacks in their storage format -> the float planes hgs_hair_image_loss reads (ABI v4).  What Hair-GS's cameras hold
 * per view (scene/cameras.py:60-85: original_image[3], float_mask, orientation_field, orientation_confidence) arrives from the
 * host as 8 bytes per pixel instead of 24: rgbm = [views, H*W] uchar4 (r, g, b in 0..255 as the image files store them, mask
 * 0 / 255), theta_conf = [views, H*W] half2 (orientation angle in [0, pi), confidence); out = [views, 6, H*W] float
 * (r/255, g/255, b/255, mask/255, theta, confidence).  One launch for all views of a step. */
int hgs_unpack_targets(int32_t views, int64_t HW, const void* rgbm, const void* theta_conf, float* out, void* stream);

/* Optimiser step over the flat parameter bucket: torch.optim.Adam as Hair-GS configures it
 * (scene/gaussian_model.py:250: per-group lr, betas (0.9, 0.999), eps 1e-15, no weight decay, no amsgrad) for every
 * parameter group in ONE launch.  Group g covers flat elements [group_end[g-1], group_end[g]) and steps with lr[g];
 * `step` is the 1-based iteration count (bias correction); the gradient is multiplied by grad_scale first (1/views for a
 * multi-view batch) and, if zero_grad != 0, cleared afterwards (optimizer.zero_grad, train.py:204).
 * All four buffers hold n floats and must be 16-byte aligned; group_end and lr are HOST arrays. */
#define HGS_ADAM_MAX_GROUPS 16
int hgs_adam_step(int64_t n, float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                  int32_t n_groups, const int64_t* group_end, const float* lr, int32_t step,
                  float beta1, float beta2, float eps, float grad_scale, int32_t zero_grad, void* stream);

/* update_densification_stats (scene/gaussian_model.py:675-682) for the Gaussians with radii > 0:
 *   max_radii2D = max(max_radii2D, radii);  xyz_gradient_accum += |dL_dmean2D.xy|;  denom += 1.
 * dL_dmean2D has grad_stride floats per Gaussian (3 for viewspace_points.grad). */
int hgs_densify_stats(int32_t P, const int32_t* radii, const float* dL_dmean2D, int32_t grad_stride,
                      float* max_radii2D, float* xyz_gradient_accum, float* denom, void* stream);

/* Strand endpoint merge search: HairGaussianModel.compute_endpoint_pair_to_merge,
 * scene/hair_gaussian_model.py:1205-1362, without the host cKDTree and the per-point Python loops.
 * The K strand ends considered are given by position, unit direction towards their neighbouring joint, endpoint id and the
 * endpoint id of the other end of their strand.  A pair (i, j) is a candidate when |p_i - p_j| <= radius (double
 * arithmetic like cKDTree), j != i, j is not the other end of i's strand and <dir_j, -dir_i> >= dir_th (absolute value if
 * bidirectional); per i at most max_num_nn candidates in ascending j are kept (<= 0: all).
 *   hgs_merge_count: counts[i] = number of candidates of end i.
 *   hgs_merge_fill:  writes end i's candidates at offsets[i] (exclusive scan of counts): p1 = id of i, p2 = id of j,
 *                    dist = float32 norm of the float32 difference (np.linalg.norm, :1321-1324).
 *   hgs_merge_greedy: pairs sorted by ascending dist -> keep[r] in {0,1} after remove_duplicate_endpoint_rows (:711-726)
 *                    and remove_complementary_rows (:1237-1255).  other_end_of[id] = other strand end of endpoint id or
 *                    -1; flags: one zeroed byte per endpoint id (scratch). */
int hgs_merge_count(int32_t K, const float* points, const float* dirs, const int32_t* global_id,
                    const int32_t* other_end, double radius, double dir_th, int32_t bidirectional,
                    int32_t max_num_nn, int32_t* counts, void* stream);
int hgs_merge_fill(int32_t K, const float* points, const float* dirs, const int32_t* global_id,
                   const int32_t* other_end, double radius, double dir_th, int32_t bidirectional,
                   int32_t max_num_nn, const int64_t* offsets, int32_t* p1, int32_t* p2, float* dist, void* stream);
int hgs_merge_greedy(int64_t n, const int32_t* p1, const int32_t* p2, const int32_t* other_end_of,
                     uint8_t* flags, uint8_t* keep, void* stream);

/* Mean squared distance to the 3 nearest neighbours of every point (distCUDA2).
 * workspace: hgs_knn_bytes(P) bytes of device scratch. */
size_t hgs_knn_bytes(int32_t P);
int hgs_dist2_knn3(int32_t P, const float* points, float* mean_dist2, void* workspace, void* stream);

/* Introspection of the opaque workspaces, for parity tests against the reference's
 * geomBuffer/binningBuffer/imgBuffer carve-up.  `what` selects a sub-array; the call copies it
 * (device to device, stream-ordered) into dst in the REFERENCE's element layout.
 * Returns the number of bytes written, or < 0. */
enum hgs_view {
    HGS_VIEW_DEPTHS = 0,        /* float[P] */
    HGS_VIEW_MEANS2D = 1,       /* float2[P] */
    HGS_VIEW_CONIC_OPACITY = 2, /* float4[P] */
    HGS_VIEW_RGB = 3,           /* float[P*channels] */
    HGS_VIEW_TILES_TOUCHED = 4, /* uint32[P] */
    HGS_VIEW_POINT_OFFSETS = 5, /* uint32[P] inclusive scan */
    HGS_VIEW_CLAMPED = 6,       /* uint8[P*3] (bool per channel) */
    HGS_VIEW_KEYS_SORTED = 7,   /* uint64[N] */
    HGS_VIEW_POINT_LIST = 8,    /* uint32[N] sorted Gaussian ids */
    HGS_VIEW_RANGES = 9,        /* uint2[tiles] */
    HGS_VIEW_FINAL_T = 10,      /* float[H*W] */
    HGS_VIEW_N_CONTRIB = 11,    /* uint32[H*W] */
    HGS_VIEW_KEYS_UNSORTED = 12,/* uint64[N]; only valid between stage emission and sort (debug builds keep a copy) */
    HGS_VIEW_COV3D = 13         /* float[P*6]; recomputed on demand */
};
int64_t hgs_state_view(int what, const hgs_raster_params* prm, const hgs_raster_inputs* in,
                       int64_t num_rendered, int64_t binning_capacity, const void* geom_ws, const void* binning_ws,
                       const void* image_ws, void* dst, void* stream);

/* Stand-alone stable LSD radix sort of (u64 key, u32 value) pairs on bits [0, end_bit): the
 * hand-written onesweep that replaces cub::DeviceRadixSort::SortPairs (rasterizer_impl.cu:303-308).
 * Exposed for parity tests and ncu captures.  workspace: hgs_sort_bytes(n). Sorted output is left in
 * keys_out/vals_out (both in/out pairs are clobbered). */
size_t hgs_sort_bytes(int64_t n);
int hgs_sort_pairs(int64_t n, int end_bit, uint64_t* keys_in, uint32_t* vals_in,
                   uint64_t* keys_out, uint32_t* vals_out, void* workspace, void* stream);

/* Per-stage device timing and launch counting (no reference counterpart; the reference only has the
 * debug-mode synchronisation of auxiliary.h:166-173).  hgs_profile_enable(1) makes every kernel launch of the
 * library bracket itself with CUDA events on its stream; hgs_profile_collect synchronises the device, adds
 * the elapsed milliseconds per stage into ms[HGS_STAGE_COUNT], the launches per stage into
 * launches[HGS_STAGE_COUNT], and resets.  Launch counts are kept even when timing is disabled. */
enum hgs_stage {
    HGS_STAGE_PREPROCESS_FWD = 0,
    HGS_STAGE_EMIT_KEYS = 1,
    HGS_STAGE_SORT_HISTOGRAM = 2,
    HGS_STAGE_SORT_ONESWEEP = 3,
    HGS_STAGE_TILE_RANGES = 4,
    HGS_STAGE_COMPOSITE_FWD = 5,
    HGS_STAGE_COMPOSITE_BWD = 6,
    HGS_STAGE_PREPROCESS_BWD = 7,
    HGS_STAGE_KNN = 8,
    HGS_STAGE_OTHER = 9,
    HGS_STAGE_TILE_SCAN = 10,
    HGS_STAGE_TILE_COUNT = 11,      /* (unused: the per-tile counts are taken by preprocess_fwd) */
    HGS_STAGE_TILE_OFFSETS = 12,    /* tile ranges + processing order */
    HGS_STAGE_TILE_SCATTER = 13,    /* instances to their tile's range */
    HGS_STAGE_TILE_SORT_PACK = 14,  /* per-tile (depth, id) sort + sorted-order record packing (3 size classes) */
    HGS_STAGE_COUNT = 15
};
/* Debug: when dev_ptr != NULL the forward compositor writes one uint4 per (tile, warp):
 * (chunks walked, cull candidates, blends summed over lanes, pixels terminated | list chunks << 8). */
int hgs_debug_set_stats(void* dev_ptr);
/* Debug / measurement: pixel-block shape of the compositors (no reference counterpart: renderCUDA, forward.cu:261, and
 * renderCUDABW_*, backward_distwar.cu:450-1014, have one thread per pixel of a 16x16 tile).  0 = default (environment
 * variable HGS_COMPOSITE_BLOCKS=4x4|8x4, else 4x4), 1 = two 4x4 blocks per warp, 2 = one 8x4 block per warp.
 * Outputs are identical in both modes. */
int hgs_debug_set_composite_blocks(int mode);
int hgs_profile_enable(int on);
int hgs_profile_collect(double* ms, int64_t* launches);
const char* hgs_stage_name(int stage);

#ifdef __cplusplus
}
#endif
#endif /* HAIRGS_RAST_H_ */
