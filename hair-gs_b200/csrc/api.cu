// api.cu — the extern "C" surface declared in include/hairgs_rast.h: argument checks, workspace
// carving, stage orchestration.  Mirrors the orchestration (not the code) of
// CudaRasterizer::Rasterizer::{forward,backward,markVisible} rasterizer_impl.cu:141-434.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>
#include "hgs_common.cuh"

namespace hgs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return HGS_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return HGS_ERR_CUDA;
}

// debug mode: synchronise and surface asynchronous kernel faults per stage (auxiliary.h:166-173)
int stage_check(const char* stage, int debug, cudaStream_t s) {
    if (!debug) return HGS_OK;
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("[CUDA ERROR] in stage %s: %s", stage, cudaGetErrorString(e));
        return HGS_ERR_CUDA;
    }
    return HGS_OK;
}

// ---- stage profiler -------------------------------------------------------------------------------
// process-wide (a profiler is), guarded: launches may come from several host threads (one per device / stream)
static std::atomic<bool> g_prof_on{false};
static std::atomic<long long> g_launches[HGS_STAGE_COUNT];
struct ProfRec { int stage; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mutex;

StageScope::StageScope(int stage_id, cudaStream_t s) : stage(stage_id), stream(s), stop(nullptr) {
    g_launches[stage].fetch_add(1, std::memory_order_relaxed);
    if (g_prof_on.load(std::memory_order_relaxed)) {
        ProfRec r;
        r.stage = stage;
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        cudaEventRecord(r.a, stream);
        stop = r.b;
        std::lock_guard<std::mutex> lock(g_prof_mutex);
        g_prof.push_back(r);
    }
}
StageScope::~StageScope() {
    if (stop) cudaEventRecord(stop, stream);
}

// launchers implemented in the other translation units
int launch_preprocess_fwd(const hgs_raster_params*, const hgs_raster_inputs*, const GeomLayout&, int32_t*, cudaStream_t);
int launch_preprocess_bwd(const hgs_raster_params*, const hgs_raster_inputs*, const GeomLayout&, const int32_t*,
                          const hgs_raster_grads*, cudaStream_t);
int launch_mark_visible(int, const float*, const float*, uint8_t*, cudaStream_t);
int launch_strand_preprocess_fwd(const hgs_raster_params*, const hgs_strand_inputs*, const GeomLayout&, int32_t*, cudaStream_t);
int launch_strand_preprocess_bwd(const hgs_raster_params*, const hgs_strand_inputs*, const GeomLayout&,
                                 const hgs_strand_grads*, cudaStream_t);
int launch_view_geom(int, const hgs_raster_params*, const hgs_raster_inputs*, const GeomLayout&, void*, cudaStream_t);
int launch_emit_keys(int, const GeomLayout&, const uint2*, uint64_t*, uint32_t*, uint32_t, uint32_t, cudaStream_t);
int launch_sort_pairs(int64_t, const uint32_t*, int, uint64_t* [2], uint32_t* [2], void*, int*, cudaStream_t, GeomHeader*, int, int);
int launch_finalize_sorted(int, int64_t, const uint32_t*, const uint64_t*, const uint32_t*, const GeomLayout&,
                           const BinningLayout&, uint2*,
                           uint32_t*, size_t, uint32_t, cudaStream_t);
int launch_composite_fwd(int, const ImageLayout&, const BinningLayout&, int, int, const float*, float*, cudaStream_t);
int launch_composite_bwd(int, const ImageLayout&, const BinningLayout&, const uint32_t*, int, int, const float*,
                         const float*, const hgs_raster_grads*, float*, cudaStream_t);
int launch_tile_binning(int, int, int64_t, const GeomLayout&, const BinningLayout&, const ImageLayout&, uint32_t, uint32_t,
                        uint32_t, int, cudaStream_t);
int set_fwd_stats(void* dev_ptr);
int set_composite_blocks(int mode);
int launch_weighted_l1(int, long long, const float*, const float*, const float*, float*, float*, cudaStream_t);
int launch_hair_image_loss(const hgs_hair_loss&, cudaStream_t);
int launch_unpack_targets(int, long long, const void*, const void*, float*, cudaStream_t);
int launch_adam_flat(long long, float*, float*, float*, float*, int, const int64_t*, const float*, int, float, float, float,
                     float, int, cudaStream_t);
int launch_densify_stats(int, const int*, const float*, int, int*, float*, float*, float*, cudaStream_t);
int launch_merge_candidates(const MergeArgs&, bool, int*, const long long*, int*, int*, float*, cudaStream_t);
int launch_merge_greedy(long long, const int*, const int*, const int*, unsigned char*, unsigned char*, cudaStream_t);
size_t knn_bytes(int P);
int launch_knn(int P, const float* points, float* out, void* ws, cudaStream_t s);

static int validate(const hgs_raster_params* prm, const hgs_raster_inputs* in, bool forward = true) {
    if (!prm || !in) { set_error("null params"); return HGS_ERR_INVALID; }
    if (prm->P < 0 || prm->width <= 0 || prm->height <= 0) { set_error("bad sizes P=%d W=%d H=%d", prm->P, prm->width, prm->height); return HGS_ERR_INVALID; }
    if (prm->channels < 1 || prm->channels > HGS_MAX_CHANNELS) { set_error("channels must be 1..%d", HGS_MAX_CHANNELS); return HGS_ERR_INVALID; }
    if (prm->channels != 3 && in->colors_precomp == nullptr) {
        // rasterizer_impl.cu:242-245
        set_error("For non-RGB, provide precomputed Gaussian colors!");
        return HGS_ERR_INVALID;
    }
    if ((prm->width + HGS_TILE - 1) / HGS_TILE > 0xffff || (prm->height + HGS_TILE - 1) / HGS_TILE > 0xffff) { set_error("image too large"); return HGS_ERR_INVALID; }
    if (prm->P > 0) {
        if (!in->means3D || (forward && !in->opacities) || !in->viewmatrix || !in->projmatrix || !in->background) { set_error("missing required input pointer"); return HGS_ERR_INVALID; }
        if (!in->colors_precomp && (!in->shs || prm->M <= 0 || !in->cam_pos)) { set_error("need shs (+campos) or colors_precomp"); return HGS_ERR_INVALID; }
        if (!in->colors_precomp && (prm->D < 0 || prm->D > 3 || (prm->D + 1) * (prm->D + 1) > prm->M)) { set_error("SH degree %d incompatible with M=%d", prm->D, prm->M); return HGS_ERR_INVALID; }
        if (!in->cov3D_precomp && (!in->scales || !in->rotations)) { set_error("need scales+rotations or cov3D_precomp"); return HGS_ERR_INVALID; }
    }
    return HGS_OK;
}

// bits of the depth part of the sort key: 32 (the reference's key, rasterizer_impl.cu:300-308) unless the caller passes a
// smaller range hint (see KeyXform in binning.cu)
static inline int depth_bits_for(const hgs_raster_params* prm) {
    return (prm->sort_depth_bits >= 1 && prm->sort_depth_bits < 32) ? prm->sort_depth_bits : 32;
}
static inline int end_bit_for(const hgs_raster_params* prm) {
    const uint32_t gx = (prm->width + HGS_TILE - 1) / HGS_TILE, gy = (prm->height + HGS_TILE - 1) / HGS_TILE;
    return depth_bits_for(prm) + tile_id_bits(gx * gy);
}

}  // namespace hgs

using namespace hgs;

extern "C" {

int hgs_debug_set_stats(void* dev_ptr) { return set_fwd_stats(dev_ptr); }

int hgs_debug_set_composite_blocks(int mode) { return set_composite_blocks(mode); }

int hgs_profile_enable(int on) {
    g_prof_on.store(on != 0);
    return HGS_OK;
}

int hgs_profile_collect(double* ms, int64_t* launches) {
    if (int e = check_cuda(cudaDeviceSynchronize(), "profile sync")) return e;
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    for (auto& r : g_prof) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && ms) ms[r.stage] += (double)t;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof.clear();
    for (int i = 0; i < HGS_STAGE_COUNT; ++i) {
        const long long n = g_launches[i].exchange(0);
        if (launches) launches[i] += n;
    }
    return HGS_OK;
}

const char* hgs_stage_name(int stage) {
    static const char* names[HGS_STAGE_COUNT] = {"preprocess_fwd", "emit_keys", "sort_histogram", "sort_onesweep",
                                                  "tile_ranges", "composite_fwd", "composite_bwd", "preprocess_bwd",
                                                  "knn", "other", "tile_scan", "tile_count", "tile_offsets", "tile_scatter",
                                                  "tile_sort_pack"};
    return (stage >= 0 && stage < HGS_STAGE_COUNT) ? names[stage] : "?";
}

int hgs_abi_version(void) { return HGS_ABI_VERSION; }
const char* hgs_last_error(void) { return g_err; }

size_t hgs_geom_bytes(int32_t P, int32_t channels, int32_t width, int32_t height) {
    return carve_geom(nullptr, P, channels, tile_count_of(width, height)).bytes;
}
size_t hgs_image_bytes(int32_t width, int32_t height) { return carve_image(nullptr, width, height).bytes; }
size_t hgs_binning_bytes(int64_t n, int32_t channels) { return carve_binning(nullptr, n, channels).bytes; }
size_t hgs_sort_bytes(int64_t n) { return carve_sort(nullptr, n).bytes; }

// Inverse of hgs_binning_bytes for the two ways a binning workspace gets sized: exactly num_rendered
// (reference-style, after the blocking read-back) or a capacity that is a multiple of 4096 (sync-free mode).
int64_t hgs_binning_capacity(size_t bytes, int32_t channels, int64_t num_rendered) {
    if (num_rendered >= 0 && carve_binning(nullptr, num_rendered, channels).bytes == bytes) return num_rendered;
    int64_t lo = 0, hi = (int64_t)1 << 19;  // in units of 4096 instances (2^31 total)
    while (lo < hi) {
        const int64_t mid = (lo + hi) / 2;
        if (carve_binning(nullptr, mid * 4096, channels).bytes < bytes) lo = mid + 1; else hi = mid;
    }
    if (carve_binning(nullptr, lo * 4096, channels).bytes == bytes) return lo * 4096;
    set_error("binning workspace of %zu bytes matches no capacity", bytes);
    return HGS_ERR_INVALID;
}

int hgs_forward_stage_a(const hgs_raster_params* prm, const hgs_raster_inputs* in, void* geom_ws, int32_t* radii,
                        void* stream) {
    if (int e = validate(prm, in)) return e;
    cudaStream_t s = (cudaStream_t)stream;
    if (!geom_ws) { set_error("null geometry workspace"); return HGS_ERR_INVALID; }
    GeomLayout g = carve_geom(geom_ws, prm->P, prm->channels, tile_count_of(prm->width, prm->height));
    if (prm->P == 0) return check_cuda(cudaMemsetAsync(g.hdr, 0, g.clear_bytes, s), "memset geom header");
    if (int e = launch_preprocess_fwd(prm, in, g, radii, s)) return e;
    return stage_check("preprocess", prm->debug, s);
}

int hgs_forward_read_num_rendered(const void* geom_ws, int32_t P, uint32_t* n_pinned_host, void* stream) {
    (void)P;
    const GeomHeader* h = (const GeomHeader*)geom_ws;
    return check_cuda(cudaMemcpyAsync(n_pinned_host, &h->num_rendered, 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                      (cudaStream_t)stream), "read num_rendered");
}

// bin + sort + finalize (parts & 1) and composite (parts & 2), shared by the generic and the strand entry
static int stage_b_impl(const hgs_raster_params* prm, const float* background, void* geom_ws, void* binning_ws,
                        void* image_ws, int64_t N, float* out_color, cudaStream_t s, int parts = 3) {
    if (!geom_ws || !image_ws || (N > 0 && !binning_ws) || ((parts & 2) && !out_color)) { set_error("null workspace/output"); return HGS_ERR_INVALID; }
    if (N < 0 || N > 0x7fffffffll) { set_error("num_rendered out of range"); return HGS_ERR_OVERFLOW; }
    GeomLayout g = carve_geom(geom_ws, prm->P, prm->channels, tile_count_of(prm->width, prm->height));
    ImageLayout im = carve_image(image_ws, prm->width, prm->height);
    BinningLayout b = carve_binning(binning_ws, N, prm->channels);
    const uint32_t gx = (prm->width + HGS_TILE - 1) / HGS_TILE, gy = (prm->height + HGS_TILE - 1) / HGS_TILE;

    if ((parts & 1) && prm->sort_mode != HGS_SORT_GLOBAL) {
        // HGS_SORT_TILE: partition by tile + in-tile sort (tilesort.cu); the sorted pairs end in ping-pong buffer 0
        if (int e = launch_tile_binning(N > 0 ? prm->P : 0, prm->channels, N, g, b, im, gx, gy, (uint32_t)prm->slice_base,
                                        prm->slice_shift <= 0 ? 32 : prm->slice_shift, s)) return e;
        if (int e = stage_check("tile_binning", prm->debug, s)) return e;
    } else if (parts & 1) {
        // the unsorted pairs go into the ping-pong buffer from which the sort's passes end in buffer 0, whatever their number
        const int start = sort_passes(end_bit_for(prm)) & 1;
        if (int e = launch_emit_keys(N > 0 ? prm->P : 0, g, g.rects, b.keys[start], b.vals[start], gx, (uint32_t)N, s)) return e;
        if (int e = stage_check("emit_keys", prm->debug, s)) return e;
        int res = 0;
        const uint32_t* n_ptr = &g.hdr->num_rendered;  // live instance count stays on the device; N is only the capacity
        if (int e = launch_sort_pairs(N, n_ptr, end_bit_for(prm), b.keys, b.vals, b.sort_ws, &res, s, g.hdr, depth_bits_for(prm),
                                      start)) return e;
        if (int e = stage_check("sort", prm->debug, s)) return e;
        if (int e = launch_finalize_sorted(prm->channels, N, n_ptr, b.keys[res], b.vals[res], g, b, im.ranges, im.tile_order, (size_t)gx * gy, gx, s)) return e;
        if (int e = stage_check("finalize_sorted", prm->debug, s)) return e;
    }
    if (parts & 2) {
        if (!background) { set_error("null background"); return HGS_ERR_INVALID; }
        if (int e = launch_composite_fwd(prm->channels, im, b, prm->width, prm->height, background, out_color, s)) return e;
        return stage_check("composite_fwd", prm->debug, s);
    }
    return HGS_OK;
}

static int validate_parts(const hgs_raster_params* prm) {
    if (!prm) { set_error("null params"); return HGS_ERR_INVALID; }
    if (prm->P < 0 || prm->width <= 0 || prm->height <= 0) { set_error("bad sizes P=%d W=%d H=%d", prm->P, prm->width, prm->height); return HGS_ERR_INVALID; }
    if (prm->channels < 1 || prm->channels > HGS_MAX_CHANNELS) { set_error("channels must be 1..%d", HGS_MAX_CHANNELS); return HGS_ERR_INVALID; }
    return HGS_OK;
}

int hgs_forward_stage_b_binning(const hgs_raster_params* prm, void* geom_ws, void* binning_ws, void* image_ws, int64_t N,
                                void* stream) {
    if (int e = validate_parts(prm)) return e;
    return stage_b_impl(prm, nullptr, geom_ws, binning_ws, image_ws, N, nullptr, (cudaStream_t)stream, 1);
}

int hgs_forward_stage_b_composite(const hgs_raster_params* prm, const float* background, void* geom_ws, void* binning_ws,
                                  void* image_ws, int64_t N, float* out_color, void* stream) {
    if (int e = validate_parts(prm)) return e;
    return stage_b_impl(prm, background, geom_ws, binning_ws, image_ws, N, out_color, (cudaStream_t)stream, 2);
}

int hgs_forward_stage_b(const hgs_raster_params* prm, const hgs_raster_inputs* in, void* geom_ws, void* binning_ws,
                        void* image_ws, int64_t N, const int32_t* radii, float* out_color, void* stream) {
    (void)radii;
    if (int e = validate(prm, in)) return e;
    return stage_b_impl(prm, in->background, geom_ws, binning_ws, image_ws, N, out_color, (cudaStream_t)stream);
}

static int validate_strands(const hgs_raster_params* prm, const hgs_strand_inputs* in) {
    if (!prm || !in) { set_error("null params"); return HGS_ERR_INVALID; }
    if (prm->channels != 7) { set_error("strand entry renders exactly 7 channels (rgb, mask, orientation)"); return HGS_ERR_INVALID; }
    if (prm->P < 0 || prm->width <= 0 || prm->height <= 0) { set_error("bad sizes"); return HGS_ERR_INVALID; }
    if ((prm->width + HGS_TILE - 1) / HGS_TILE > 0xffff || (prm->height + HGS_TILE - 1) / HGS_TILE > 0xffff) { set_error("image too large"); return HGS_ERR_INVALID; }
    if (prm->D < 0 || prm->D > 3 || (prm->D + 1) * (prm->D + 1) > prm->M) { set_error("SH degree %d incompatible with M=%d", prm->D, prm->M); return HGS_ERR_INVALID; }
    if (prm->P > 0 && (!in->endpoints || !in->endpoint_pairs || !in->width || !in->opacity_logit || !in->mask_logit ||
                       !in->features || !in->viewmatrix || !in->projmatrix || !in->cam_pos || !in->background)) {
        set_error("missing required strand input pointer");
        return HGS_ERR_INVALID;
    }
    return HGS_OK;
}

int hgs_strands_forward_stage_a(const hgs_raster_params* prm, const hgs_strand_inputs* in, void* geom_ws, int32_t* radii,
                                void* stream) {
    if (int e = validate_strands(prm, in)) return e;
    cudaStream_t s = (cudaStream_t)stream;
    if (!geom_ws) { set_error("null geometry workspace"); return HGS_ERR_INVALID; }
    GeomLayout g = carve_geom(geom_ws, prm->P, prm->channels, tile_count_of(prm->width, prm->height));
    if (prm->P == 0) return check_cuda(cudaMemsetAsync(g.hdr, 0, g.clear_bytes, s), "memset geom header");
    if (int e = launch_strand_preprocess_fwd(prm, in, g, radii, s)) return e;
    return stage_check("strand preprocess", prm->debug, s);
}

int hgs_strands_forward_stage_b(const hgs_raster_params* prm, const hgs_strand_inputs* in, void* geom_ws, void* binning_ws,
                                void* image_ws, int64_t capacity, float* out_color, void* stream) {
    if (int e = validate_strands(prm, in)) return e;
    return stage_b_impl(prm, in->background, geom_ws, binning_ws, image_ws, capacity, out_color, (cudaStream_t)stream);
}

int hgs_strands_backward(const hgs_raster_params* prm, const hgs_strand_inputs* in, int64_t R, const void* geom_ws,
                         const void* binning_ws, const void* image_ws, const float* dL_dpix, const hgs_strand_grads* gr,
                         void* stream) {
    return hgs_strands_backward_parts(prm, in, R, geom_ws, binning_ws, image_ws, dL_dpix, gr, 3, stream);
}

int hgs_strands_backward_parts(const hgs_raster_params* prm, const hgs_strand_inputs* in, int64_t R, const void* geom_ws,
                               const void* binning_ws, const void* image_ws, const float* dL_dpix, const hgs_strand_grads* gr,
                               int32_t parts, void* stream) {
    if (int e = validate_strands(prm, in)) return e;
    cudaStream_t s = (cudaStream_t)stream;
    if (!gr || !gr->dL_dmean2D || !gr->dL_dendpoints || !gr->dL_dwidth || !gr->dL_dopacity_logit || !gr->dL_dmask_logit ||
        !gr->dL_dfeatures || (!gr->acc16 && (!gr->dL_dconic || !gr->dL_dopacity || !gr->dL_dcolor))) { set_error("missing gradient output pointer"); return HGS_ERR_INVALID; }
    if (gr->acc16 && ((uintptr_t)gr->acc16 & 15)) { set_error("acc16 must be 16-byte aligned"); return HGS_ERR_INVALID; }
    if (!geom_ws || !image_ws || !dL_dpix || (R > 0 && !binning_ws)) { set_error("null workspace"); return HGS_ERR_INVALID; }
    const int P = prm->P;
    if ((parts & 2) && !gr->accumulate)
        if (int e = check_cuda(cudaMemsetAsync(gr->dL_dendpoints, 0, (size_t)in->num_endpoints * 3 * 4, s), "memset dL_dendpoints")) return e;
    if (P == 0) return HGS_OK;
    GeomLayout g = carve_geom((void*)geom_ws, P, prm->channels, tile_count_of(prm->width, prm->height));
    ImageLayout im = carve_image((void*)image_ws, prm->width, prm->height);
    BinningLayout b = carve_binning((void*)binning_ws, R, prm->channels);
    const int res = 0;  // the sorted pairs always end in ping-pong buffer 0 (stage_b_impl)
    const size_t Pz = (size_t)P;
    if (!(parts & 1)) {
        // part 2 only: the compositor's accumulators are already filled
    } else if (gr->acc16) {
        if (int e = check_cuda(cudaMemsetAsync(gr->acc16, 0, Pz * 16 * 4, s), "memset acc16")) return e;
    } else if (gr->dL_dconic == gr->dL_dmean2D + 3 * Pz && gr->dL_dopacity == gr->dL_dconic + 4 * Pz &&
        gr->dL_dcolor == gr->dL_dopacity + Pz) {
        if (int e = check_cuda(cudaMemsetAsync(gr->dL_dmean2D, 0, Pz * (8 + 7) * 4, s), "memset grads")) return e;
    } else {
        if (int e = check_cuda(cudaMemsetAsync(gr->dL_dmean2D, 0, Pz * 3 * 4, s), "memset dL_dmean2D")) return e;
        if (int e = check_cuda(cudaMemsetAsync(gr->dL_dconic, 0, Pz * 4 * 4, s), "memset dL_dconic")) return e;
        if (int e = check_cuda(cudaMemsetAsync(gr->dL_dopacity, 0, Pz * 4, s), "memset dL_dopacity")) return e;
        if (int e = check_cuda(cudaMemsetAsync(gr->dL_dcolor, 0, Pz * 7 * 4, s), "memset dL_dcolor")) return e;
    }
    if ((parts & 1) && R > 0) {
        hgs_raster_grads rg;
        memset(&rg, 0, sizeof(rg));
        rg.dL_dmean2D = gr->dL_dmean2D; rg.dL_dconic = gr->dL_dconic; rg.dL_dopacity = gr->dL_dopacity; rg.dL_dcolor = gr->dL_dcolor;
        if (int e = launch_composite_bwd(7, im, b, b.vals[res], prm->width, prm->height, in->background, dL_dpix, &rg, gr->acc16, s)) return e;
        if (int e = stage_check("composite_bwd", prm->debug, s)) return e;
    }
    if (!(parts & 2)) return HGS_OK;
    if (int e = launch_strand_preprocess_bwd(prm, in, g, gr, s)) return e;
    return stage_check("strand preprocess_bwd", prm->debug, s);
}

int hgs_graph_instantiate(void* graph, int32_t use_node_priority, void** exec) {
    if (!graph || !exec) { set_error("null graph"); return HGS_ERR_INVALID; }
    cudaGraphExec_t e = nullptr;
    const unsigned long long flags = use_node_priority ? cudaGraphInstantiateFlagUseNodePriority : 0ull;
    if (int err = check_cuda(cudaGraphInstantiateWithFlags(&e, (cudaGraph_t)graph, flags), "cudaGraphInstantiateWithFlags")) return err;
    *exec = (void*)e;
    return HGS_OK;
}

int hgs_graph_launch(void* exec, void* stream) {
    if (!exec) { set_error("null graph exec"); return HGS_ERR_INVALID; }
    return check_cuda(cudaGraphLaunch((cudaGraphExec_t)exec, (cudaStream_t)stream), "cudaGraphLaunch");
}

int hgs_graph_exec_destroy(void* exec) {
    if (!exec) return HGS_OK;
    return check_cuda(cudaGraphExecDestroy((cudaGraphExec_t)exec), "cudaGraphExecDestroy");
}

int hgs_rasterize_forward(hgs_alloc_fn geom_alloc, void* geom_user, hgs_alloc_fn binning_alloc, void* binning_user,
                          hgs_alloc_fn image_alloc, void* image_user, const hgs_raster_params* prm,
                          const hgs_raster_inputs* in, float* out_color, int32_t* radii, void* stream) {
    if (int e = validate(prm, in)) return e;
    if (!geom_alloc || !binning_alloc || !image_alloc) { set_error("null allocator"); return HGS_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    void* geom = geom_alloc(geom_user, hgs_geom_bytes(prm->P, prm->channels, prm->width, prm->height));
    void* img = image_alloc(image_user, hgs_image_bytes(prm->width, prm->height));
    if (!geom || !img) { set_error("allocator returned NULL"); return HGS_ERR_ALLOC; }
    if (int e = hgs_forward_stage_a(prm, in, geom, radii, s)) return e;
    // the one blocking read-back of the pass (rasterizer_impl.cu:281)
    uint32_t host[3] = {0, 0, 0};
    if (int e = check_cuda(cudaMemcpyAsync(host, geom, sizeof(host), cudaMemcpyDeviceToHost, s), "read num_rendered")) return e;
    if (int e = check_cuda(cudaStreamSynchronize(s), "sync num_rendered")) return e;
    if ((host[2] & 1u) || host[0] > 0x7fffffffu) { set_error("instance count overflows int32"); return HGS_ERR_OVERFLOW; }
    const int64_t N = host[0];
    void* bin = binning_alloc(binning_user, hgs_binning_bytes(N, prm->channels));
    if (!bin) { set_error("allocator returned NULL"); return HGS_ERR_ALLOC; }
    // stage A counted the tile lists: one longer than HGS_TILE_SORT_MAX (bit 2) needs the global radix sort
    hgs_raster_params p2 = *prm;
    if (host[2] & 4u) p2.sort_mode = HGS_SORT_GLOBAL;
    if (int e = hgs_forward_stage_b(&p2, in, geom, bin, img, N, radii, out_color, s)) return e;
    return (int)N;
}

int hgs_rasterize_backward(const hgs_raster_params* prm, const hgs_raster_inputs* in, int64_t R, const int32_t* radii,
                           const void* geom_ws, const void* binning_ws, const void* image_ws, const float* dL_dpix,
                           const hgs_raster_grads* gr, void* stream) {
    // R is the CAPACITY the binning workspace was carved with (== num_rendered in the reference-style exact mode)
    if (int e = validate(prm, in, false)) return e;
    cudaStream_t s = (cudaStream_t)stream;
    if (!gr || !gr->dL_dmean2D || !gr->dL_dconic || !gr->dL_dopacity || !gr->dL_dcolor || !gr->dL_dmean3D ||
        !gr->dL_dcov3D || !gr->dL_dscale || !gr->dL_drot || (prm->M > 0 && !gr->dL_dsh)) { set_error("missing gradient output pointer"); return HGS_ERR_INVALID; }
    if (!geom_ws || !image_ws || !dL_dpix || (R > 0 && !binning_ws)) { set_error("null workspace"); return HGS_ERR_INVALID; }
    const int P = prm->P;
    if (P == 0) return HGS_OK;
    GeomLayout g = carve_geom((void*)geom_ws, P, prm->channels, tile_count_of(prm->width, prm->height));
    ImageLayout im = carve_image((void*)image_ws, prm->width, prm->height);
    BinningLayout b = carve_binning((void*)binning_ws, R, prm->channels);
    const int res = 0;  // the sorted pairs always end in ping-pong buffer 0 (stage_b_impl)

    // accumulation targets of the compositor (the only arrays that need clearing)
    const size_t Pz = (size_t)P;
    if (gr->dL_dconic == gr->dL_dmean2D + 3 * Pz && gr->dL_dopacity == gr->dL_dconic + 4 * Pz &&
        gr->dL_dcolor == gr->dL_dopacity + Pz) {
        // caller packed the four arrays back to back: one memset
        if (int e = check_cuda(cudaMemsetAsync(gr->dL_dmean2D, 0, Pz * (8 + prm->channels) * 4, s), "memset grads")) return e;
    } else {
        if (int e = check_cuda(cudaMemsetAsync(gr->dL_dmean2D, 0, Pz * 3 * 4, s), "memset dL_dmean2D")) return e;
        if (int e = check_cuda(cudaMemsetAsync(gr->dL_dconic, 0, Pz * 4 * 4, s), "memset dL_dconic")) return e;
        if (int e = check_cuda(cudaMemsetAsync(gr->dL_dopacity, 0, Pz * 4, s), "memset dL_dopacity")) return e;
        if (int e = check_cuda(cudaMemsetAsync(gr->dL_dcolor, 0, Pz * prm->channels * 4, s), "memset dL_dcolor")) return e;
    }
    if (R > 0) {
        if (int e = launch_composite_bwd(prm->channels, im, b, b.vals[res], prm->width, prm->height, in->background,
                                         dL_dpix, gr, nullptr, s)) return e;
        if (int e = stage_check("composite_bwd", prm->debug, s)) return e;
    }
    if (int e = launch_preprocess_bwd(prm, in, g, radii, gr, s)) return e;
    return stage_check("preprocess_bwd", prm->debug, s);
}

int hgs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream) {
    (void)projmatrix;
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) { set_error("bad mark_visible args"); return HGS_ERR_INVALID; }
    if (P == 0) return HGS_OK;
    return launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
}

int hgs_weighted_l1(int32_t C, int64_t HW, const float* image, const float* target, const float* weights, float* loss,
                    float* dL_dimage, void* stream) {
    if (C < 0 || HW < 0 || (C > 0 && HW > 0 && (!image || !target || !weights || !dL_dimage)) || !loss) { set_error("bad weighted_l1 args"); return HGS_ERR_INVALID; }
    return launch_weighted_l1(C, HW, image, target, weights, loss, dL_dimage, (cudaStream_t)stream);
}

int hgs_hair_image_loss(const hgs_hair_loss* a, void* stream) {
    if (!a || a->height <= 0 || a->width <= 0 || !a->image7 || !a->gt_rgb || !a->gt_mask || !a->gt_theta || !a->confidence ||
        !a->terms || !a->scratch || !a->dL_dimage) { set_error("bad hair_image_loss args"); return HGS_ERR_INVALID; }
    return launch_hair_image_loss(*a, (cudaStream_t)stream);
}

int hgs_unpack_targets(int32_t views, int64_t HW, const void* rgbm, const void* theta_conf, float* out, void* stream) {
    if (views < 0 || HW < 0 || (views > 0 && HW > 0 && (!rgbm || !theta_conf || !out))) { set_error("bad unpack_targets args"); return HGS_ERR_INVALID; }
    return launch_unpack_targets(views, HW, rgbm, theta_conf, out, (cudaStream_t)stream);
}

int hgs_adam_step(int64_t n, float* param, float* grad, float* exp_avg, float* exp_avg_sq, int32_t n_groups,
                  const int64_t* group_end, const float* lr, int32_t step, float beta1, float beta2, float eps,
                  float grad_scale, int32_t zero_grad, void* stream) {
    if (n < 0 || n_groups < 1 || n_groups > HGS_ADAM_MAX_GROUPS || step < 1 || !group_end || !lr ||
        (n > 0 && (!param || !grad || !exp_avg || !exp_avg_sq))) { set_error("bad adam_step args"); return HGS_ERR_INVALID; }
    for (int g = 0; g < n_groups; ++g)
        if (group_end[g] < (g ? group_end[g - 1] : 0) || group_end[g] > n) { set_error("adam_step: group_end must be non-decreasing and <= n"); return HGS_ERR_INVALID; }
    if (group_end[n_groups - 1] != n) { set_error("adam_step: groups must cover the bucket"); return HGS_ERR_INVALID; }
    return launch_adam_flat(n, param, grad, exp_avg, exp_avg_sq, n_groups, group_end, lr, step, beta1, beta2, eps, grad_scale,
                            zero_grad, (cudaStream_t)stream);
}

int hgs_densify_stats(int32_t P, const int32_t* radii, const float* dL_dmean2D, int32_t grad_stride, float* max_radii2D,
                      float* xyz_gradient_accum, float* denom, void* stream) {
    if (P < 0 || grad_stride < 2 || (P > 0 && (!radii || !dL_dmean2D || !max_radii2D || !xyz_gradient_accum || !denom))) {
        set_error("bad densify_stats args"); return HGS_ERR_INVALID;
    }
    return launch_densify_stats(P, radii, dL_dmean2D, grad_stride, nullptr, max_radii2D, xyz_gradient_accum, denom,
                                (cudaStream_t)stream);
}

static int merge_args(MergeArgs& a, int32_t K, const float* points, const float* dirs, const int32_t* global_id,
                      const int32_t* other_end, double radius, double dir_th, int32_t bidirectional, int32_t max_num_nn) {
    if (K < 0 || !(radius >= 0.0) || (K > 0 && (!points || !dirs || !global_id || !other_end))) {
        set_error("bad merge search args"); return HGS_ERR_INVALID;
    }
    a = MergeArgs{K, points, dirs, global_id, other_end, radius * radius, dir_th, bidirectional, max_num_nn};
    return HGS_OK;
}

int hgs_merge_count(int32_t K, const float* points, const float* dirs, const int32_t* global_id, const int32_t* other_end,
                    double radius, double dir_th, int32_t bidirectional, int32_t max_num_nn, int32_t* counts, void* stream) {
    MergeArgs a;
    if (int e = merge_args(a, K, points, dirs, global_id, other_end, radius, dir_th, bidirectional, max_num_nn)) return e;
    if (K > 0 && !counts) { set_error("bad merge search args"); return HGS_ERR_INVALID; }
    return launch_merge_candidates(a, false, counts, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int hgs_merge_fill(int32_t K, const float* points, const float* dirs, const int32_t* global_id, const int32_t* other_end,
                   double radius, double dir_th, int32_t bidirectional, int32_t max_num_nn, const int64_t* offsets,
                   int32_t* p1, int32_t* p2, float* dist, void* stream) {
    MergeArgs a;
    if (int e = merge_args(a, K, points, dirs, global_id, other_end, radius, dir_th, bidirectional, max_num_nn)) return e;
    if (K > 0 && (!offsets || !p1 || !p2 || !dist)) { set_error("bad merge search args"); return HGS_ERR_INVALID; }
    return launch_merge_candidates(a, true, nullptr, (const long long*)offsets, p1, p2, dist, (cudaStream_t)stream);
}

int hgs_merge_greedy(int64_t n, const int32_t* p1, const int32_t* p2, const int32_t* other_end_of, uint8_t* flags,
                     uint8_t* keep, void* stream) {
    if (n < 0 || (n > 0 && (!p1 || !p2 || !other_end_of || !flags || !keep))) { set_error("bad merge_greedy args"); return HGS_ERR_INVALID; }
    return launch_merge_greedy(n, p1, p2, other_end_of, flags, keep, (cudaStream_t)stream);
}

size_t hgs_knn_bytes(int32_t P) { return knn_bytes(P); }

int hgs_dist2_knn3(int32_t P, const float* points, float* mean_dist2, void* workspace, void* stream) {
    if (P < 0 || (P > 0 && (!points || !mean_dist2 || !workspace))) { set_error("bad knn args"); return HGS_ERR_INVALID; }
    if (P == 0) return HGS_OK;
    return launch_knn(P, points, mean_dist2, workspace, (cudaStream_t)stream);
}

int hgs_sort_pairs(int64_t n, int end_bit, uint64_t* keys_in, uint32_t* vals_in, uint64_t* keys_out, uint32_t* vals_out,
                   void* workspace, void* stream) {
    if (n < 0 || end_bit < 0 || end_bit > 64) { set_error("bad sort args"); return HGS_ERR_INVALID; }
    if (n == 0) return HGS_OK;
    cudaStream_t s = (cudaStream_t)stream;
    uint64_t* keys[2] = {keys_in, keys_out};
    uint32_t* vals[2] = {vals_in, vals_out};
    int res = 0;
    if (int e = launch_sort_pairs(n, nullptr, end_bit, keys, vals, workspace, &res, s, nullptr, 32, 0)) return e;
    if (res == 0) {
        if (int e = check_cuda(cudaMemcpyAsync(keys_out, keys_in, (size_t)n * 8, cudaMemcpyDeviceToDevice, s), "copy keys")) return e;
        if (int e = check_cuda(cudaMemcpyAsync(vals_out, vals_in, (size_t)n * 4, cudaMemcpyDeviceToDevice, s), "copy vals")) return e;
    }
    return HGS_OK;
}

int64_t hgs_state_view(int what, const hgs_raster_params* prm, const hgs_raster_inputs* in, int64_t N, int64_t capacity,
                       const void* geom_ws, const void* binning_ws, const void* image_ws, void* dst, void* stream) {
    if (!prm || !dst) { set_error("null args"); return HGS_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    const int P = prm->P;
    const size_t hw = (size_t)prm->width * prm->height;
    const size_t tiles = (size_t)((prm->width + HGS_TILE - 1) / HGS_TILE) * ((prm->height + HGS_TILE - 1) / HGS_TILE);
    auto d2d = [&](const void* src, size_t bytes) -> int64_t {
        if (bytes == 0) return 0;
        if (int e = check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s), "state view copy")) return e;
        return (int64_t)bytes;
    };
    switch (what) {
        case HGS_VIEW_DEPTHS: case HGS_VIEW_MEANS2D: case HGS_VIEW_CONIC_OPACITY: case HGS_VIEW_RGB:
        case HGS_VIEW_CLAMPED: case HGS_VIEW_COV3D: {
            if (!geom_ws || !in) { set_error("null geometry workspace"); return HGS_ERR_INVALID; }
            if (P == 0) return 0;
            GeomLayout g = carve_geom((void*)geom_ws, P, prm->channels, tile_count_of(prm->width, prm->height));
            if (int e = launch_view_geom(what, prm, in, g, dst, s)) return e;
            static const int per[] = {4, 8, 16, 0, 0, 0, 3};
            if (what == HGS_VIEW_RGB) return (int64_t)P * prm->channels * 4;
            if (what == HGS_VIEW_COV3D) return (int64_t)P * 24;
            return (int64_t)P * per[what];
        }
        case HGS_VIEW_TILES_TOUCHED: case HGS_VIEW_POINT_OFFSETS: {
            if (!geom_ws) { set_error("null geometry workspace"); return HGS_ERR_INVALID; }
            GeomLayout g = carve_geom((void*)geom_ws, P, prm->channels, tile_count_of(prm->width, prm->height));
            return d2d(what == HGS_VIEW_TILES_TOUCHED ? (void*)g.tiles_touched : (void*)g.offsets, (size_t)P * 4);
        }
        case HGS_VIEW_KEYS_SORTED: case HGS_VIEW_POINT_LIST: {
            if (N > 0 && !binning_ws) { set_error("null binning workspace"); return HGS_ERR_INVALID; }
            BinningLayout b = carve_binning((void*)binning_ws, capacity, prm->channels);
            const int res = 0;  // the sorted pairs always end in ping-pong buffer 0 (stage_b_impl)
            if (what == HGS_VIEW_KEYS_SORTED) return d2d(b.keys[res], (size_t)N * 8);
            return d2d(b.vals[res], (size_t)N * 4);
        }
        case HGS_VIEW_RANGES: case HGS_VIEW_FINAL_T: case HGS_VIEW_N_CONTRIB: {
            if (!image_ws) { set_error("null image workspace"); return HGS_ERR_INVALID; }
            ImageLayout im = carve_image((void*)image_ws, prm->width, prm->height);
            if (what == HGS_VIEW_RANGES) return d2d(im.ranges, tiles * 8);
            if (what == HGS_VIEW_FINAL_T) return d2d(im.final_T, hw * 4);
            return d2d(im.n_contrib, hw * 4);
        }
        default: break;
    }
    set_error("unknown view %d", what);
    return HGS_ERR_INVALID;
}

}  // extern "C"
