// hgs_torch_ext.cpp — the torch-extension form of the drop-in binding: the same three entry points, argument order and
// return tuples as the reference's pybind module (submodules/diff-gaussian-rasterization/ext.cpp:15-19, signatures
// rasterize_points.h:18-66), implemented over the C ABI of libhairgs_rast.so (include/hairgs_rast.h) exactly the way
// INTEGRATION.md describes: a maintainer who prefers a compiled binding over the shipped ctypes one (1 ms of Python per
// eager view) builds this file against the library.  Built and tested by tests/test_torch_ext.py
// (hair-gs_b200/torch_ext/build.py); no CUDA code here - marshalling only.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <stdexcept>
#include <tuple>

#include "../../include/hairgs_rast.h"

namespace {

void* resize_cb(void* user, size_t bytes) {  // the reference's resizeFunctional (rasterize_points.cu:27-33)
    auto* t = static_cast<torch::Tensor*>(user);
    t->resize_({(long long)(bytes ? bytes : 1)});
    return t->data_ptr();
}

const float* fptr(const torch::Tensor& t) { return t.numel() ? t.data_ptr<float>() : nullptr; }

torch::Tensor contig(const torch::Tensor& t) { return t.numel() ? t.contiguous() : t; }

void check(int status, const char* what) {
    if (status < 0) throw std::runtime_error(std::string(what) + ": " + hgs_last_error());
}

}  // namespace

std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> RasterizeGaussiansCUDA(
    const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors, const torch::Tensor& opacity,
    const torch::Tensor& scales, const torch::Tensor& rotations, const float scale_modifier, const torch::Tensor& cov3D_precomp,
    const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
    const int image_height, const int image_width, const torch::Tensor& sh, const int degree, const torch::Tensor& campos,
    const bool prefiltered, const bool debug) {
    if (means3D.ndimension() != 2 || means3D.size(1) != 3) AT_ERROR("means3D must have dimensions (num_points, 3)");
    TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor: this rasterizer has no CPU path");
    const c10::cuda::CUDAGuard guard(means3D.device());
    const int P = (int)means3D.size(0), H = image_height, W = image_width;
    const int C = colors.numel() ? (int)colors.size(-1) : 3;
    auto f32 = means3D.options().dtype(torch::kFloat32);
    auto out_color = torch::empty({C, H, W}, f32);
    auto radii = torch::empty({P}, means3D.options().dtype(torch::kInt32));
    auto u8 = torch::TensorOptions(torch::kByte).device(means3D.device());
    torch::Tensor geom = torch::empty({0}, u8), binning = torch::empty({0}, u8), img = torch::empty({0}, u8);
    if (P == 0) {  // rasterize_points.cu:81
        out_color.zero_();
        return std::make_tuple(0, out_color, radii, geom, binning, img);
    }
    const auto bg = contig(background), m3 = contig(means3D), col = contig(colors), op = contig(opacity), sc = contig(scales),
               rot = contig(rotations), cov = contig(cov3D_precomp), vm = contig(viewmatrix), pm = contig(projmatrix),
               shc = contig(sh), cp = contig(campos);
    hgs_raster_params prm{};
    prm.P = P; prm.D = degree; prm.M = sh.numel() ? (int)sh.size(1) : 0; prm.width = W; prm.height = H; prm.channels = C;
    prm.tan_fovx = tan_fovx; prm.tan_fovy = tan_fovy; prm.scale_modifier = scale_modifier;
    prm.prefiltered = prefiltered; prm.debug = debug; prm.sort_depth_bits = 0; prm.sort_mode = HGS_SORT_GLOBAL;
    hgs_raster_inputs in{fptr(bg), fptr(m3), fptr(shc), fptr(col), fptr(op), fptr(sc), fptr(rot), fptr(cov), fptr(vm), fptr(pm), fptr(cp)};
    const int rendered = hgs_rasterize_forward(resize_cb, &geom, resize_cb, &binning, resize_cb, &img, &prm, &in,
                                               out_color.data_ptr<float>(), radii.data_ptr<int>(),
                                               at::cuda::getCurrentCUDAStream().stream());
    check(rendered, "hgs_rasterize_forward");
    return std::make_tuple(rendered, out_color, radii, geom, binning, img);
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansBackwardCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
                               const torch::Tensor& colors, const torch::Tensor& scales, const torch::Tensor& rotations,
                               const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                               const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                               const torch::Tensor& dL_dout_color, const torch::Tensor& sh, const int degree,
                               const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
                               const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer, const bool debug) {
    const c10::cuda::CUDAGuard guard(means3D.device());
    const int P = (int)means3D.size(0), C = (int)dL_dout_color.size(0), H = (int)dL_dout_color.size(1), W = (int)dL_dout_color.size(2);
    const int M = sh.numel() ? (int)sh.size(1) : 0;
    auto f32 = means3D.options().dtype(torch::kFloat32);
    // no zero fill needed: the library clears what the compositor accumulates into and writes everything else once
    auto acc = torch::empty({(long long)P * (3 + 4 + 1 + C)}, f32);
    auto dL_dmeans2D = acc.narrow(0, 0, 3LL * P).view({P, 3});
    auto dL_dconic = acc.narrow(0, 3LL * P, 4LL * P).view({P, 4});
    auto dL_dopacity = acc.narrow(0, 7LL * P, P).view({P, 1});
    auto dL_dcolors = acc.narrow(0, 8LL * P, (long long)P * C).view({P, C});
    auto dL_drotations = torch::empty({P, 4}, f32);
    auto dL_dmeans3D = torch::empty({P, 3}, f32);
    auto dL_dcov3D = torch::empty({P, 6}, f32);
    auto dL_dscales = torch::empty({P, 3}, f32);
    auto dL_dsh = torch::empty({P, M, 3}, f32);
    if (P != 0) {
        const auto bg = contig(background), m3 = contig(means3D), col = contig(colors), sc = contig(scales), rot = contig(rotations),
                   cov = contig(cov3D_precomp), vm = contig(viewmatrix), pm = contig(projmatrix), shc = contig(sh), cp = contig(campos),
                   dpix = contig(dL_dout_color);
        hgs_raster_params prm{};
        prm.P = P; prm.D = degree; prm.M = M; prm.width = W; prm.height = H; prm.channels = C;
        prm.tan_fovx = tan_fovx; prm.tan_fovy = tan_fovy; prm.scale_modifier = scale_modifier; prm.debug = debug;
        hgs_raster_inputs in{fptr(bg), fptr(m3), fptr(shc), fptr(col), nullptr, fptr(sc), fptr(rot), fptr(cov), fptr(vm), fptr(pm), fptr(cp)};
        hgs_raster_grads gr{dL_dmeans2D.data_ptr<float>(), dL_dconic.data_ptr<float>(), dL_dopacity.data_ptr<float>(),
                            dL_dcolors.data_ptr<float>(), dL_dmeans3D.data_ptr<float>(), dL_dcov3D.data_ptr<float>(),
                            M ? dL_dsh.data_ptr<float>() : nullptr, dL_dscales.data_ptr<float>(), dL_drotations.data_ptr<float>()};
        const int64_t cap = hgs_binning_capacity((size_t)binningBuffer.numel(), C, R);
        check((int)(cap < 0 ? cap : 0), "hgs_binning_capacity");
        check(hgs_rasterize_backward(&prm, &in, cap, radii.data_ptr<int>(), geomBuffer.data_ptr(),
                                     binningBuffer.numel() ? binningBuffer.data_ptr() : nullptr, imageBuffer.data_ptr(),
                                     dpix.data_ptr<float>(), &gr, at::cuda::getCurrentCUDAStream().stream()),
              "hgs_rasterize_backward");
    }
    return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations);
}

torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix) {
    const c10::cuda::CUDAGuard guard(means3D.device());
    const int P = (int)means3D.size(0);
    auto present = torch::zeros({P}, means3D.options().dtype(torch::kBool));
    if (P != 0) {
        const auto m3 = means3D.contiguous(), vm = viewmatrix.contiguous(), pm = projmatrix.contiguous();
        check(hgs_mark_visible(P, m3.data_ptr<float>(), vm.data_ptr<float>(), pm.data_ptr<float>(),
                               reinterpret_cast<uint8_t*>(present.data_ptr<bool>()), at::cuda::getCurrentCUDAStream().stream()),
              "hgs_mark_visible");
    }
    return present;
}

torch::Tensor distCUDA2(const torch::Tensor& points) {  // simple-knn/spatial.cu:15-26
    const c10::cuda::CUDAGuard guard(points.device());
    const int P = (int)points.size(0);
    auto means = torch::zeros({P}, points.options().dtype(torch::kFloat32));
    if (P != 0) {
        const auto pts = points.contiguous();
        auto ws = torch::empty({(long long)hgs_knn_bytes(P)}, torch::TensorOptions(torch::kByte).device(points.device()));
        check(hgs_dist2_knn3(P, pts.data_ptr<float>(), means.data_ptr<float>(), ws.data_ptr(), at::cuda::getCurrentCUDAStream().stream()),
              "hgs_dist2_knn3");
    }
    return means;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("rasterize_gaussians", &RasterizeGaussiansCUDA);
    m.def("rasterize_gaussians_backward", &RasterizeGaussiansBackwardCUDA);
    m.def("mark_visible", &markVisible);
    m.def("distCUDA2", &distCUDA2);
}
