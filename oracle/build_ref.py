"""Build recipe for oracle/_ref: the *unmodified* reference rasterizer and simple-knn,
compiled for sm_100a straight from the sources where they lie under /root/reference.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import oracle/.
Only build artefacts (ninja files, objects, the two .so) are written, and only into
oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).  No reference source
is copied into this repository.

Sources compiled (reference file list, cf. submodules/diff-gaussian-rasterization/setup.py:21-32
and submodules/simple-knn/setup.py:21-31):
  DGR: cuda_rasterizer/{rasterizer_impl,forward,backward_distwar}.cu, rasterize_points.cu, ext.cpp
  KNN: spatial.cu, simple_knn.cu, ext.cpp
Flags: torch defaults (-O3, no fast-math) + -gencode arch=compute_100a,code=sm_100a, as the
reference's setup.py passes no arch flag of its own.
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("HAIRGS_REFERENCE", "/root/reference")
DGR = os.path.join(REF, "submodules", "diff-gaussian-rasterization")
KNN = os.path.join(REF, "submodules", "simple-knn")


def build(verbose=False):
    if not os.path.isdir(DGR):
        print(f"[oracle/_ref] reference not present at {REF}; using prebuilt files if any")
        return False
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    # This image's default CXX (/opt/gcc wrapper) links libstdc++ STATICALLY into shared objects; the
    # reference prints integers through std::cout in BACKWARD::render (backward_distwar.cu:1203-1205) and a
    # private, uninitialised libstdc++ copy segfaults there.  Use the distro compiler (dynamic libstdc++).
    if os.path.exists("/usr/bin/g++"):
        os.environ["CXX"] = "/usr/bin/g++"
        os.environ["CC"] = "/usr/bin/gcc"
    from torch.utils.cpp_extension import load
    ok = True
    for name, srcs, inc in (
        ("ref_dgr_C",
         [f"{DGR}/cuda_rasterizer/rasterizer_impl.cu", f"{DGR}/cuda_rasterizer/forward.cu",
          f"{DGR}/cuda_rasterizer/backward_distwar.cu", f"{DGR}/rasterize_points.cu", f"{DGR}/ext.cpp"],
         [f"{DGR}/third_party/glm", DGR]),
        ("ref_knn_C", [f"{KNN}/spatial.cu", f"{KNN}/simple_knn.cu", f"{KNN}/ext.cpp"], []),
    ):
        bd = os.path.join(OUT, name)
        so = os.path.join(bd, name + ".so")
        if os.path.exists(so):
            print(f"[oracle/_ref] {so} already built")
            continue
        os.makedirs(bd, exist_ok=True)
        t = time.time()
        try:
            load(name=name, sources=srcs, extra_include_paths=inc, build_directory=bd,
                 extra_cuda_cflags=["-lineinfo"], verbose=verbose, is_python_module=False)
            print(f"[oracle/_ref] built {name} in {time.time() - t:.0f}s")
        except Exception as e:  # noqa
            ok = False
            print(f"[oracle/_ref] FAILED {name}: {e}")
    return ok


if __name__ == "__main__":
    sys.exit(0 if build(verbose="-v" in sys.argv) else 1)
