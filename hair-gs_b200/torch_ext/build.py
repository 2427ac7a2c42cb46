"""Builds hgs_torch_ext (hair-gs_b200/torch_ext/hgs_torch_ext.cpp: the compiled torch binding over the C ABI of
libhairgs_rast.so) in-tree with torch.utils.cpp_extension.  No CUDA sources: host C++ only, linked against the library.

    python hair-gs_b200/torch_ext/build.py      -> hair-gs_b200/lib/hgs_torch_ext.so
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIBDIR = os.path.join(PKG, "lib")
NAME = "hgs_torch_ext"


def build(verbose=False):
    sys.path.insert(0, PKG)
    import importlib.util
    spec = importlib.util.spec_from_file_location("hgs_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()                                          # the library it links against
    so = os.path.join(LIBDIR, NAME + ".so")
    src = os.path.join(HERE, NAME + ".cpp")
    if os.path.exists(so) and os.path.getmtime(so) >= max(os.path.getmtime(src), os.path.getmtime(os.path.join(PKG, "..", "include", "hairgs_rast.h"))):
        return so
    if os.path.exists("/usr/bin/g++"):
        os.environ.setdefault("CXX", "/usr/bin/g++")
    import ctypes
    ctypes.CDLL(os.path.join(LIBDIR, "libhairgs_rast.so"), mode=ctypes.RTLD_GLOBAL)   # the extension links against it
    from torch.utils.cpp_extension import load
    bd = os.path.join(PKG, "build", NAME)
    os.makedirs(bd, exist_ok=True)
    load(name=NAME, sources=[src], build_directory=bd, with_cuda=True, verbose=verbose, is_python_module=False,
         extra_ldflags=[f"-L{LIBDIR}", "-lhairgs_rast", f"-Wl,-rpath,{LIBDIR}"], extra_cflags=["-O2"])
    os.replace(os.path.join(bd, NAME + ".so"), so)
    return so


def load_module():
    import ctypes
    import importlib.util
    import torch  # noqa: F401
    so = build()
    ctypes.CDLL(os.path.join(LIBDIR, "libhairgs_rast.so"), mode=ctypes.RTLD_GLOBAL)
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
