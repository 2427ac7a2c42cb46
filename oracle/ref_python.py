"""TEST INFRASTRUCTURE ONLY — the reference's own PYTHON side of the render path, unmodified, as a parity oracle and
as the `bench.py --impl reference` arm.  Never imported by anything under hair-gs_b200/.

/root/reference does not exist on the GPU box and reference sources must not be copied into this repository, so —
exactly like the CUDA sources that oracle/build_ref.py compiles into oracle/_ref/*.so — the reference's Python
modules are COMPILED where they lie (compile() -> marshalled code objects, one `<module>.hgsref` file each) into
oracle/_ref/pyref/ (git-ignored, shipped to the box) and imported through a private meta-path finder.  Modules compiled
(reference paths):

    submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py   autograd surface (A2-A4)
    gaussian_renderer/__init__.py                                                    render() (A1)
    scene/gaussian_model.py, scene/hair_gaussian_model.py, scene/cameras.py          getters (A21/A22), Camera (A24)
    loss/losses.py                                                                   loss_function (N3)
    utils/{system,transform,sh,graphics,general,logging}.py                          torch helpers (A23, A24)
    arguments/__init__.py                                                            OptimizationParams defaults

What is NOT the reference's (absent third-party packages, stubbed at import):
    pytorch3d.transforms.matrix_to_quaternion   restated (hairgs_b200/scenes.py; un-vendored, un-pinned dependency,
                                                SURVEY.md 8c) — the one arithmetic stub on the path
    pytorch3d.ops.knn_points, plyfile, c_utils, wandb, tensorboard   never called on the render path: inert stubs
    diff_gaussian_rasterization._C / simple_knn._C                   = oracle/_ref/ref_dgr_C.so / ref_knn_C.so
    scene/__init__.py, utils/__init__.py, loss/__init__.py           not executed (they import dataset readers,
                                                pyvista, ...); the packages are namespace stubs over the compiled files

load() leaves sys.modules as it found it (the reference's package names collide with the drop-in's), so the product
packages and the reference can live in one test process.
"""
import importlib
import importlib.abc
import importlib.util
import marshal
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PYREF = os.path.join(HERE, "_ref", "pyref")
REF = os.environ.get("HAIRGS_REFERENCE", "/root/reference")

EXT = ".hgsref"
MAGIC = b"HGSREF" + bytes(sys.version_info[:2]) + importlib.util.MAGIC_NUMBER
MODULES = {
    "diff_gaussian_rasterization/__init__": "submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py",
    "gaussian_renderer/__init__": "gaussian_renderer/__init__.py",
    "scene/gaussian_model": "scene/gaussian_model.py",
    "scene/hair_gaussian_model": "scene/hair_gaussian_model.py",
    "scene/cameras": "scene/cameras.py",
    "loss/losses": "loss/losses.py",
    "utils/system": "utils/system.py",
    "utils/transform": "utils/transform.py",
    "utils/sh": "utils/sh.py",
    "utils/graphics": "utils/graphics.py",
    "utils/general": "utils/general.py",
    "utils/logging": "utils/logging.py",
    "arguments/__init__": "arguments/__init__.py",
}
_UTILS_SUBMODULES = ("system", "transform", "sh", "graphics", "general", "logging")
_COLLIDING = ("diff_gaussian_rasterization", "gaussian_renderer", "simple_knn", "scene", "utils", "loss", "arguments",
              "pytorch3d", "plyfile", "c_utils", "wandb")


def build(verbose=False):
    """Compile the reference's Python modules into oracle/_ref/pyref/ (only where the reference is present)."""
    if not os.path.isdir(REF):
        print(f"[oracle/_ref/pyref] reference not present at {REF}; using prebuilt byte code if any")
        return available()
    for dst, src in MODULES.items():
        out = os.path.join(PYREF, dst + EXT)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        srcp = os.path.join(REF, src)
        if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(srcp):
            continue
        with open(srcp, "rb") as fh:
            code = compile(fh.read(), f"<reference>/{src}", "exec", dont_inherit=True, optimize=0)
        with open(out, "wb") as fh:
            fh.write(MAGIC + marshal.dumps(code))
        if verbose:
            print(f"[oracle/_ref/pyref] {src} -> {os.path.relpath(out, ROOT)}")
    return True


def available():
    return all(os.path.exists(os.path.join(PYREF, d + EXT)) for d in MODULES)


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Imports the compiled reference modules by name while load() runs (first on sys.meta_path, removed afterwards)."""

    def __init__(self):
        self.table = {}
        for dst in MODULES:
            name = dst.replace("/", ".")
            pkg = name.endswith(".__init__")
            self.table[name[:-len(".__init__")] if pkg else name] = (os.path.join(PYREF, dst + EXT), pkg)

    def find_spec(self, fullname, path=None, target=None):
        if fullname not in self.table:
            return None
        file, pkg = self.table[fullname]
        spec = importlib.util.spec_from_loader(fullname, self, origin=file, is_package=pkg)
        if pkg:
            spec.submodule_search_locations = [os.path.dirname(file)]
        return spec

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        file, _ = self.table[module.__name__]
        with open(file, "rb") as fh:
            blob = fh.read()
        if blob[:len(MAGIC)] != MAGIC:
            raise ImportError(f"{file}: compiled by another Python ({sys.version_info[:2]} needed); rebuild oracle/_ref/pyref")
        exec(marshal.loads(blob[len(MAGIC):]), module.__dict__)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def _unavailable(what):
    def f(*a, **k):
        raise NotImplementedError(f"{what} is not installed in this image and is not on the render path")
    return f


def _load_file(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _ref_ext(name):
    so = os.path.join(HERE, "_ref", name, name + ".so")
    if not os.path.exists(so):
        return None
    return _load_file(name, so)


_cache = {}


def load(with_cuda_ext=True):
    """-> namespace with the reference's modules: .dgr (diff_gaussian_rasterization), .gaussian_renderer,
    .gaussian_model, .hair_gaussian_model, .cameras, .losses, .utils (namespace of the compiled utils submodules),
    .transform, .sh, .graphics, .general, .arguments.  with_cuda_ext=False: the CUDA extension modules are inert stubs
    (CPU-only pinning tests of the torch-side helpers)."""
    key = bool(with_cuda_ext)
    if key in _cache:
        return _cache[key]
    if not available():
        raise ImportError("oracle/_ref/pyref is not built (run oracle/ref_python.py where /root/reference exists)")
    import torch  # noqa: F401
    # the one arithmetic stub: pytorch3d.transforms.matrix_to_quaternion, restated in hairgs_b200/scenes.py (pure torch;
    # loaded by file path so that the product package and its shared library are NOT imported by the reference arm)
    scenes = _load_file("_hgs_scenes_for_ref", os.path.join(ROOT, "hair-gs_b200", "hairgs_b200", "scenes.py"))
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _COLLIDING}
    for k in saved:
        del sys.modules[k]
    finder = _Finder()
    sys.meta_path.insert(0, finder)
    try:
        os.environ.setdefault("BW_IMPLEMENTATION", "1")   # train.py:278 -> utils/general.py:119-124
        os.environ.setdefault("BALANCE_THRESHOLD", "8")
        dgr_c = _ref_ext("ref_dgr_C") if with_cuda_ext else None
        knn_c = _ref_ext("ref_knn_C") if with_cuda_ext else None
        sys.modules["pytorch3d"] = _stub("pytorch3d")
        sys.modules["pytorch3d.transforms"] = _stub("pytorch3d.transforms", matrix_to_quaternion=scenes.matrix_to_quaternion)
        sys.modules["pytorch3d.ops"] = _stub("pytorch3d.ops", knn_points=_unavailable("pytorch3d.ops.knn_points"))
        sys.modules["pytorch3d"].transforms = sys.modules["pytorch3d.transforms"]
        sys.modules["pytorch3d"].ops = sys.modules["pytorch3d.ops"]
        sys.modules["plyfile"] = _stub("plyfile", PlyData=type("PlyData", (), {}), PlyElement=type("PlyElement", (), {}))
        sys.modules["c_utils"] = _stub("c_utils", filter_strand_list_segments=_unavailable("c_utils"))
        sys.modules["wandb"] = _stub("wandb")
        tb_saved = sys.modules.get("torch.utils.tensorboard")
        sys.modules["torch.utils.tensorboard"] = _stub("torch.utils.tensorboard", SummaryWriter=type("SummaryWriter", (), {}))
        sys.modules["simple_knn"] = _stub("simple_knn")
        sys.modules["simple_knn._C"] = knn_c if knn_c is not None else _stub("simple_knn._C", distCUDA2=_unavailable("simple_knn"))
        if dgr_c is not None:
            sys.modules["diff_gaussian_rasterization._C"] = dgr_c
        else:
            sys.modules["diff_gaussian_rasterization._C"] = _stub("diff_gaussian_rasterization._C")
        for pkg in ("scene", "utils", "loss"):
            m = _stub(pkg)
            m.__path__ = [os.path.join(PYREF, pkg)]
            sys.modules[pkg] = m
        utils = sys.modules["utils"]
        for sub in _UTILS_SUBMODULES:      # what utils/__init__.py star-imports, minus the viewer / dataset modules
            mod = importlib.import_module("utils." + sub)
            for k, v in vars(mod).items():
                if not k.startswith("_"):
                    setattr(utils, k, v)
        ns = types.SimpleNamespace()
        ns.utils = utils
        ns.transform, ns.sh, ns.graphics, ns.general = (sys.modules["utils." + s] for s in ("transform", "sh", "graphics", "general"))
        ns.dgr = importlib.import_module("diff_gaussian_rasterization")
        ns.dgr_C, ns.knn_C = dgr_c, knn_c
        ns.gaussian_model = importlib.import_module("scene.gaussian_model")
        ns.hair_gaussian_model = importlib.import_module("scene.hair_gaussian_model")
        ns.cameras = importlib.import_module("scene.cameras")
        ns.gaussian_renderer = importlib.import_module("gaussian_renderer")
        # two signatures in loss/losses.py build a default `bg` tensor on "cuda" at import time (:228,:296)
        real_tensor = torch.tensor

        def tensor_any(*a, **k):
            if str(k.get("device", "")).startswith("cuda") and not torch.cuda.is_available():
                k["device"] = "cpu"
            return real_tensor(*a, **k)
        torch.tensor = tensor_any
        try:
            ns.losses = importlib.import_module("loss.losses")
        finally:
            torch.tensor = real_tensor
        ns.arguments = importlib.import_module("arguments")
        ns.matrix_to_quaternion_stub = scenes.matrix_to_quaternion
    finally:
        sys.meta_path.remove(finder)
        for k in [k for k in sys.modules if k.split(".")[0] in _COLLIDING]:
            del sys.modules[k]
        sys.modules.update(saved)
        if "tb_saved" in locals():
            if tb_saved is not None:
                sys.modules["torch.utils.tensorboard"] = tb_saved
            else:
                sys.modules.pop("torch.utils.tensorboard", None)
    _cache[key] = ns
    return ns


class cuda_as_cpu:
    """Context for the GPU-less build container: the reference hard-codes device="cuda" in its torch helpers
    (utils/transform.py:14,34, utils/general.py:72); inside this context those factory calls land on the CPU so the helpers
    can be run to record golden vectors.  Does nothing when a CUDA device exists."""
    NAMES = ("zeros", "ones", "empty", "tensor", "eye", "zeros_like", "ones_like", "full")

    def __enter__(self):
        import torch
        self.torch, self.saved = torch, {}
        if torch.cuda.is_available():
            return self
        for n in self.NAMES:
            real = getattr(torch, n)
            self.saved[n] = real

            def wrap(*a, __real=real, **k):
                if str(k.get("device", "")).startswith("cuda"):
                    k["device"] = "cpu"
                return __real(*a, **k)
            setattr(torch, n, wrap)
        return self

    def __exit__(self, *exc):
        for n, real in self.saved.items():
            setattr(self.torch, n, real)
        return False


if __name__ == "__main__":
    ok = build(verbose=True)
    print("[oracle/_ref/pyref]", "ok" if ok else "unavailable")
    sys.exit(0 if ok else 1)
