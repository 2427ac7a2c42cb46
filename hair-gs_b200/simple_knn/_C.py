"""`simple_knn._C.distCUDA2` over the C ABI (replaces submodules/simple-knn/spatial.cu:15-26)."""
import torch

from hairgs_b200 import _lib as L


def distCUDA2(points):
    """Mean squared distance to the 3 nearest neighbours of each point: f32[P,3] cuda -> f32[P]."""
    lib = L.load()
    if not points.is_cuda:
        raise L.HgsError("points must be a CUDA tensor: distCUDA2 has no CPU path")
    dev = points.device
    P = points.size(0)
    means = torch.zeros((P,), dtype=torch.float32, device=dev)
    if P == 0:
        return means
    pts = L.f32c(points, "points", dev)
    with torch.cuda.device(dev):
        ws = torch.empty((lib.hgs_knn_bytes(P),), dtype=torch.uint8, device=dev)
        L.check(lib.hgs_dist2_knn3(P, pts.data_ptr(), means.data_ptr(), ws.data_ptr(), L.stream_ptr(dev)),
                "distCUDA2")
    return means
