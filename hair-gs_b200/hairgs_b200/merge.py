"""Strand endpoint merge search on the device (second half).

Replaces the body of HairGaussianModel.compute_endpoint_pair_to_merge (scene/hair_gaussian_model.py:1205-1362): scipy's
cKDTree ball query on the host, the per-point Python filter loop and the per-row Python matching loop become three kernel
launches (hgs_merge_count / hgs_merge_fill / hgs_merge_greedy) plus a scan and a sort of the (small) candidate list.
No CPU path.
"""
import math

import torch

from . import _lib as L


def _i32(t, name, dev):
    if not t.is_cuda:
        raise L.HgsError(f"merge search: {name} must be a CUDA tensor (no CPU path)")
    return t.to(device=dev, dtype=torch.int32).contiguous()


def merge_candidates(points, dirs, global_id, other_end, dist_th, angle_th_deg, bidirectional=False, max_num_nn=-1):
    """All candidate pairs in the reference's order (by strand end, then by neighbour index): (p1, p2, dist).

    points/dirs: [K,3] float32 — positions of the strand ends (roots/tips) and unit directions towards their neighbouring
    joint; global_id: [K] endpoint ids; other_end: [K] endpoint id of the other end of the same strand."""
    lib = L.load()
    if not points.is_cuda:
        raise L.HgsError("merge search: points must be a CUDA tensor (no CPU path)")
    dev = points.device
    pts = L.f32c(points, "points", dev)
    drs = L.f32c(dirs, "dirs", dev)
    K = pts.shape[0]
    if pts.shape != (K, 3) or drs.shape != (K, 3) or global_id.shape[0] != K or other_end.shape[0] != K:
        raise L.HgsError("merge search: points/dirs must be [K,3], global_id/other_end [K]")
    gid, oth = _i32(global_id, "global_id", dev), _i32(other_end, "other_end", dev)
    dir_th = math.cos(math.radians(float(angle_th_deg)))   # np.cos(np.deg2rad(angle_th)), :1259
    args = (K, pts.data_ptr(), drs.data_ptr(), gid.data_ptr(), oth.data_ptr(), float(dist_th), dir_th,
            1 if bidirectional else 0, int(max_num_nn))
    counts = torch.zeros(K, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.hgs_merge_count(*args, counts.data_ptr(), L.stream_ptr(dev)), "merge_count")
        incl = torch.cumsum(counts, 0, dtype=torch.int64)
        total = int(incl[-1]) if K else 0      # host sync: a topology edit is host-synchronous in the reference too
        offsets = (incl - counts).contiguous()
        p1 = torch.empty(total, dtype=torch.int32, device=dev)
        p2 = torch.empty(total, dtype=torch.int32, device=dev)
        dist = torch.empty(total, dtype=torch.float32, device=dev)
        if total:
            L.check(lib.hgs_merge_fill(*args, offsets.data_ptr(), p1.data_ptr(), p2.data_ptr(), dist.data_ptr(),
                                       L.stream_ptr(dev)), "merge_fill")
    return p1, p2, dist


def endpoint_pairs_to_merge(points, dirs, global_id, other_end, other_end_of, dist_th, angle_th_deg, bidirectional=False,
                            max_num_nn=-1):
    """index_pairs_to_merge [n,2] (int64) as compute_endpoint_pair_to_merge returns it.  other_end_of: int tensor over
    ALL endpoint ids (strands_info.strand_endpoint_id_to_complementary), -1 where an id is not a strand end."""
    lib = L.load()
    p1, p2, dist = merge_candidates(points, dirs, global_id, other_end, dist_th, angle_th_deg, bidirectional, max_num_nn)
    dev = p1.device
    n = p1.shape[0]
    if n == 0:
        return torch.empty((0, 2), dtype=torch.int64, device=dev)
    order = torch.sort(dist, stable=True).indices    # the reference's torch.sort is unstable; ties keep search order here
    p1s, p2s = p1[order].contiguous(), p2[order].contiguous()
    comp = _i32(other_end_of, "other_end_of", dev)
    flags = torch.zeros(comp.shape[0], dtype=torch.uint8, device=dev)
    keep = torch.empty(n, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.hgs_merge_greedy(n, p1s.data_ptr(), p2s.data_ptr(), comp.data_ptr(), flags.data_ptr(), keep.data_ptr(),
                                     L.stream_ptr(dev)), "merge_greedy")
    m = keep.bool()
    return torch.stack([p1s[m], p2s[m]], 1).to(torch.int64)
