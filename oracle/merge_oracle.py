"""TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of HairGaussianModel.compute_endpoint_pair_to_merge
(scene/hair_gaussian_model.py:1205-1362) with the reference's own third-party dependency, scipy.spatial.cKDTree
(query_ball_point, return_sorted=True), numpy float32 arithmetic and the two sequential filters
(remove_duplicate_endpoint_rows :711-726, remove_complementary_rows :1237-1255).  The strand bookkeeping around it
(which endpoints are strand ends, foreground mask) is the caller's input here exactly as for the device entry points.
Only tests/ may import this module."""
import numpy as np
from scipy.spatial import cKDTree


def merge_candidates(points, dirs, global_id, other_end, dist_th, angle_th_deg, bidirectional=False, max_num_nn=-1):
    points = np.asarray(points, dtype=np.float32)
    dirs = np.asarray(dirs, dtype=np.float32)
    dir_th = np.cos(np.deg2rad(angle_th_deg))                        # :1259
    tree = cKDTree(points)                                            # :1283
    nns = tree.query_ball_point(points, r=dist_th, return_sorted=True)  # :1295-1300
    p1, p2, dist = [], [], []
    for i in range(points.shape[0]):                                  # :1302-1335
        nn = np.array(nns[i], dtype=np.int64)
        g = global_id[nn]
        nn = nn[np.logical_and(g != other_end[i], g != global_id[i])]
        if len(nn) == 0:
            continue
        dot = dirs[nn] @ (-dirs[i]).T
        if bidirectional:
            dot = np.abs(dot)
        nn = nn[dot >= dir_th]
        d = np.linalg.norm(points[i] - points[nn], axis=1)
        k = len(nn) if max_num_nn <= 0 else min(max_num_nn, len(nn))
        for j in range(k):
            p1.append(global_id[i]); p2.append(global_id[nn[j]]); dist.append(d[j])
    return np.array(p1, dtype=np.int64), np.array(p2, dtype=np.int64), np.array(dist, dtype=np.float32)


def greedy_filter(p1, p2, other_end_of):
    """Rows already sorted by distance.  Returns the boolean keep mask."""
    n = len(p1)
    flat = np.stack([p1, p2], 1).reshape(-1)
    first = np.zeros(2 * n, dtype=bool)
    seen = set()
    for k, v in enumerate(flat):                                      # get_first_occurence_index :773-784
        if v not in seen:
            seen.add(v); first[k] = True
    keep = np.logical_and(first[0::2], first[1::2])                   # :720-722
    disabled = set()
    for r in range(n):                                                # :1246-1253 over the surviving rows
        if not keep[r]:
            continue
        if p1[r] in disabled or p2[r] in disabled:
            keep[r] = False
        else:
            disabled.add(int(other_end_of[p1[r]])); disabled.add(int(other_end_of[p2[r]]))
    return keep
