"""Strand endpoint merge search: device kernels vs the cKDTree restatement in oracle/merge_oracle.py
(scene/hair_gaussian_model.py:1205-1362).  Index outputs must be identical; distances bit-equal."""
import numpy as np
import pytest
import torch

from oracle import merge_oracle


def _scene(S, V, seed, extent):
    """S strands of V joints: ids s*V .. s*V+V-1; the strand ends are the first and the last joint."""
    rng = np.random.default_rng(seed)
    ends = np.stack([np.arange(S) * V, np.arange(S) * V + V - 1], 1).reshape(-1)          # global ids of the 2S ends
    other = np.stack([np.arange(S) * V + V - 1, np.arange(S) * V], 1).reshape(-1)
    points = rng.uniform(0, extent, (2 * S, 3)).astype(np.float32)
    dirs = rng.normal(size=(2 * S, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    other_end_of = -np.ones(S * V, dtype=np.int64)
    other_end_of[ends] = other
    return points, dirs, ends.astype(np.int64), other.astype(np.int64), other_end_of


def test_oracle_on_a_hand_case():
    """Two tips facing each other 1 mm apart merge; a third end nearby pointing the wrong way does not."""
    points = np.array([[0, 0, 0], [0, 0, 1], [0.001, 0, 0], [0.001, 0, 1], [0.0005, 0.0005, 0], [0.5, 0.5, 1]], np.float32)
    dirs = np.array([[-1, 0, 0], [0, 0, -1], [1, 0, 0], [0, 0, -1], [0, 1, 0], [0, 0, -1]], np.float32)
    gid = np.array([0, 9, 10, 19, 20, 29])
    other = np.array([9, 0, 19, 10, 29, 20])
    p1, p2, d = merge_oracle.merge_candidates(points, dirs, gid, other, 2e-3, 20.0)
    assert sorted(zip(p1.tolist(), p2.tolist())) == [(0, 10), (10, 0)]
    other_end_of = -np.ones(30, dtype=np.int64)
    other_end_of[gid] = other
    keep = merge_oracle.greedy_filter(p1, p2, other_end_of)
    assert keep.tolist() == [True, False]      # the mirrored row repeats both ids


@pytest.mark.gpu
@pytest.mark.parametrize("S,extent,bidir,max_nn", [(300, 0.02, False, -1), (2000, 0.05, True, -1), (1500, 0.03, False, 2),
                                                   (1, 0.01, False, -1)])
def test_merge_search_matches_ckdtree_reference(S, extent, bidir, max_nn):
    from hairgs_b200 import merge
    dev = torch.device("cuda:0")
    points, dirs, gid, other, other_end_of = _scene(S, 10, S + 7, extent)
    dist_th, angle = 4e-3, 40.0
    rp1, rp2, rd = merge_oracle.merge_candidates(points, dirs, gid, other, dist_th, angle, bidir, max_nn)
    t = lambda a: torch.from_numpy(a).to(dev)
    p1, p2, d = merge.merge_candidates(t(points), t(dirs), t(gid), t(other), dist_th, angle, bidir, max_nn)
    assert p1.shape[0] == rp1.shape[0]
    assert np.array_equal(p1.cpu().numpy(), rp1) and np.array_equal(p2.cpu().numpy(), rp2)
    assert np.array_equal(d.cpu().numpy().view(np.uint32), rd.view(np.uint32))
    if S > 1:
        assert rp1.shape[0] > 20          # the case exercises the filters
    # full routine: stable distance sort + duplicate / complementary filters
    order = np.argsort(rd, kind="stable")
    keep = merge_oracle.greedy_filter(rp1[order], rp2[order], other_end_of)
    ref_pairs = np.stack([rp1[order][keep], rp2[order][keep]], 1)
    pairs = merge.endpoint_pairs_to_merge(t(points), t(dirs), t(gid), t(other), t(other_end_of), dist_th, angle, bidir, max_nn)
    assert pairs.dtype == torch.int64 and np.array_equal(pairs.cpu().numpy(), ref_pairs.reshape(-1, 2))
    # one-to-one: no endpoint id appears twice, and never both ends of one strand
    flat = pairs.cpu().numpy().reshape(-1)
    assert len(set(flat.tolist())) == flat.shape[0]


@pytest.mark.gpu
def test_merge_search_rejects_cpu_tensors():
    from hairgs_b200 import merge
    points, dirs, gid, other, _ = _scene(4, 10, 1, 0.01)
    with pytest.raises(Exception):
        merge.merge_candidates(torch.from_numpy(points), torch.from_numpy(dirs), torch.from_numpy(gid), torch.from_numpy(other),
                               1e-3, 20.0)
