"""N>1 host logic on CPU: two gloo ranks shard views round-robin, accumulate per-view gradients (produced by the
CPU oracle — the CUDA rasterizer needs a GPU) into the flat bucket and all-reduce once; the result must equal the
single-process sum of per-view gradients ('parity for the batch against the sum of per-view oracle
gradients')."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common
from hairgs_b200 import multiview, scenes

N_VIEWS = 3
NAMES = ("dL_dmeans3D", "dL_dscales", "dL_drotations", "dL_dopacity", "dL_dsh")


def _view_grads(view):
    from oracle import pyoracle
    d = common.blob_inputs(400, 64, 48, "cpu", seed=5, view=view, n_views=N_VIEWS + 1, sh_degree=1, M=4)
    f = pyoracle.Forward({k: (v.numpy() if torch.is_tensor(v) else v) for k, v in d.items()})
    dL = np.random.default_rng(view).standard_normal(f.color.shape).astype(np.float32)
    g = f.backward(dL)
    f.close()
    return {k: torch.tensor(g[k]) for k in NAMES}


def _shapes():
    return {"dL_dmeans3D": (400, 3), "dL_dscales": (400, 3), "dL_drotations": (400, 4), "dL_dopacity": (400, 1),
            "dL_dsh": (400, 4, 3)}


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bucket = multiview.GradBucket(_shapes(), "cpu")
        views, weights = multiview.shard_views(N_VIEWS, world, rank)
        for v, w in zip(views, weights):
            for k, g in _view_grads(v).items():
                bucket.accumulate(k, g, w)
        bucket.all_reduce()
        if rank == 0:
            torch.save(bucket.flat.clone(), out_path)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4])
def test_two_rank_bucket_allreduce_equals_sum_of_views(tmp_path, world):
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = torch.load(out)
    ref = multiview.GradBucket(_shapes(), "cpu")
    for v in range(N_VIEWS):
        for k, g in _view_grads(v).items():
            ref.accumulate(k, g)
    assert torch.allclose(got, ref.flat, rtol=1e-5, atol=1e-6)


def test_shard_views_round_robin():
    for n in (1, 3, 16, 64):
        for w in (1, 2, 4, 8):
            seen = []
            for r in range(w):
                v, wt = multiview.shard_views(n, w, r)
                assert len(v) >= 1 and len(v) == len(wt)
                seen += [x for x, y in zip(v, wt) if y > 0]
            assert sorted(seen) == list(range(n))
            assert multiview.steps_per_epoch(n, w) == -(-n // w)
    with pytest.raises(ValueError):
        multiview.shard_views(4, 2, 2)


def test_bucket_attach_accumulates_autograd_in_place():
    params = {"a": torch.nn.Parameter(torch.randn(5, 3)), "b": torch.nn.Parameter(torch.randn(7))}
    bucket = multiview.GradBucket({k: p.shape for k, p in params.items()}, "cpu")
    bucket.attach_to(params)
    (params["a"].sum() * 2 + (params["b"] ** 2).sum()).backward()
    assert torch.allclose(bucket.view("a"), torch.full((5, 3), 2.0))
    assert torch.allclose(bucket.view("b"), 2 * params["b"].detach())
    assert bucket.flat.data_ptr() == params["a"].grad.data_ptr()


def test_shard_views_cost_sorted():
    """Views of similar cost meet in the same lock-step iteration; every view is dealt exactly once."""
    costs = [5, 1, 9, 3, 7, 2, 8, 4]
    world = 4
    dealt = [multiview.shard_views(len(costs), world, r, costs=costs)[0] for r in range(world)]
    assert sorted(v for d in dealt for v in d) == list(range(len(costs)))
    step0 = sorted(costs[d[0]] for d in dealt)
    step1 = sorted(costs[d[1]] for d in dealt)
    assert step0 == [5, 7, 8, 9] and step1 == [1, 2, 3, 4]
    # the second group is dealt in reverse rank order: the per-rank sums (the cost of a multi-view step) are balanced
    sums = [sum(costs[v] for v in d) for d in dealt]
    assert sums == [10, 10, 10, 9]
    for n, w in ((64, 8), (10, 4), (3, 8)):
        c = [((7 * i) % 11) + 1 for i in range(n)]
        parts = [multiview.shard_views(n, w, r, costs=c)[0] for r in range(w)]
        assert sorted(set(v for d in parts for v in d)) == list(range(n))
    with pytest.raises(ValueError):
        multiview.shard_views(3, 2, 0, costs=[1, 2])


def _stats_worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hairgs_b200 import optim
        P = 50
        st = optim.DensifyStats(P, "cpu", distributed=True)
        red = multiview.AsyncReducer(P * 3, "cpu")
        flat = torch.zeros(P * 3)
        for rnd in range(2):                              # two densification intervals
            for view in range(rank, 6, world):            # this rank's views of the interval
                g = torch.Generator().manual_seed(100 * rnd + view)
                radii = torch.randint(0, 30, (P,), generator=g).float() * (torch.rand(P, generator=g) < 0.7)
                grad = torch.randn(P, 3, generator=g)
                vis = radii > 0
                # what hgs_densify_stats does on the device (scene/gaussian_model.py:675-682), into the LOCAL buffers
                st.local_max_radii2D[vis] = torch.max(st.local_max_radii2D[vis], radii[vis])
                st.local_xyz_gradient_accum[vis] += torch.norm(grad[vis, :2], dim=-1, keepdim=True)
                st.local_denom[vis] += 1
                flat += grad.reshape(-1)
            st.reduce()
        summed = red.launch(flat).wait().clone()
        if rank == 0:
            torch.save((st.max_radii2D, st.xyz_gradient_accum, st.denom, summed), out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_densify_stats_and_async_reducer_across_ranks(tmp_path, world):
    """densification statistics are per-view nonlinear reductions -> accumulated locally, folded with
    (sum, sum, max) across ranks on densification iterations; the gradient bucket goes through the side-stream reducer.
    Both must equal the single-process result over all views."""
    out = str(tmp_path / "stats.pt")
    mp.spawn(_stats_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    mr, acc, den, summed = torch.load(out)
    P = 50
    rmr, racc, rden, rflat = torch.zeros(P), torch.zeros(P, 1), torch.zeros(P, 1), torch.zeros(P * 3)
    for rnd in range(2):
        for view in range(6):
            g = torch.Generator().manual_seed(100 * rnd + view)
            radii = torch.randint(0, 30, (P,), generator=g).float() * (torch.rand(P, generator=g) < 0.7)
            grad = torch.randn(P, 3, generator=g)
            vis = radii > 0
            rmr[vis] = torch.max(rmr[vis], radii[vis])
            racc[vis] += torch.norm(grad[vis, :2], dim=-1, keepdim=True)
            rden[vis] += 1
            rflat += grad.reshape(-1)
    assert torch.equal(mr, rmr) and torch.equal(den, rden)
    assert torch.allclose(acc, racc, rtol=1e-6, atol=1e-6)
    assert torch.allclose(summed, rflat, rtol=1e-5, atol=1e-6)
