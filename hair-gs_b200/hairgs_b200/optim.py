"""Optimiser step and densification statistics on the flat bucket.

FlatAdam is the subset of torch.optim.Adam that Hair-GS uses (scene/gaussian_model.py:216-250: one parameter group per
named tensor, per-group lr rewritten every iteration by update_learning_rate, betas (0.9, 0.999), eps 1e-15), laid out
for the data-parallel step: parameters, gradients and both moments live in four flat fp32 tensors; autograd accumulates
straight into the gradient bucket (multiview.GradBucket), the bucket is all-reduced once and ONE kernel
(hgs_adam_step) updates every group and clears the gradient.  There is no CPU path.
"""
import ctypes
from collections import OrderedDict

import torch

from . import _lib as L
from .multiview import GradBucket


class FlatAdam:
    def __init__(self, param_groups, betas=(0.9, 0.999), eps=1e-15):
        """param_groups: [{"params": [tensor], "lr": float, "name": str}, ...] as built in training_setup()."""
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.param_groups = []
        shapes = OrderedDict()
        dev = None
        for g in param_groups:
            ps = [p for p in g["params"]]
            for k, p in enumerate(ps):
                if not p.is_cuda or p.dtype != torch.float32:
                    raise L.HgsError("FlatAdam: parameters must be float32 CUDA tensors (no CPU path)")
                dev = p.device
                shapes[f'{g["name"]}.{k}'] = tuple(p.shape)
            self.param_groups.append({"params": ps, "lr": float(g.get("lr", 0.0)), "name": g["name"]})
        if len(self.param_groups) > 16:
            raise L.HgsError("FlatAdam: at most 16 parameter groups (HGS_ADAM_MAX_GROUPS)")
        self.device = dev
        self.grads = GradBucket(shapes, dev)
        n = self.grads.flat.numel()
        self.flat_param = torch.empty(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step_count = 0
        ends = []
        with torch.no_grad():
            for g in self.param_groups:
                for k, p in enumerate(g["params"]):
                    o, numel, shape = self.grads.slices[f'{g["name"]}.{k}']
                    view = self.flat_param[o:o + numel].view(shape)
                    view.copy_(p.data)
                    p.data = view                                     # the parameter now lives in the flat buffer
                    p.grad = self.grads.view(f'{g["name"]}.{k}')      # and its gradient in the bucket
                    end = o + numel
                ends.append(end if g["params"] else (ends[-1] if ends else 0))
        self._ends = (ctypes.c_int64 * len(ends))(*ends)
        self._lrs = (ctypes.c_float * len(ends))()

    def attach_grads(self):
        """Re-point p.grad at the bucket (after zero_grad(set_to_none=True)-style code dropped it)."""
        for g in self.param_groups:
            for k, p in enumerate(g["params"]):
                p.grad = self.grads.view(f'{g["name"]}.{k}')

    def all_reduce(self, group=None):
        self.grads.all_reduce(group=group)
        return self

    def step(self, grad_scale=1.0, zero_grad=True):
        lib = L.load()
        self.step_count += 1
        for i, g in enumerate(self.param_groups):
            self._lrs[i] = float(g["lr"])
        n = self.flat_param.numel()
        with torch.cuda.device(self.device):
            L.check(lib.hgs_adam_step(n, self.flat_param.data_ptr(), self.grads.flat.data_ptr(), self.exp_avg.data_ptr(),
                                      self.exp_avg_sq.data_ptr(), len(self.param_groups), self._ends, self._lrs,
                                      self.step_count, self.betas[0], self.betas[1], self.eps, float(grad_scale),
                                      1 if zero_grad else 0, L.stream_ptr(self.device)), "adam_step")

    def zero_grad(self, set_to_none=False):
        """Gradients are views of one bucket that step() already cleared; kept for call-site compatibility
        (train.py:204).  Clears explicitly when called without a preceding step()."""
        self.grads.zero_()
        self.attach_grads()

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "lr": [g["lr"] for g in self.param_groups], "names": [g["name"] for g in self.param_groups]}

    def load_state_dict(self, sd):
        if sd["names"] != [g["name"] for g in self.param_groups] or sd["exp_avg"].numel() != self.exp_avg.numel():
            raise L.HgsError("FlatAdam.load_state_dict: parameter groups do not match")
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, lr in zip(self.param_groups, sd["lr"]):
            g["lr"] = float(lr)


class DensifyStats:
    """max_radii2D / xyz_gradient_accum / denom of scene/gaussian_model.py:208-211, updated by one kernel per view
    (update_densification_stats, :675-682)."""

    def __init__(self, P, device, distributed=False):
        """distributed=True (view-sharded training): update() accumulates this rank's views into LOCAL
        buffers and reduce() folds them into the public tensors with (sum, sum, max) across ranks — called on
        densification iterations only, so the per-view path stays collective-free."""
        self.max_radii2D = torch.zeros(P, dtype=torch.float32, device=device)
        self.xyz_gradient_accum = torch.zeros(P, 1, dtype=torch.float32, device=device)
        self.denom = torch.zeros(P, 1, dtype=torch.float32, device=device)
        self.distributed = bool(distributed)
        if self.distributed:
            self.local_max_radii2D = torch.zeros_like(self.max_radii2D)
            self.local_xyz_gradient_accum = torch.zeros_like(self.xyz_gradient_accum)
            self.local_denom = torch.zeros_like(self.denom)

    def reduce(self, group=None):
        """Fold the views every rank has seen since the last call into max_radii2D / xyz_gradient_accum / denom:
        all-reduce SUM of the two accumulators, MAX of the radii (the nonlinear reductions of
        scene/gaussian_model.py:675-682 commute with this split: max of maxima, sum of per-view norms, sum of counts)."""
        import torch.distributed as dist
        if not self.distributed:
            return self
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.local_xyz_gradient_accum, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.local_denom, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.local_max_radii2D, op=dist.ReduceOp.MAX, group=group)
        self.xyz_gradient_accum.add_(self.local_xyz_gradient_accum)
        self.denom.add_(self.local_denom)
        torch.maximum(self.max_radii2D, self.local_max_radii2D, out=self.max_radii2D)
        self.local_xyz_gradient_accum.zero_()
        self.local_denom.zero_()
        self.local_max_radii2D.zero_()
        return self

    def update(self, viewspace_point_grad, radii):
        lib = L.load()
        dev = self.max_radii2D.device
        P = self.max_radii2D.shape[0]
        if not viewspace_point_grad.is_cuda:
            raise L.HgsError("DensifyStats.update: CUDA tensors only (no CPU path)")
        g = L.f32c(viewspace_point_grad, "viewspace_point_grad", dev)
        if g.dim() != 2 or g.shape[0] != P or g.shape[1] < 2 or radii.shape[0] != P:
            raise L.HgsError("DensifyStats.update: expected grad [P,>=2] and radii [P]")
        r = radii if (radii.dtype == torch.int32 and radii.is_contiguous()) else radii.to(torch.int32).contiguous()
        with torch.cuda.device(dev):
            mr, acc, den = ((self.local_max_radii2D, self.local_xyz_gradient_accum, self.local_denom) if self.distributed
                            else (self.max_radii2D, self.xyz_gradient_accum, self.denom))
            L.check(lib.hgs_densify_stats(P, r.data_ptr(), g.data_ptr(), g.shape[1], mr.data_ptr(), acc.data_ptr(),
                                          den.data_ptr(), L.stream_ptr(dev)), "densify_stats")
