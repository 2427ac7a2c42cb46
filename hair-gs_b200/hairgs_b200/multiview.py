"""View-sharded data parallelism for the render path.

The path shards naturally over camera views: every rank holds a full replica of the Gaussian parameters, renders
the views dealt to it, accumulates their parameter gradients IN PLACE into one flat fp32 bucket and the bucket is
summed across ranks with a single all-reduce per optimiser step (NCCL over NVLink on the B200 box, gloo in the CPU
tests).  There is no other data-path collective.  The reference has no distributed code at all
(utils/general.py:116 pins cuda:0), so this is new surface, kept deliberately small.
"""
from collections import OrderedDict

import torch
import torch.distributed as dist


def shard_views(n_views, world_size, rank, costs=None):
    """View i goes to rank i mod world_size (round-robin).  Every rank gets at least one view so that
    collectives stay matched when n_views < world_size (the extra ranks repeat a view with zero weight).

    costs (optional, one number per view, identical on every rank — e.g. the view's tile-instance count N): the views
    are first ordered by decreasing cost and then dealt in groups of world_size, so the views that meet in one lock-step
    iteration cost about the same; every other group is dealt in reverse rank order (snake), so the SUM over a rank's views
    — what a multi-view step costs — is balanced too (plain round-robin hands rank 0 the dearest view of every group)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    order = list(range(n_views))
    if costs is not None:
        if len(costs) != n_views:
            raise ValueError("one cost per view")
        order.sort(key=lambda v: (-float(costs[v]), v))
        mine = []
        for g in range(0, n_views, world_size):
            i = g + (rank if (g // world_size) % 2 == 0 else world_size - 1 - rank)
            if i < n_views:
                mine.append(order[i])
    else:
        mine = order[rank::world_size]
    weights = [1.0] * len(mine)
    if not mine:
        mine, weights = [rank % max(n_views, 1)], [0.0]
    return mine, weights


def steps_per_epoch(n_views, world_size):
    """Number of lock-step iterations needed to cover all views: ceil(n_views / world_size)."""
    return (n_views + world_size - 1) // world_size


class GradBucket:
    """One flat fp32 tensor with named views; gradients are accumulated into the views in place and the whole
    bucket is all-reduced at once."""

    def __init__(self, shapes, device):
        self.slices = OrderedDict()
        n = 0
        for name, shape in shapes.items():
            numel = 1
            for s in shape:
                numel *= int(s)
            self.slices[name] = (n, numel, tuple(shape))
            n += numel
        self.flat = torch.zeros(n, dtype=torch.float32, device=device)

    def view(self, name):
        o, n, shape = self.slices[name]
        return self.flat[o:o + n].view(shape)

    def zero_(self):
        self.flat.zero_()
        return self

    def accumulate(self, name, grad, weight=1.0):
        v = self.view(name)
        if weight == 1.0:
            v.add_(grad.reshape(v.shape))
        elif weight != 0.0:
            v.add_(grad.reshape(v.shape), alpha=weight)
        return self

    def attach_to(self, params):
        """Point p.grad of every parameter at its slice so autograd accumulates straight into the bucket."""
        for name, p in params.items():
            p.grad = self.view(name)

    def all_reduce(self, group=None, average_over=None):
        """Sum over ranks (one collective for every parameter); optionally divide by the number of views."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average_over:
            self.flat.div_(float(average_over))
        return self


class AsyncReducer:
    """All-reduce of a flat gradient bucket OFF the critical path ("on a side stream, overlapped with the
    first view of the next batch").  launch(flat) snapshots the bucket into a staging buffer (one device copy on the
    caller's stream) and sums the snapshot across ranks on a side stream; the caller's stream goes straight on to the
    next batch, which may overwrite the bucket at once.  wait() makes the caller's stream wait for the sum, which is then
    in `.staging` (what the optimiser consumes).  On CPU tensors (gloo tests) it degrades to a blocking all-reduce."""

    def __init__(self, numel, device, group=None):
        self.staging = torch.zeros(int(numel), dtype=torch.float32, device=device)
        self.group = group
        self.cuda = torch.device(device).type == "cuda"
        self.stream = torch.cuda.Stream(device=device) if self.cuda else None
        self.done = None
        self.launched = 0
        self.timing = False      # record (begin, end) CUDA events around every collective on the side stream
        self._timed = []

    def _active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def launch(self, flat):
        if flat.numel() != self.staging.numel():
            raise ValueError("AsyncReducer: bucket size changed")
        self.launched += 1
        if not self.cuda:
            self.staging.copy_(flat)
            if self._active():
                dist.all_reduce(self.staging, op=dist.ReduceOp.SUM, group=self.group)
            return self
        dev = self.staging.device
        cur = torch.cuda.current_stream(dev)
        if self.done is not None:
            cur.wait_event(self.done)          # the previous reduction has finished with the staging buffer
        self.staging.copy_(flat, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            begin = None
            if self.timing:
                begin = torch.cuda.Event(enable_timing=True)
                begin.record(self.stream)
            if self._active():
                dist.all_reduce(self.staging, op=dist.ReduceOp.SUM, group=self.group)
            self.done = torch.cuda.Event(enable_timing=self.timing)
            self.done.record(self.stream)
            if begin is not None:
                self._timed.append((begin, self.done))
        return self

    def collective_ms(self):
        """Device time of the recorded collectives (timing=True), one number per launch; synchronises the side stream.
        A collective starts when the LAST rank arrives, so this includes the wait for the slowest rank of the step."""
        if not self._timed:
            return []
        self.stream.synchronize()
        out = [a.elapsed_time(b) for a, b in self._timed]
        self._timed = []
        return out

    def wait(self):
        """The caller's stream waits for the last launched reduction; returns the staging buffer holding the sum."""
        if self.cuda and self.done is not None:
            torch.cuda.current_stream(self.staging.device).wait_event(self.done)
        return self.staging
