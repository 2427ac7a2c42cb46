"""CUDA-graph replay of one Hair-GS Stage-III training view (SURVEY.md §8f N1-N3 in one graph).

The eager step — `fused.render_strands` -> `losses.hair_image_loss` -> `loss.backward()` — is ~25 kernel launches and
~0.9 ms of Python / autograd bookkeeping per view; on cfg3 that is as long as the device work itself (0.94 ms), so the
eager loop is host-bound (profiles/r1_e2e_host_time.txt).  The forward was built without host synchronisation for
exactly this purpose: with a fixed *launch plan* (instance capacity and sort depth bits decided up front) nothing the
host does depends on the view, and the whole view — strand parameterisation, binning, sort, both compositors, the image
loss and every gradient down to the raw parameters — is captured once and replayed with one `cudaGraphLaunch`.

What varies per view lives in device memory the caller fills between replays: the camera (world_view_transform,
full_proj_transform, camera_center: 35 floats) and the six target planes.  The reference has no counterpart (its forward
blocks on `num_rendered` mid-pass, rasterizer_impl.cu:281, which cannot be captured).

Validation is deferred, never skipped: every replay copies (num_rendered, overflow, depth range) to pinned memory; the
next use of the slot checks it against the plan, and `validate()` checks every slot used so far.  A training loop calls
`validate()` after the last view of a step and BEFORE the all-reduce / optimiser step (`HgsPlanError` if a view did not fit:
its image and gradients came from a truncated instance list; the parameters are still untouched, so the caller re-plans —
`measure_plan` — re-captures and redoes the step).  Widths and positions train, so a long run re-measures the plan
periodically (e.g. with every topology edit, which needs a re-capture anyway).
"""
import math

import torch

from . import _lib as L
from . import fused, losses
from diff_gaussian_rasterization import _C as _dgr


class HgsPlanError(L.HgsError):
    """A replayed view needed more instances or more depth bits than the captured launch plan provides."""


class LaunchPlan:
    """Fixed launch plan of the strand forward: capacity of the binning workspace (instances) and the sort's depth bits.
    `host` is the pinned read-back target of hgs_forward_read_num_rendered."""

    def __init__(self, capacity, depth_bits, sort_mode=None, slices=(0, 0)):
        self.capacity, self.depth_bits = int(capacity), int(depth_bits)
        self.slice_base, self.slice_shift = int(slices[0]), int(slices[1])   # depth-slice hints of the in-tile sort
        # binning formulation of the captured view: in-tile sort unless a view of this scene had a tile list longer than
        # HGS_TILE_SORT_MAX (diff_gaussian_rasterization._C._sort_mode_hint, filled by the eager passes of measure_plan)
        self.sort_mode = int(_dgr.DEFAULT_SORT_MODE if sort_mode is None else sort_mode)
        self.host = torch.zeros(8, dtype=torch.int32).pin_memory()

    def check(self):
        """Call once the work that used this plan has completed.  Returns the view's instance count."""
        N, overflow = int(self.host[0]), int(self.host[2])
        if (overflow & 1) != 0 or N < 0:
            raise HgsPlanError("instance count overflows int32")
        if N > self.capacity:
            raise HgsPlanError(f"view has {N} tile instances, the captured plan holds {self.capacity}")
        if self.sort_mode == L.SORT_TILE:
            if overflow & 4:
                raise HgsPlanError("a tile list is longer than HGS_TILE_SORT_MAX: the captured plan sorts inside the tiles; "
                                   "re-plan (measure_plan switches the scene to the global sort)")
            return N
        need = _dgr._depth_range_bits(self.host)
        if self.depth_bits not in (0, 32) and need > self.depth_bits:
            raise HgsPlanError(f"view needs {need} depth bits in the sort keys, the captured plan compares {self.depth_bits}")
        return N


def measure_plan(model, cameras, bg7):
    """(capacity, depth_bits) that fit every camera in `cameras` for the current parameters: one eager forward per camera
    (no grad).  The eager path keeps, per (device, P, H, W), the largest instance count seen plus 25 % head-room and the
    widest depth range rounded up to whole sort passes (diff_gaussian_rasterization._C hints); the plan is those."""
    H, W = int(cameras[0].image_height), int(cameras[0].image_width)
    key = (model._endpoints.device.index, int(model.endpoint_pairs.shape[0]), H, W, 7)
    if not _dgr.SYNC_FREE:
        raise L.HgsError("measure_plan needs the capacity hints of the sync-free forward (diff_gaussian_rasterization._C."
                         "SYNC_FREE is off)")
    if len(cameras) == 0:
        raise L.HgsError("measure_plan: no cameras")
    with torch.no_grad():
        for cam in cameras:
            if int(cam.image_height) != H or int(cam.image_width) != W:
                raise L.HgsError("measure_plan: all cameras of one plan must have the same resolution")
            fused.render_strands(cam, model, bg7)
    return int(_dgr._capacity_hint[key]), int(_dgr._depth_bits_hint[key])


class GraphedStrandStep:
    """One training view as a CUDA graph: render_strands + hair_image_loss + backward into a fused.GradSink.

        step = GraphedStrandStep(model, sink, bg7, H, W, fovx, fovy, capacity, depth_bits, lambdas=dict(...))
        step.cam_buf[s].copy_(...); step.tgt_buf[s].copy_(...)      # fill every slot once with a real view
        step.capture()
        loop:  copy the view's camera (35 floats: world_view_transform | full_proj_transform | camera_center) and targets
               (image[3] | mask | orientation field | confidence) into slot s, then
               loss = step.replay(s)       # device scalar; parameter gradients are in the sink's tensors,
                                           # step.mean2d_grad[s] / step.radii[s] feed the densification statistics

    `slots` static input buffers (default 2) let the next view's host->device copies overlap the current replay; each
    slot has its own graph (same kernels, different input addresses) sharing one memory pool.
    """

    def __init__(self, model, sink, bg7, H, W, fovx, fovy, capacity, depth_bits, lambdas=None, slots=2, cam_buf=None,
                 tgt_buf=None, dimage=None, accumulate_variant=True):
        dev = model._endpoints.device
        if dev.type != "cuda":
            raise L.HgsError("GraphedStrandStep needs a CUDA model: this rasterizer has no CPU path")
        if sink is None or not all(k in sink.tensors for k in ("endpoints", "width", "opacity", "mask", "features")):
            raise L.HgsError("GraphedStrandStep needs a fused.GradSink with endpoints / width / opacity / mask / features "
                             "tensors: the captured backward writes the parameter gradients there")
        if int(capacity) <= 0:
            raise L.HgsError("GraphedStrandStep: capacity must be positive (see measure_plan)")
        self.model, self.sink, self.bg7, self.dev = model, sink, bg7, dev
        self.H, self.W, self.fovx, self.fovy = int(H), int(W), float(fovx), float(fovy)
        self.lambdas = dict(lambdas or {})
        # dimage: a static [7,H,W] dL/dimage7 to back-propagate INSTEAD of the image loss (the caller's own loss gradient,
        # or a fixed one for measurements); the target slots are then unused
        if dimage is not None and (dimage.shape != (7, self.H, self.W) or dimage.dtype != torch.float32
                                   or dimage.device != dev or not dimage.is_contiguous()):
            raise L.HgsError("GraphedStrandStep: dimage must be a contiguous float32 [7,H,W] tensor on the model's device")
        self.dimage = dimage
        # also capture the view with the backward ADDING to the sink (gradient accumulation over the views of a batch)
        self.accumulate_variant = bool(accumulate_variant)
        # static inputs: the caller's own staging buffers (cam_buf / tgt_buf, one per slot) or fresh ones
        self.cam_buf = list(cam_buf) if cam_buf is not None else [torch.zeros(35, device=dev) for _ in range(slots)]
        slots = len(self.cam_buf)
        if tgt_buf is not None:
            self.tgt_buf = list(tgt_buf)
        elif dimage is not None:
            self.tgt_buf = [torch.zeros(6, self.H, self.W, device=dev)] * slots     # never read
        else:
            self.tgt_buf = [torch.zeros(6, self.H, self.W, device=dev) for _ in range(slots)]
        for c, t in zip(self.cam_buf, self.tgt_buf):
            if c.shape != (35,) or t.shape != (6, self.H, self.W) or c.device != dev or t.device != dev \
                    or c.dtype != torch.float32 or t.dtype != torch.float32 or not t.is_contiguous():
                raise L.HgsError("GraphedStrandStep: slot buffers must be float32 [35] and contiguous [6,H,W] on the model's "
                                 "device")
        key = (dev.index, int(model.endpoint_pairs.shape[0]), self.H, self.W, 7)
        self.plans = [LaunchPlan(capacity, depth_bits, _dgr.sort_mode_for(key), _dgr.slice_params(key)) for _ in range(slots)]
        self.done = [None] * slots          # event after the slot's last replay
        self.graphs, self.loss, self.terms = [], [], []
        self.mean2d_grad, self.radii, self.image = [None] * slots, [None] * slots, [None] * slots  # static outputs
        self.replays = 0

    # the body that is captured.  It calls the forward, the loss and the backward directly (no autograd engine inside the
    # capture: the engine synchronises with the streams on which the parameters' AccumulateGrad nodes were created, which
    # are outside the capture whenever an eager autograd graph of the same parameters is still alive).
    def _view(self, slot, accumulate=False):
        cd, tgt = self.cam_buf[slot], self.tgt_buf[slot]
        wvt = cd[0:16].view(4, 4)
        m = self.model
        lam = self.lambdas
        l_dssim = float(lam.get("lambda_dssim", 0.2))
        # defaults: the reference's OptimizationParams (arguments/__init__.py:84-86)
        weights = (max(0.0, 1.0 - l_dssim), l_dssim, float(lam.get("lambda_mask", 0.01)),
                   float(lam.get("lambda_orientation", 100.0)))
        settings = dict(image_height=self.H, image_width=self.W, tanfovx=math.tan(self.fovx * 0.5),
                        tanfovy=math.tan(self.fovy * 0.5), bg=self.bg7, scale_modifier=1.0, viewmatrix=wvt,
                        projmatrix=cd[16:32].view(4, 4), sh_degree=m.active_sh_degree, campos=cd[32:35], debug=False,
                        grad_sink=self.sink, plan=self.plans[slot])
        with torch.no_grad():
            args = (m._endpoints, m.endpoint_pairs, m._width, m._opacity, m._mask, m.get_features)
            image, radii, state = fused.strands_forward(*args, settings)
            if self.dimage is not None:
                terms, dimage = torch.zeros(8, device=self.dev), self.dimage
            else:
                terms, dimage = losses.hair_image_loss_raw(image, tgt[0:3], tgt[3], tgt[4], tgt[5], tgt[3] > 0.5, wvt,
                                                           weights, lam.get("bg_orient", (0.0, 0.0, 0.0)))
            # first view of an optimiser step: the backward OVERWRITES the sink; later views of the step add to it
            self.sink.accumulate = bool(accumulate)
            grads = fused.strands_backward(*args, settings, state, dimage)
        self.mean2d_grad[slot] = grads[5]      # screen-space mean gradients (densification statistics)
        self.radii[slot] = radii
        self.image[slot] = image
        return terms[0], terms

    def capture(self, warmup=2):
        """Runs `warmup` eager views per slot on a side stream (lazy initialisation inside the library and autograd must not
        happen during capture), then captures one graph per slot.  The slots must already hold a real view."""
        cur = torch.cuda.current_stream(self.dev)
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for slot in range(len(self.cam_buf)):
                for _ in range(warmup):
                    self._view(slot)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        for p in self.plans:
            p.check()
        pool = None
        for slot in range(len(self.cam_buf)):
            per_mode = []
            for accumulate in ((False, True) if self.accumulate_variant else (False,)):
                g = torch.cuda.CUDAGraph()
                # thread_local: calls of OTHER host threads (e.g. the NCCL watchdog of a multi-GPU job) must not
                # invalidate the capture; the captured body itself runs on this thread only (no autograd worker threads)
                with torch.cuda.graph(g, pool=pool, capture_error_mode="thread_local"):
                    loss, terms = self._view(slot, accumulate)
                pool = g.pool()
                per_mode.append((g, loss, terms, self.mean2d_grad[slot], self.radii[slot], self.image[slot]))
            self.graphs.append(per_mode)
            self.loss.append(per_mode[0][1])
            self.terms.append(per_mode[0][2])
        return self

    def replay(self, slot, accumulate=False):
        """Replays the view in `slot` on the current stream; returns the (static) device scalar of the loss.
        accumulate=False: first view of an optimiser step, the parameter gradients in the sink are overwritten;
        accumulate=True: a later view of the same step, its gradients are added (needs accumulate_variant=True).
        Raises HgsPlanError if the PREVIOUS replay of this slot did not fit the plan."""
        ev = self.done[slot]
        if ev is not None:
            ev.synchronize()
            self.plans[slot].check()
        if accumulate and not self.accumulate_variant:
            raise L.HgsError("GraphedStrandStep was built without the accumulate variant (accumulate_variant=True)")
        g, loss, terms, mean2d, radii, image = self.graphs[slot][1 if accumulate else 0]
        self.loss[slot], self.terms[slot] = loss, terms
        self.mean2d_grad[slot], self.radii[slot], self.image[slot] = mean2d, radii, image
        g.replay()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self.done[slot] = ev
        self.replays += 1
        return self.loss[slot]

    def validate(self):
        """Call BEFORE anything irreversible consumes the gradients of the views replayed so far (the all-reduce and the
        optimiser step mutate the parameters and their moments in place): waits for the last replay of every slot — one
        event wait each, the read-back words are already in pinned memory — and raises HgsPlanError if any view did not
        fit the captured plan, while the step can still be redone with a new plan.  Returns the instance counts."""
        out = []
        for p, ev in zip(self.plans, self.done):
            if ev is not None:
                ev.synchronize()
                out.append(p.check())
        return out

    def check(self):
        """Synchronises the device and validates the last replay of every slot."""
        torch.cuda.synchronize(self.dev)
        return [p.check() for p, ev in zip(self.plans, self.done) if ev is not None]


class GraphedStrandBatch:
    """The V views of one optimiser step as ONE CUDA graph with two branches, so that the binning of view k+1 runs UNDER the
    compositing of view k (SURVEY.md 8e batch schedule):

        binning stream (high priority):  [stage A + scan + keys + sort + ranges/packing](0) -> (1) -> ... -> (V-1)
        main stream:                      wait bin(0): [composite fwd + image loss + composite bwd + preprocess bwd](0) -> ...

    The binning chain of a cfg3 view is ~0.23 ms of one-wave, latency-bound kernels at <= 30 % issue utilisation; the
    compositors are issue-bound.  Neither can use the SM alone, together they can: the binning kernels take the issue slots
    and the CTA slots the compositors leave idle.  Every view has its own workspaces (geometry, binning, image: ~0.5 GB per
    cfg3 view), so the binning branch never waits for the main branch; the main branch is ordered (the backward of view k
    adds to the gradients view k-1 wrote), which is the only dependency between views.  The first view OVERWRITES the
    gradient sink (or adds, `replay(accumulate=True)`), the others add: one replay = the gradient of the whole batch.

    Inputs that change per step live in static device buffers the caller fills before replay(): cam_buf [V,35]
    (world_view_transform | full_proj_transform | camera_center per view) and tgt_buf [V,6,H,W] (image[3] | mask |
    orientation field | confidence); or a fixed dL/dimage7 (`dimage`) instead of the image loss.  Outputs: losses [V],
    terms [V,8], per-view image / radii / mean2d_grad.  validate() / check() as for GraphedStrandStep."""

    def __init__(self, model, sink, bg7, H, W, fovx, fovy, capacity, depth_bits, views, lambdas=None, dimage=None,
                 cam_buf=None, tgt_buf=None):
        dev = model._endpoints.device
        if dev.type != "cuda":
            raise L.HgsError("GraphedStrandBatch needs a CUDA model: this rasterizer has no CPU path")
        if sink is None or not all(k in sink.tensors for k in ("endpoints", "width", "opacity", "mask", "features")):
            raise L.HgsError("GraphedStrandBatch needs a fused.GradSink with endpoints / width / opacity / mask / features")
        if int(capacity) <= 0 or int(views) < 1:
            raise L.HgsError("GraphedStrandBatch: capacity and views must be positive (see measure_plan)")
        self.lib = L.load()
        self.model, self.sink, self.dev = model, sink, dev
        self.bg7 = L.f32c(bg7, "bg7", dev)
        self.H, self.W, self.V = int(H), int(W), int(views)
        self.tanfovx, self.tanfovy = math.tan(float(fovx) * 0.5), math.tan(float(fovy) * 0.5)
        self.lambdas = dict(lambdas or {})
        if dimage is not None and (dimage.shape != (7, self.H, self.W) or dimage.dtype != torch.float32
                                   or dimage.device != dev or not dimage.is_contiguous()):
            raise L.HgsError("GraphedStrandBatch: dimage must be a contiguous float32 [7,H,W] tensor on the model's device")
        self.dimage = dimage
        V, P = self.V, int(model.endpoint_pairs.shape[0])
        self.P = P
        f32 = dict(dtype=torch.float32, device=dev)
        self.cam_buf = cam_buf if cam_buf is not None else torch.zeros(V, 35, **f32)
        self.tgt_buf = tgt_buf if tgt_buf is not None else (torch.zeros(V, 6, self.H, self.W, **f32) if dimage is None else None)
        if self.cam_buf.shape != (V, 35) or self.cam_buf.dtype != torch.float32 or not self.cam_buf.is_contiguous():
            raise L.HgsError("GraphedStrandBatch: cam_buf must be a contiguous float32 [V,35] tensor")
        if self.tgt_buf is not None and (self.tgt_buf.shape != (V, 6, self.H, self.W) or not self.tgt_buf.is_contiguous()):
            raise L.HgsError("GraphedStrandBatch: tgt_buf must be a contiguous float32 [V,6,H,W] tensor")
        key = (dev.index, int(model.endpoint_pairs.shape[0]), self.H, self.W, 7)
        self.plans = [LaunchPlan(capacity, depth_bits, _dgr.sort_mode_for(key), _dgr.slice_params(key)) for _ in range(V)]
        lib = self.lib
        u8 = dict(dtype=torch.uint8, device=dev)
        # per-view workspaces and outputs, allocated once (outside any capture)
        self.geom = [torch.empty(lib.hgs_geom_bytes(P, 7, self.W, self.H), **u8) for _ in range(V)]
        self.img_ws = [torch.empty(lib.hgs_image_bytes(self.W, self.H), **u8) for _ in range(V)]
        self.binning = [torch.empty(lib.hgs_binning_bytes(int(capacity), 7), **u8) for _ in range(V)]
        self.image = [torch.empty(7, self.H, self.W, **f32) for _ in range(V)]
        self.radii = [torch.empty(P, dtype=torch.int32, device=dev) for _ in range(V)]
        self.vec = fused.BWD_VECTOR_RED
        if self.vec:       # interleaved [P,16] accumulation records + [P,3] mean2D output (hgs_strand_grads.acc16)
            self.acc = [torch.empty(P * 19, **f32) for _ in range(V)]
            self.mean2d_grad = [a[16 * P:].view(P, 3) for a in self.acc]
        else:              # mean2D 3 | conic 4 | opacity 1 | colour 7
            self.acc = [torch.empty(P * 15, **f32) for _ in range(V)]
            self.mean2d_grad = [a[:3 * P].view(P, 3) for a in self.acc]
        self.losses = torch.zeros(V, **f32)
        self.terms = torch.zeros(V, 8, **f32)
        self.bin_stream = torch.cuda.Stream(device=dev, priority=-1)
        # HGS_BATCH_COMP_BRANCHES: streams the views' compositing is dealt over (see _batch); 1 = one compositing branch
        import os
        self.comp_branches = max(1, min(4, int(os.environ.get("HGS_BATCH_COMP_BRANCHES", "2")), self.V))
        self.comp_streams = [torch.cuda.Stream(device=dev) for _ in range(self.comp_branches - 1)]
        self._keep = []
        self.graphs, self.execs, self.use_priority = {}, {}, True
        self.done = None
        self.replays = 0

    def _prm_inp(self, v, features):
        m = self.model
        cd = self.cam_buf[v]
        prm = L.RasterParams(P=self.P, D=int(m.active_sh_degree), M=int(features.shape[1]), width=self.W, height=self.H,
                             channels=7, tan_fovx=self.tanfovx, tan_fovy=self.tanfovy, scale_modifier=1.0, prefiltered=0,
                             debug=0, sort_depth_bits=int(self.plans[v].depth_bits), sort_mode=int(self.plans[v].sort_mode),
                             slice_base=int(self.plans[v].slice_base), slice_shift=int(self.plans[v].slice_shift))
        inp = L.StrandInputs(num_endpoints=int(m._endpoints.shape[0]), background=self.bg7.data_ptr(),
                             endpoints=m._endpoints.data_ptr(), endpoint_pairs=m.endpoint_pairs.data_ptr(),
                             width=m._width.data_ptr(), opacity_logit=m._opacity.data_ptr(), mask_logit=m._mask.data_ptr(),
                             features=features.data_ptr(), viewmatrix=cd[0:16].data_ptr(), projmatrix=cd[16:32].data_ptr(),
                             cam_pos=cd[32:35].data_ptr())
        return prm, inp

    def _batch(self, accumulate_first):
        """The body that is captured (also run eagerly as warm-up): two branches, see the class docstring."""
        import ctypes
        lib, dev, V = self.lib, self.dev, self.V
        m = self.model
        for t in (m._endpoints, m._width, m._opacity, m._mask):
            if not t.is_contiguous() or t.dtype != torch.float32:
                raise L.HgsError("GraphedStrandBatch: parameters must be contiguous float32 tensors")
        main = torch.cuda.current_stream(dev)
        side = self.bin_stream
        cap = int(self.plans[0].capacity)
        lam = self.lambdas
        l_dssim = float(lam.get("lambda_dssim", 0.2))
        weights = (max(0.0, 1.0 - l_dssim), l_dssim, float(lam.get("lambda_mask", 0.01)),
                   float(lam.get("lambda_orientation", 100.0)))
        s = self.sink
        E = int(m._endpoints.shape[0])
        Mf = int(m._features_dc.shape[1] + m._features_rest.shape[1])
        for k, t in s.tensors.items():
            want = {"endpoints": 3 * E, "width": self.P, "opacity": self.P, "mask": self.P, "features": 3 * Mf * self.P}.get(k)
            if want is not None and (t.numel() != want or t.dtype != torch.float32 or t.device != dev or not t.is_contiguous()):
                raise L.HgsError("grad_sink: tensors must be contiguous float32 on the render device, shaped like the parameters")
        self._keep = []
        with torch.no_grad(), torch.cuda.device(dev):
            features = m.get_features.contiguous()      # cat(dc, rest): once per batch, shared by the views
            self._features = features
            import os
            parts = os.environ.get("HGS_BATCH_PARTS", "both")   # debug / measurement: "bin" or "comp" runs one branch only
            side.wait_stream(main)                      # fork
            bin_done = []
            for v in range(V):
                prm, inp = self._prm_inp(v, features)
                if parts == "comp":
                    continue
                with torch.cuda.stream(side):
                    st = side.cuda_stream
                    L.check(lib.hgs_strands_forward_stage_a(ctypes.byref(prm), ctypes.byref(inp), self.geom[v].data_ptr(),
                                                            self.radii[v].data_ptr(), st), "strands stage A")
                    L.check(lib.hgs_forward_read_num_rendered(self.geom[v].data_ptr(), self.P, self.plans[v].host.data_ptr(), st),
                            "read num_rendered")
                    L.check(lib.hgs_forward_stage_b_binning(ctypes.byref(prm), self.geom[v].data_ptr(), self.binning[v].data_ptr(),
                                                            self.img_ws[v].data_ptr(), cap, st), "binning")
                    ev = torch.cuda.Event()
                    ev.record(side)
                    bin_done.append(ev)
            # compositing branches: view v composites, takes its loss and back-propagates on stream v % B.  With B > 1 the
            # forward side of view v+1 runs next to the backward of view v and fills the tail of its grid (a compositor
            # launch is ~80 % busy: heavy tiles first, the last ones trail); the backwards themselves are chained - each adds
            # to the gradients the previous one wrote.
            B = self.comp_branches
            comp_streams = [main] + self.comp_streams[:B - 1]
            for cs in comp_streams[1:]:
                cs.wait_stream(main)
            bwd_done = []
            for v in range(V):
                if parts == "bin":
                    break
                prm, inp = self._prm_inp(v, features)
                cs = comp_streams[v % B]
                with torch.cuda.stream(cs):
                    if parts != "comp":
                        cs.wait_event(bin_done[v])
                    st = cs.cuda_stream
                    L.check(lib.hgs_forward_stage_b_composite(ctypes.byref(prm), self.bg7.data_ptr(), self.geom[v].data_ptr(),
                                                              self.binning[v].data_ptr(), self.img_ws[v].data_ptr(), cap,
                                                              self.image[v].data_ptr(), st), "composite")
                    if self.dimage is not None:
                        dimage = self.dimage
                    else:
                        tgt = self.tgt_buf[v]
                        terms, dimage = losses.hair_image_loss_raw(self.image[v], tgt[0:3], tgt[3], tgt[4], tgt[5], tgt[3] > 0.5,
                                                                   self.cam_buf[v][0:16].view(4, 4), weights,
                                                                   lam.get("bg_orient", (0.0, 0.0, 0.0)))
                        self.terms[v].copy_(terms)
                        self._keep.append((terms, dimage))     # alive until the join: other branches must not reuse them
                    acc = self.acc[v]
                    vec = self.vec
                    grads = L.StrandGrads(dL_dmean2D=self.mean2d_grad[v].data_ptr(), dL_dconic=None if vec else acc[3 * self.P:].data_ptr(),
                                          dL_dopacity=None if vec else acc[7 * self.P:].data_ptr(),
                                          dL_dcolor=None if vec else acc[8 * self.P:].data_ptr(), acc16=acc.data_ptr() if vec else None,
                                          dL_dendpoints=s.tensors["endpoints"].data_ptr(), dL_dwidth=s.tensors["width"].data_ptr(),
                                          dL_dopacity_logit=s.tensors["opacity"].data_ptr(),
                                          dL_dmask_logit=s.tensors["mask"].data_ptr(), dL_dfeatures=s.tensors["features"].data_ptr(),
                                          accumulate=1 if (accumulate_first or v > 0) else 0)
                    # part 1 (backward compositor: this view's scratch only) may overlap other views' backwards; part 2
                    # (preprocess backward: adds to the shared parameter gradients) is chained over the views
                    bargs = (ctypes.byref(prm), ctypes.byref(inp), cap, self.geom[v].data_ptr(), self.binning[v].data_ptr(),
                             self.img_ws[v].data_ptr(), dimage.data_ptr(), ctypes.byref(grads))
                    L.check(lib.hgs_strands_backward_parts(*bargs, 1, st), "strands backward (compositor)")
                    if B > 1 and v > 0:
                        cs.wait_event(bwd_done[v - 1])
                    L.check(lib.hgs_strands_backward_parts(*bargs, 2, st), "strands backward (preprocess)")
                    if B > 1:
                        ev = torch.cuda.Event()
                        ev.record(cs)
                        bwd_done.append(ev)
            for cs in comp_streams[1:]:
                main.wait_stream(cs)
            if self.dimage is None:
                self.losses.copy_(self.terms[:, 0])
            main.wait_stream(side)                      # join

    def capture(self, warmup=2, accumulate_variant=True):
        """`warmup` eager batches on a side stream, then one graph per mode (first view overwrites / adds).  cam_buf (and
        tgt_buf) must already hold real views."""
        cur = torch.cuda.current_stream(self.dev)
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._batch(False)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        for p in self.plans:
            p.check()
        import ctypes
        import os
        # Node priorities: torch instantiates its graphs without cudaGraphInstantiateFlagUseNodePriority, so every node would
        # run at the launch stream's priority and the binning branch would queue behind the compositors' blocks.  The
        # captured cudaGraph_t is therefore kept (keep_graph=True) and instantiated by the library with the flag
        # (hgs_graph_instantiate); HGS_GRAPH_PRIORITY=0 falls back to torch's own instantiation (A/B measurements).
        self.use_priority = os.environ.get("HGS_GRAPH_PRIORITY", "1") != "0"
        pool = None
        for acc in ((False, True) if accumulate_variant else (False,)):
            g = torch.cuda.CUDAGraph(keep_graph=True) if self.use_priority else torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool, capture_error_mode="thread_local"):
                self._batch(acc)
            pool = g.pool()
            self.graphs[acc] = g
            if self.use_priority:
                ex = ctypes.c_void_p()
                L.check(self.lib.hgs_graph_instantiate(ctypes.c_void_p(g.raw_cuda_graph()), 1, ctypes.byref(ex)), "graph instantiate")
                self.execs[acc] = ex
        return self

    def replay(self, accumulate=False):
        """Replays the whole batch on the current stream; returns the static [V] tensor of the views' losses."""
        if self.done is not None:
            self.done.synchronize()
            for p in self.plans:
                p.check()
        if accumulate not in self.graphs:
            raise L.HgsError("GraphedStrandBatch was captured without the accumulate variant")
        if self.use_priority:
            with torch.cuda.device(self.dev):
                L.check(self.lib.hgs_graph_launch(self.execs[accumulate], L.stream_ptr(self.dev)), "graph launch")
        else:
            self.graphs[accumulate].replay()
        self.done = torch.cuda.Event()
        self.done.record(torch.cuda.current_stream(self.dev))
        self.sink.accumulate = True
        self.replays += 1
        return self.losses

    def __del__(self):
        try:
            for ex in self.execs.values():
                self.lib.hgs_graph_exec_destroy(ex)
        except Exception:
            pass

    def validate(self):
        """Before the all-reduce / optimiser step: waits for the last replay and raises HgsPlanError if a view did not fit."""
        if self.done is None:
            return []
        self.done.synchronize()
        return [p.check() for p in self.plans]

    def check(self):
        torch.cuda.synchronize(self.dev)
        return [p.check() for p in self.plans] if self.done is not None else []
