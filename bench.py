#!/usr/bin/env python
"""bench.py — headline benchmark of the Hair-GS render path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg5|cfg1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): training views/s of fwd+bwd rasterization.  A *step* is one batch of camera views,
--views-per-step (default 8) per GPU — 8 GPUs x 8 views is the 64-view batch of BASELINE configs[3], SURVEY 8(e) — each
rendered the way Hair-GS trains on it (train.py:146-155, loss/losses.py:341-346): every colour set of the workload
(cfg3: SH-RGB, mask, strand orientation) is rasterized forward and backward, the Gaussian-parameter gradients of the
step's views accumulate in one flat fp32 bucket, and with N > 1 GPUs the bucket is all-reduced over NCCL ONCE per step
(views shard over ranks, one process per GPU, no other data-path collective: weak scaling, per-GPU work fixed).

  value  — device-resident: Gaussian inputs, cameras and dL/dimage already in HBM, the `_C` entry points of the
           drop-in called directly (through the C ABI of libhairgs_rast.so).
  e2e    — the public API a Hair-GS user calls: render(camera, model, bg) + autograd, with the view's camera and
           target images copied host->device from pinned memory and the loss read back device->host every step.
  roofline / stages — per-kernel device times from CUDA events recorded by the library on its launch stream.
  cpu_baseline — the CPU port (oracle/, OpenMP C) on a bounded sample of the same workload, host cores stated.

`--impl reference` runs the same harness on the UNMODIFIED reference rasterizer (oracle/_ref: the reference's own
CUDA sources compiled for sm_100a; the reference has no CPU implementation of this path), falling back to the CPU
port when that build is absent.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "hair-gs_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # name: (kind, S/P, V, W, H, views, colour sets, sh_degree, M)
    "cfg1": dict(kind="strands", S=500, V=101, W=512, H=512, views=1, sets=("sh", "mask", "orientation"), D=0, M=1,
                 desc="50k strand segments, 1 view 512x512"),
    "cfg2": dict(kind="blobs", P=300000, W=512, H=512, views=16, sets=("sh",), D=3, M=16,
                 desc="Stage I GaussianModel: 300k Gaussians, SH degree 3, 16 views 512x512"),
    "cfg3": dict(kind="strands", S=10000, V=100, W=1024, H=1024, views=16, sets=("sh", "mask", "orientation"), D=0, M=1,
                 desc="Stage III HairGaussianModel: 990k strand-aligned Gaussians (10k strands x 99 segments), "
                      "1024x1024, RGB + mask + orientation"),
    "cfg5": dict(kind="strands", S=40000, V=101, W=2048, H=2048, views=4, sets=("sh",), D=0, M=1,
                 desc="stress: 4M strand-aligned Gaussians at 2048x2048"),
}
L2_BYTES = 126 * 1024 * 1024
# arguments/__init__.py:84-86 defaults
# arguments/__init__.py optimisation defaults (position / feature / opacity / scaling / mask learning rates)
ADAM_LRS = {"_endpoints": 1.6e-4, "_xyz": 1.6e-4, "_features_dc": 0.025, "_features_rest": 0.00125, "_opacity": 0.05,
            "_width": 5e-3, "_scaling": 5e-3, "_rotation": 1e-3, "_mask": 0.01}
LOSS_LAMBDAS = dict(lambda_dssim=0.2, lambda_mask=0.01, lambda_orientation=100.0)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="e2e through the eager path only (no CUDA-graph replay)")
    ap.add_argument("--cpu-sample-views", type=int, default=6)
    ap.add_argument("--views-per-step", type=int, default=8,
                    help="views each rank renders per step (one gradient all-reduce / optimiser step per step); "
                         "8 views x 8 GPUs = the 64-view batch of BASELINE configs[3]")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------
def build_workload(cfg, dev):
    import torch
    from hairgs_b200 import models, scenes
    if cfg["kind"] == "strands":
        scene = scenes.strand_scene(cfg["S"], cfg["V"], seed=0, sh_coeffs=cfg["M"]).to(dev)
        model = models.StrandModel(scene, sh_degree=cfg["D"]).to(dev)
    else:
        scene = scenes.blob_scene(cfg["P"], seed=0, sh_coeffs=cfg["M"]).to(dev)
        model = models.BlobModel(scene, sh_degree=cfg["D"]).to(dev)
    cams = scenes.orbit_cameras(max(cfg["views"], 2), cfg["W"], cfg["H"], device=dev)[:cfg["views"]]
    return model, cams


def colour_override(model, which):
    if which == "sh":
        return None
    if which == "mask":
        return model.get_mask.repeat(1, 3)
    return model.get_orientation


class Harness:
    """Everything both arms share; `backend` is the `_C`-shaped extension module under test."""

    def __init__(self, cfg, dev, backend, world, rank, views_per_step=1):
        import torch
        self.torch, self.cfg, self.dev, self.C, self.world, self.rank = torch, cfg, dev, backend, world, rank
        self.vps = max(1, int(views_per_step))
        self.model, self.cams = build_workload(cfg, dev)
        from hairgs_b200 import multiview
        self.bg = torch.zeros(3, device=dev)
        H, W = cfg["H"], cfg["W"]
        self.empty = torch.Tensor([])
        costs = None
        if world > 1:
            # deal views of similar cost (tile-instance count N) into the same lock-step iteration
            with torch.no_grad():
                m = self.model
                costs = [backend.rasterize_gaussians(
                    self.bg, m.get_xyz.contiguous(), self.empty, m.get_opacity.contiguous(), m.get_scaling.contiguous(),
                    m.get_rotation.contiguous(), 1.0, self.empty, c.world_view_transform, c.full_proj_transform,
                    c.tanfovx, c.tanfovy, H, W, m.get_features.contiguous(), cfg["D"], c.camera_center, False, False)[0]
                    for c in self.cams]
        self.view_costs = costs
        self.my_views, _ = multiview.shard_views(len(self.cams), world, rank, costs=costs)
        g = torch.Generator(device="cpu").manual_seed(1234)
        # synthetic targets per view (pinned host memory).  Workloads that train on RGB + mask + orientation carry
        # what Hair-GS's cameras hold (scene/cameras.py:60-85): original_image[3], float_mask[1], orientation_field[1]
        # in [0, pi), orientation_confidence[1]; the others an RGB image.
        self.hair_loss = tuple(cfg["sets"]) == ("sh", "mask", "orientation")
        self.n_tgt = 6 if self.hair_loss else 7
        self.targets_host = []
        for _ in self.my_views:
            t = torch.rand(self.n_tgt, H, W, generator=g)
            if self.hair_loss:
                t[3] = (t[3] < 0.5).float()
                t[4] *= math.pi
            self.targets_host.append(t.pin_memory())
        # device-resident inputs of the `value` loop
        with torch.no_grad():
            m = self.model
            self.inputs = dict(means3D=m.get_xyz.contiguous(), opacity=m.get_opacity.contiguous(),
                               scales=m.get_scaling.contiguous(), rotations=m.get_rotation.contiguous(),
                               sh=m.get_features.contiguous())
            self.colours = {s: (None if s == "sh" else colour_override(m, s).contiguous()) for s in cfg["sets"]}
        self.dL = {s: torch.randn(3, H, W, generator=g).to(dev) / (H * W) for s in cfg["sets"]}
        P, M = self.inputs["means3D"].shape[0], self.inputs["sh"].shape[1]
        self.P, self.M = P, M
        # flat gradient bucket: means3D 3, scales 3, rotations 4, opacity 1, sh 3M, + 3 per override colour set
        shapes = {"means3D": (P, 3), "scales": (P, 3), "rotations": (P, 4), "opacity": (P, 1), "sh": (P, M, 3)}
        for s in cfg["sets"]:
            if s != "sh":
                shapes["colour_" + s] = (P, 3)
        self.bucket = multiview.GradBucket(shapes, dev)
        self.empty = torch.Tensor([])
        self.last_N = 0

    # ---- device-resident step: direct _C calls ------------------------------------------------------
    def step_resident(self, it):
        b = self.bucket.zero_()
        for k in range(self.vps):
            self._resident_view(it * self.vps + k, b)
        b.all_reduce()  # one collective per step; no-op on a single rank

    def _resident_view(self, vi, b):
        torch, C, cfg, i = self.torch, self.C, self.cfg, self.inputs
        cam = self.cams[self.my_views[vi % len(self.my_views)]]
        for s in cfg["sets"]:
            col = self.colours[s]
            sh = i["sh"] if col is None else self.empty
            colors = self.empty if col is None else col
            N, color, radii, geom, binning, img = C.rasterize_gaussians(
                self.bg, i["means3D"], colors, i["opacity"], i["scales"], i["rotations"], 1.0, self.empty,
                cam.world_view_transform, cam.full_proj_transform, cam.tanfovx, cam.tanfovy, cfg["H"], cfg["W"], sh,
                cfg["D"], cam.camera_center, False, False)
            g2d, gcol, gop, gm3, gcov, gsh, gsc, grot = C.rasterize_gaussians_backward(
                self.bg, i["means3D"], radii, colors, i["scales"], i["rotations"], 1.0, self.empty,
                cam.world_view_transform, cam.full_proj_transform, cam.tanfovx, cam.tanfovy, self.dL[s], sh, cfg["D"],
                cam.camera_center, geom, N, binning, img, False)
            b.accumulate("means3D", gm3).accumulate("scales", gsc).accumulate("rotations", grot).accumulate("opacity", gop)
            if col is None:
                b.accumulate("sh", gsh)
            else:
                b.accumulate("colour_" + s, gcol)
            self.last_N = N

    # ---- device-resident fused step: strand parameterisation + 7 channels in one pass ------------------
    def setup_fused(self):
        torch = self.torch
        from hairgs_b200 import fused
        self.fused_mod = fused
        H, W = self.cfg["H"], self.cfg["W"]
        g = torch.Generator(device="cpu").manual_seed(99)
        self.dL7 = (torch.randn(7, H, W, generator=g) / (H * W)).to(self.dev)
        self.bg7 = torch.zeros(7, device=self.dev)
        m = self.model
        self.fparams = {n: p for n, p in m.named_parameters() if p.numel() > 0}
        from hairgs_b200 import multiview
        self.fbucket = multiview.GradBucket({n: p.shape for n, p in self.fparams.items()}, self.dev)
        self.fbucket.attach_to(self.fparams)
        self.fsink = self._grad_sink()

    def setup_fused_graph(self):
        """The resident fused step as a CUDA-graph replay (hairgs_b200.graphs, dL/dimage given): same kernels, one launch."""
        torch = self.torch
        self.fgraph, self.fgraph_note = None, None
        if self.fsink is None:
            return
        try:
            from hairgs_b200 import graphs
            mine = [self.cams[v] for v in self.my_views]
            cap, bits = graphs.measure_plan(self.model, mine, self.bg7)
            g = graphs.GraphedStrandStep(self.model, self.fsink, self.bg7, self.cfg["H"], self.cfg["W"], mine[0].FoVx,
                                         mine[0].FoVy, cap, bits, dimage=self.dL7.contiguous())
            self.cam_flat = [torch.cat([c.world_view_transform.reshape(-1), c.full_proj_transform.reshape(-1),
                                        c.camera_center.reshape(-1)]).contiguous() for c in mine]
            for slot in range(2):
                g.cam_buf[slot].copy_(self.cam_flat[slot % len(mine)])
            torch.cuda.synchronize(self.dev)
            g.capture()
            self.fgraph = g
            self.fgraph_note = f"CUDA-graph replay, plan: capacity {cap} instances, {bits} depth bits"
        except Exception as e:
            self.fgraph_note = f"graph capture failed, eager path used: {type(e).__name__}: {e}"
            torch.cuda.synchronize(self.dev)

    def step_resident_fused_graph(self, it):
        for k in range(self.vps):
            vi = it * self.vps + k
            slot = vi % 2
            # the view's camera is resident; 140 bytes device-to-device into the graph's input slot
            self.fgraph.cam_buf[slot].copy_(self.cam_flat[vi % len(self.cam_flat)], non_blocking=True)
            self.fgraph.replay(slot, accumulate=k > 0)
        self.fbucket.all_reduce()

    def _grad_sink(self):
        """The strand backward writes the parameter gradients straight into the slices of the flat bucket that the
        parameters' .grad point at (fused.GradSink): no autograd accumulation kernels, no bucket clear."""
        m = self.model
        if m._features_rest.numel() != 0 or os.environ.get("HGS_BENCH_NO_SINK"):
            return None     # features = cat(dc, rest): left to autograd
        return self.fused_mod.GradSink({"endpoints": m._endpoints.grad, "width": m._width.grad, "opacity": m._opacity.grad,
                                        "mask": m._mask.grad, "features": m._features_dc.grad})

    def step_resident_fused(self, it):
        m = self.model
        if self.fsink is None:
            self.fbucket.zero_()
        else:
            self.fsink.begin_step()      # the first view of the step overwrites the bucket, the others add
        for k in range(self.vps):
            cam = self.cams[self.my_views[(it * self.vps + k) % len(self.my_views)]]
            out = self.fused_mod.render_strands(cam, m, self.bg7, grad_sink=self.fsink)
            out["image7"].backward(self.dL7)
        self.fbucket.all_reduce()
        self.last_N = 0

    # ---- end-to-end step: render() + autograd, host<->device copies inside --------------------------
    def setup_e2e(self, fused=False, optimizer=None, graph=False):
        """optimizer: None (gradients only), "flat" (hairgs_b200.optim.FlatAdam: one launch over the flat bucket) or
        "torch" (torch.optim.Adam(eps=1e-15) + zero_grad(set_to_none=True), scene/gaussian_model.py:250, train.py:203-204).
        graph: replay the view (render_strands + hair_image_loss + backward) as one CUDA graph
        (hairgs_b200.graphs.GraphedStrandStep); the copies, the all-reduce and the optimiser stay outside the graph."""
        torch = self.torch
        self.fused = fused
        self.opt_mode = optimizer
        self.graphed, self.graph_note = None, None
        if fused:
            from hairgs_b200 import fused as fused_mod
            from hairgs_b200.fused import render_strands
            from hairgs_b200 import losses
            self.fused_mod = fused_mod
            self.render_strands = render_strands
            self.weighted_l1 = losses.weighted_l1
            self.hair_image_loss = losses.hair_image_loss
            self.w7 = losses.l1_groups([(0, 3, 1.0), (3, 4, 0.01), (4, 7, 1.0)], self.cfg["H"], self.cfg["W"], self.dev)
            self.bg7 = torch.zeros(7, device=self.dev)
        import diff_gaussian_rasterization as dgr
        dgr._RasterizeGaussians.backend = self.C
        from gaussian_renderer import render
        self.render = render
        self.params = [p for p in self.model.parameters() if p.numel() > 0]
        n = sum(p.numel() for p in self.params)
        self.flat_grad = torch.zeros(n, device=self.dev)
        o = 0
        for p in self.params:
            p.grad = self.flat_grad[o:o + p.numel()].view_as(p)
            o += p.numel()
        if optimizer is not None:
            # Adam moves every parameter by ~lr per step whatever the gradient's size; with random synthetic targets the
            # reference learning rates would scramble the strands within the timed loop (segments are ~2 mm), so they
            # are scaled by 1e-3: identical optimiser work per step, a scene that stays the stated workload.
            groups = [{"params": [p], "lr": 1e-3 * ADAM_LRS.get(n, 1e-3), "name": n}
                      for n, p in self.model.named_parameters() if p.numel() > 0]
            if optimizer == "flat":
                from hairgs_b200.optim import FlatAdam
                self.opt = FlatAdam(groups)          # re-homes p.data / p.grad into its flat buffers
            else:
                for p in self.params:
                    p.grad = None
                self.opt = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
        self.esink = self._grad_sink() if fused else None
        if getattr(self, "copy_stream", None) is not None:
            self._prefetched = -1
            self._setup_graph(graph)
            return
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        H, W = self.cfg["H"], self.cfg["W"]
        self.tgt_dev = [torch.empty(self.n_tgt, H, W, device=self.dev) for _ in range(2)]
        from hairgs_b200 import losses
        self.hair_image_loss_torch = losses.hair_image_loss_torch
        self.view_rot = [losses.view_rot_of(self.cams[v].world_view_transform.cpu()) for v in self.my_views]
        self.cam_host = []
        for v in self.my_views:
            c = self.cams[v]
            self.cam_host.append(torch.cat([c.world_view_transform.reshape(-1), c.full_proj_transform.reshape(-1),
                                            c.camera_center.reshape(-1)]).cpu().pin_memory())
        self.cam_dev = [torch.empty(35, device=self.dev) for _ in range(2)]
        self.copy_done = [torch.cuda.Event() for _ in range(2)]
        self.slot_free = [torch.cuda.Event() for _ in range(2)]
        self.loss_host = torch.zeros(1).pin_memory()
        self.h2d_bytes = self.vps * (self.n_tgt * H * W * 4 + 35 * 4)
        self.d2h_bytes = self.vps * 4
        self._prefetched = -1
        self._setup_graph(graph)

    def _setup_graph(self, graph):
        """Captures the fused view into CUDA graphs over the SAME double-buffered input slots the copy stream fills."""
        if not (graph and self.fused and self.hair_loss and self.esink is not None):
            return
        torch = self.torch
        try:
            from hairgs_b200 import graphs
            mine = [self.cams[v] for v in self.my_views]
            cap, bits = graphs.measure_plan(self.model, mine, self.bg7)
            c0 = mine[0]
            g = graphs.GraphedStrandStep(self.model, self.esink, self.bg7, self.cfg["H"], self.cfg["W"], c0.FoVx, c0.FoVy,
                                         cap, bits, lambdas=LOSS_LAMBDAS, cam_buf=self.cam_dev, tgt_buf=self.tgt_dev)
            for slot in range(2):  # a real view in every slot before the warm-up / capture
                k = slot % len(self.my_views)
                self.tgt_dev[slot].copy_(self.targets_host[k])
                self.cam_dev[slot].copy_(self.cam_host[k])
            torch.cuda.synchronize(self.dev)
            g.capture()
            self.graphed = g
            self.graph_note = f"one CUDA graph per input slot, plan: capacity {cap} instances, {bits} depth bits"
        except Exception as e:  # stay measurable: fall back to the eager path and say so in the JSON line
            self.graphed = None
            self.graph_note = f"graph capture failed, eager path used: {type(e).__name__}: {e}"
            torch.cuda.synchronize(self.dev)

    def _prefetch(self, it):
        """stage view `it` into slot it%2 on the copy stream (double-buffered data loader)."""
        torch = self.torch
        slot, k = it % 2, it % len(self.my_views)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.slot_free[slot])
            self.tgt_dev[slot].copy_(self.targets_host[k], non_blocking=True)
            self.cam_dev[slot].copy_(self.cam_host[k], non_blocking=True)
            self.copy_done[slot].record(self.copy_stream)
        self._prefetched = it

    def step_e2e(self, it):
        """One step = views_per_step views (each: H2D of its camera/targets, render, loss, backward, D2H of the loss), then ONE
        gradient all-reduce and ONE optimiser step."""
        torch = self.torch
        for k in range(self.vps):
            self._e2e_view(it * self.vps + k, first=k == 0)
        if self.world > 1:
            torch.distributed.all_reduce(self.opt.grads.flat if self.opt_mode == "flat" else self.flat_grad)
        if self.opt_mode == "flat":
            # the sink overwrites the bucket on the next step, so the optimiser kernel need not clear it
            self.opt.step(grad_scale=1.0 / (self.world * self.vps), zero_grad=self.esink is None)
        elif self.opt_mode == "torch":
            self.opt.step()
            self.opt.zero_grad(set_to_none=True)

    def _e2e_view(self, it, first):
        torch, cfg = self.torch, self.cfg
        from hairgs_b200.scenes import Camera
        if self._prefetched < it:
            self._prefetch(it)
        slot = it % 2
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self.copy_done[slot])
        if self._prefetched < it + 1:
            self._prefetch(it + 1)  # next view's copies overlap this view's kernels
        base = self.cams[self.my_views[it % len(self.my_views)]]
        cd = self.cam_dev[slot]
        cam = Camera(base.image_width, base.image_height, base.FoVx, base.FoVy, cd[0:16].view(4, 4), cd[16:32].view(4, 4),
                     cd[32:35])
        tgt = self.tgt_dev[slot]
        if first:
            if self.esink is not None:
                self.esink.begin_step()      # the first view of the step overwrites the gradients, the others add
            elif self.opt_mode is None:
                self.flat_grad.zero_()
        m = self.model
        loss = None
        lam = LOSS_LAMBDAS
        if self.graphed is not None:
            # the same view as the branch below, replayed as ONE graph launch (inputs: this slot's camera / targets)
            loss = self.graphed.replay(slot, accumulate=not first)
        elif self.fused and self.hair_loss:
            # ONE fused pass: strand parameterisation + 7 channels (hairgs_b200.fused.render_strands), then Hair-GS's
            # image loss (l1 + d-ssim + BCE mask + orientation, loss/losses.py:319-346) as one fused op
            out = self.render_strands(cam, m, self.bg7, grad_sink=self.esink)
            loss, _ = self.hair_image_loss(out["image7"], tgt[0:3], tgt[3], tgt[4], tgt[5],
                                           self.view_rot[it % len(self.my_views)], orient_mask=tgt[3] > 0.5, **lam)
        elif self.fused:
            out = self.render_strands(cam, m, self.bg7, grad_sink=self.esink)
            loss = self.weighted_l1(out["image7"], tgt, self.w7)
        elif self.hair_loss:
            # the reference's composition: three render() calls (loss/losses.py:245-248, 311-312, train.py:146-155) and
            # the torch ops of loss_function
            outs = {s: self.render(cam, m, self.bg, override_color=colour_override(m, s))["render"] for s in cfg["sets"]}
            loss, _ = self.hair_image_loss_torch(outs["sh"], outs["mask"][0], outs["orientation"], tgt[0:3], tgt[3], tgt[4],
                                                 tgt[5], cam.world_view_transform, orient_mask=tgt[3] > 0.5, **lam)
        else:
            for s in cfg["sets"]:
                out = self.render(cam, m, self.bg, override_color=colour_override(m, s))["render"]
                term = (out - tgt[0:3]).abs().mean()
                loss = term if loss is None else loss + term
        if self.graphed is None:
            loss.backward()
        self.slot_free[slot].record(cur)
        self.loss_host.copy_(loss.detach().reshape(1), non_blocking=True)


def timed_loop(torch, step_fn, steps, warmup, world, dev, flush=None):
    """W untimed + exactly K timed steps; barrier + synchronize on both sides; device time via CUDA events;
    returns the max over ranks of the elapsed milliseconds."""
    for it in range(warmup):
        step_fn(it)
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(dev)
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(warmup, warmup + steps):
            step_fn(it)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
    else:
        evs = []
        for it in range(warmup, warmup + steps):
            flush.zero_()  # evict L2 between timed iterations (outside the event brackets)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_fn(it)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize(dev)
        ms = sum(a.elapsed_time(b) for a, b in evs)
    if world > 1:
        torch.distributed.barrier()
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    return ms


# ---------------------------------------------------------------------------------------------------
# CPU port (oracle) timing — bounded sample
# ---------------------------------------------------------------------------------------------------
def cpu_port_views_per_s(cfg, n_views):
    import numpy as np
    import torch
    from oracle import pyoracle
    model, cams = build_workload(cfg, "cpu")
    with torch.no_grad():
        base = dict(background=np.zeros(3, np.float32), means3D=model.get_xyz.numpy(), opacity=model.get_opacity.numpy(),
                    scales=model.get_scaling.numpy(), rotations=model.get_rotation.numpy(), cov3D_precomp=None,
                    scale_modifier=1.0, image_height=cfg["H"], image_width=cfg["W"], degree=cfg["D"])
        cols = {s: (None if s == "sh" else colour_override(model, s).numpy()) for s in cfg["sets"]}
        sh = model.get_features.numpy()
    rng = np.random.default_rng(0)
    dL = rng.standard_normal((3, cfg["H"], cfg["W"])).astype(np.float32)
    stats = None
    t0 = time.perf_counter()
    for v in range(n_views):
        cam = cams[v % len(cams)]
        for s in cfg["sets"]:
            d = dict(base, viewmatrix=cam.world_view_transform.numpy(), projmatrix=cam.full_proj_transform.numpy(),
                     campos=cam.camera_center.numpy(), tan_fovx=cam.tanfovx, tan_fovy=cam.tanfovy,
                     sh=sh if cols[s] is None else None, colors=cols[s])
            f = pyoracle.Forward(d)
            f.backward(dL)
            if stats is None:
                # what SURVEY 8(d) asks the generator to report about the synthetic scene (first sampled view)
                r = f.array("ranges").astype(np.int64)
                ln = r[:, 1] - r[:, 0]
                stats = {"P": int(f.P), "num_rendered": int(f.N), "visible_fraction": round(float((f.radii > 0).mean()), 4),
                         "non_empty_tiles": int((ln > 0).sum()), "tiles": int(ln.shape[0]),
                         "mean_tile_list": round(float(ln[ln > 0].mean()) if (ln > 0).any() else 0.0, 1),
                         "max_tile_list": int(ln.max()) if ln.size else 0,
                         "median_radius_px": float(np.median(f.radii[f.radii > 0])) if (f.radii > 0).any() else 0.0}
            f.close()
    dt = time.perf_counter() - t0
    cpu_port_views_per_s.scene_stats = stats
    return n_views / dt, pyoracle.num_threads(), dt


def cpu_torch_naive(cfg, max_tiles=96):
    """The CPU baseline as BASELINE.json's north_star words it: the reference's torch-side preprocessing plus a naive torch
    compositor (oracle/torch_baseline.py), forward + autograd backward of one colour set on a
    bounded sample of tiles, extrapolated; a view renders len(cfg["sets"]) colour sets."""
    import torch
    from oracle import torch_baseline as tb
    model, cams = build_workload(cfg, "cpu")
    cam = cams[0]
    with torch.no_grad():
        d = dict(background=torch.zeros(3), means3D=model.get_xyz, colors=torch.Tensor([]), opacity=model.get_opacity,
                 scales=model.get_scaling, rotations=model.get_rotation, scale_modifier=1.0, cov3D_precomp=torch.Tensor([]),
                 viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, tan_fovx=cam.tanfovx,
                 tan_fovy=cam.tanfovy, image_height=cfg["H"], image_width=cfg["W"], sh=model.get_features.contiguous(),
                 degree=cfg["D"], campos=cam.camera_center)
        d = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in d.items()}
    strand = None
    if cfg["kind"] == "strands":
        strand = (model._endpoints.detach().clone(), model.endpoint_pairs, model._width.detach().clone())
    torch.set_num_threads(os.cpu_count() or 1)
    r = tb.time_view(d, strand=strand, max_tiles=max_tiles)
    sets = len(cfg["sets"])
    return {"value": round(1000.0 / (sets * r["ms_per_view"]), 5), "unit": "views/s", "cores": r["cores"],
            "ms_per_colour_set": round(r["ms_per_view"], 1), "colour_sets_per_view": sets, "items_ms": r["items"],
            "sample": r["sample"]}


# ---------------------------------------------------------------------------------------------------
def algorithmic_bytes(stage, P, N, HW, T, M, D, sh_mode, C=3):
    """SURVEY.md §8(d) compulsory traffic per launch of each stage."""
    if stage == "preprocess_fwd":
        return P * (44 + (12 * (D + 1) ** 2 if sh_mode else 12) + 8) + P * (28 + 12)
    if stage == "tile_scan":
        return 8 * P
    if stage == "emit_keys":
        return 20 * P + 12 * N
    if stage == "sort_histogram":
        return 8 * N
    if stage == "sort_onesweep":
        return 24 * N  # one pass: read + write of (u64 key, u32 value)
    if stage == "tile_ranges":  # finalize_sorted: keys + ids read, 32 B record + colour gathered and re-written sorted
        cs = 16 if C <= 4 else 32
        return 12 * N + 2 * (32 + cs) * N + 12 * T
    if stage == "composite_fwd":
        return (28 + 4 * C) * N + (8 + 4 * C) * HW
    if stage == "composite_bwd":
        return (28 + 4 * C) * N + (8 + 4 * C) * HW + 4 * (6 + C) * N
    if stage == "preprocess_bwd":
        return P * (96 + 40) + (P * (12 * (D + 1) ** 2 + 15) + P * 12 * M if sh_mode else 0)
    return 0


class StdoutToStderr:
    """Route fd 1 to stderr while native code may print (the reference writes a banner to stdout from
    BACKWARD::render, backward_distwar.cu:1121-1205; NCCL may print its version) so that stdout carries exactly ONE
    line: the JSON result."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    with StdoutToStderr():
        line = run()
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


def run():
    args = parse_args()
    cfg = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import torch
    config = {"workload": f"{args.workload}: {cfg['desc']}", "views": cfg["views"], "colour_sets": list(cfg["sets"]),
              "resolution": [cfg["W"], cfg["H"]], "sharding": f"views dealt over {world} rank(s), ordered by tile-instance count" if world > 1 else "1 rank",
              "e2e_loss": ("Hair-GS image loss: (1-0.2) l1 + 0.2 d-ssim + 0.01 BCE mask + 100 orientation "
                           "(loss/losses.py:319-346, arguments/__init__.py:84-86)"
                           if tuple(cfg["sets"]) == ("sh", "mask", "orientation") else "l1")}
    base_line = {"metric": "train views/s (fwd+bwd rasterize incl. grad accumulation" +
                           (", NCCL all-reduce" if world > 1 else "") + ")", "unit": "views/s", "n_gpus": world,
                 "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
                 "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config}

    use_ref_gpu = False
    if args.impl == "reference":
        if rank != 0:
            return None  # the reference is single-GPU (utils/general.py:116): rank 0 alone measures it
        world = 1
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import refload
        use_ref_gpu = torch.cuda.is_available() and refload.ref_dgr() is not None
        if not use_ref_gpu:
            # CPU port of the path on all host cores, bounded sample per step
            vps, cores, dt = cpu_port_views_per_s(cfg, max(1, min(args.steps, args.cpu_sample_views)))
            line = dict(base_line, impl="reference", value=vps, n_gpus=1, ms_per_step=1000.0 / vps,
                        cpu_baseline={"value": vps, "unit": "views/s", "cores": cores, "kind": "port",
                                      "sample": f"{args.cpu_sample_views} view(s) of {args.workload}, all colour sets, fwd+bwd"},
                        e2e={"value": vps, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                        gpu_launches=0)
            return line

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # the version banner goes to stdout and would precede the JSON line
        torch.distributed.init_process_group("nccl", device_id=dev)

    if args.impl == "reference":
        import refload
        backend = refload.ref_dgr()
    else:
        import diff_gaussian_rasterization._C as backend
    from hairgs_b200 import _lib as L
    lib = L.load()

    h = Harness(cfg, dev, backend, world, rank, views_per_step=args.views_per_step)
    config["views_per_step"] = (f"{h.vps} per rank ({h.vps * world}-view batch): gradients of the step's views accumulate in the "
                                f"flat bucket, ONE all-reduce and ONE optimiser step per step")
    P, M, D = h.P, h.M, cfg["D"]
    HW = cfg["W"] * cfg["H"]
    T = ((cfg["W"] + 15) // 16) * ((cfg["H"] + 15) // 16)

    # working-set estimate of one step -> L2 policy
    sets = len(cfg["sets"])
    h.step_resident(0)
    torch.cuda.synchronize(dev)
    N = h.last_N
    ws = P * (56 + 12 * (M - 1)) + sets * (P * 69 + N * 24 + HW * 44 + P * 4 * (11 + 19 + 3 * M)) + h.bucket.flat.numel() * 4
    flush = None
    if ws < 2 * L2_BYTES:
        flush = torch.empty(2 * L2_BYTES // 4, device=dev)
        config["l2"] = f"explicit flush: {2 * L2_BYTES >> 20} MiB written between timed steps (working set {ws >> 20} MiB)"
    else:
        config["l2"] = f"no flush: per-step working set ~{ws >> 20} MiB > 126 MiB L2, views rotate every step"
    config["P"], config["num_rendered"] = P, N

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    import ctypes
    ms_res = timed_loop(torch, h.step_resident, args.steps, args.warmup, world, dev, flush)
    # kernels of libhairgs_rast.so launched per step (the library counts its own launches)
    launches = (ctypes.c_int64 * 16)()
    lib.hgs_profile_collect(None, launches)  # reset
    h.step_resident(0)
    launches = (ctypes.c_int64 * 16)()
    lib.hgs_profile_collect(None, launches)
    launches_per_step = int(sum(launches))

    can_fuse = args.impl == "ours" and cfg["kind"] == "strands" and tuple(cfg["sets"]) == ("sh", "mask", "orientation")
    ms_res_3pass, ms_e2e_3pass = ms_res, None
    ms_e2e_eager, graph_note = None, None
    ms_res_eager, res_graph_note = None, None
    h.setup_e2e(fused=False)
    ms_e2e = timed_loop(torch, h.step_e2e, args.steps, args.warmup, world, dev, flush)
    if can_fuse:
        # the same view (same 7 output planes, same parameter gradients) through the fused strand entry
        ms_e2e_3pass = ms_e2e
        h.setup_fused()
        ms_res = timed_loop(torch, h.step_resident_fused, args.steps, args.warmup, world, dev, flush)
        launches = (ctypes.c_int64 * 16)()
        lib.hgs_profile_collect(None, launches)
        h.step_resident_fused(0)
        launches = (ctypes.c_int64 * 16)()
        lib.hgs_profile_collect(None, launches)
        launches_per_step = int(sum(launches))
        if not args.no_graph:
            h.setup_fused_graph()
            if h.fgraph is not None:
                ms_res_eager = ms_res
                ms_res = timed_loop(torch, h.step_resident_fused_graph, args.steps, args.warmup, world, dev, flush)
                h.fgraph.check()
            res_graph_note = h.fgraph_note
        h.setup_e2e(fused=True)
        ms_e2e = timed_loop(torch, h.step_e2e, args.steps, args.warmup, world, dev, flush)
        if h.hair_loss and not args.no_graph:
            # the same step with the view replayed as one CUDA graph (the eager loop is host-bound: ~25 launches and
            # ~0.9 ms of Python per view); eager number kept as e2e.value_eager
            h.setup_e2e(fused=True, graph=True)
            if h.graphed is not None:
                ms_e2e_eager = ms_e2e
                ms_e2e = timed_loop(torch, h.step_e2e, args.steps, args.warmup, world, dev, flush)
                h.graphed.check()
            graph_note = h.graph_note
    # the same e2e step with the optimiser included ("optimiser step excluded and also reported included");
    # runs last because it moves the parameters
    h.setup_e2e(fused=can_fuse, optimizer="flat" if args.impl == "ours" else "torch",
                graph=can_fuse and ms_e2e_eager is not None)
    ms_e2e_opt = timed_loop(torch, h.step_e2e, args.steps, args.warmup, world, dev, flush)
    if h.graphed is not None:
        h.graphed.check()
    clk = clocks.stop() if rank == 0 else None
    import diff_gaussian_rasterization as dgr
    dgr._RasterizeGaussians.backend = dgr._C

    views = world * args.steps * h.vps
    value = views / (ms_res / 1000.0)
    e2e_value = views / (ms_e2e / 1000.0)

    # ---- per-stage device times (CUDA events recorded by the library on its launch stream) ----------
    stages = {}
    roofline = None
    if args.impl == "ours":
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json hbm_gbs)") if peaks.get("hbm_gbs") else (6650.0, "fallback 6.65 TB/s")
        lib.hgs_profile_collect(None, None)  # reset the launch counters accumulated by the e2e loop
        lib.hgs_profile_enable(1)
        prof_steps = min(args.steps, 8)
        prof_step = h.step_resident_fused if can_fuse else h.step_resident
        C_prof = 7 if can_fuse else 3
        for it in range(prof_steps):
            prof_step(it)
        ms = (ctypes.c_double * 16)()
        cnt = (ctypes.c_int64 * 16)()
        lib.hgs_profile_collect(ms, cnt)
        lib.hgs_profile_enable(0)
        total = sum(ms)
        for sidx in range(11):
            if cnt[sidx] == 0:
                continue
            name = lib.hgs_stage_name(sidx).decode()
            per_launch_ms = ms[sidx] / cnt[sidx]
            ab = algorithmic_bytes(name, P, N, HW, T, M, D, True, C_prof)
            stages[name] = {"ms_per_launch": round(per_launch_ms, 5), "launches_per_step": cnt[sidx] / prof_steps,
                            "share": round(ms[sidx] / total, 4) if total else None,
                            "achieved_GBps": round(ab / per_launch_ms / 1e6, 1) if per_launch_ms > 0 and ab else None,
                            "frac_of_hbm_peak": round(ab / per_launch_ms / 1e6 / peak, 4) if per_launch_ms > 0 and ab else None}
        dom = max(stages, key=lambda k: stages[k]["ms_per_launch"] * stages[k]["launches_per_step"])
        ab = algorithmic_bytes(dom, P, N, HW, T, M, D, True, C_prof)
        ach = ab / stages[dom]["ms_per_launch"] / 1e6
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get(args.workload, {}).get(dom)
        except Exception:
            pass
        roofline = {"kernel": dom, "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": ab,
                    "note": "compositors are issue/latency-bound (serial transmittance chain), see DESIGN.md; "
                            "HBM-bound stages are listed under 'stages'"}

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return None

    line = dict(base_line, value=round(value, 2), ms_per_step=round(ms_res / args.steps, 4),
                e2e={"value": round(e2e_value, 2), "unit": "views/s", "h2d_bytes_per_step": h.h2d_bytes,
                     "d2h_bytes_per_step": h.d2h_bytes, "ms_per_step": round(ms_e2e / args.steps, 4),
                     "api": "gaussian_renderer.render() + torch loss (loss/losses.py composition) + autograd, "
                            "targets/camera prefetched from pinned host memory"},
                gpu_launches=launches_per_step * args.steps, clocks=clk)
    line["e2e"]["value_incl_optimizer"] = round(views / (ms_e2e_opt / 1000.0), 2)
    line["e2e"]["optimizer"] = (("hairgs_b200.optim.FlatAdam (hgs_adam_step, one launch)" if args.impl == "ours"
                                 else "torch.optim.Adam(eps=1e-15) + zero_grad(set_to_none=True)") +
                                "; Hair-GS learning rates x 1e-3 (random targets must not scramble the scene)")
    if can_fuse:
        config["path"] = ("fused strand entry: strand parameterisation + RGB/mask/orientation in ONE rasterization pass "
                          "(hairgs_b200.fused.render_strands); the three-pass drop-in path is reported as *_dropin_3pass")
        line["value_dropin_3pass"] = round(views / (ms_res_3pass / 1000.0), 2)
        if res_graph_note is not None:
            config["value_path"] = res_graph_note
        if ms_res_eager is not None:
            line["value_eager"] = round(views / (ms_res_eager / 1000.0), 2)
        line["e2e"]["value_dropin_3pass"] = round(views / (ms_e2e_3pass / 1000.0), 2)
        line["e2e"]["api"] = ("hairgs_b200.fused.render_strands() + hairgs_b200.losses."
                              + ("hair_image_loss()" if h.hair_loss else "weighted_l1()") +
                              " + autograd, targets/camera prefetched from pinned host memory")
        if graph_note is not None:
            line["e2e"]["graph"] = graph_note
        if ms_e2e_eager is not None:
            line["e2e"]["api"] = ("hairgs_b200.graphs.GraphedStrandStep.replay(): render_strands() + hair_image_loss() + "
                                  "backward captured as one CUDA graph per input slot; targets/camera copied from pinned "
                                  "host memory into the slot every step, loss copied back; all-reduce / optimiser outside "
                                  "the graph")
            line["e2e"]["value_eager"] = round(views / (ms_e2e_eager / 1000.0), 2)
    line["n_gpus"] = world
    if args.impl == "reference":
        line["impl"] = "reference"
        line["cpu_baseline"] = {"value": round(value, 2), "unit": "views/s", "cores": 0, "kind": "reference",
                                "sample": "the reference has no CPU implementation of this path: this is its own CUDA "
                                          "rasterizer (oracle/_ref, sm_100a build, BW_IMPLEMENTATION=1 "
                                          "BALANCE_THRESHOLD=8 as train.py:278) on the same GPU, full workload"}
        line["gpu_launches"] = 0
    else:
        line["roofline"] = roofline
        line["stages"] = stages
        if not args.no_cpu_baseline and world == 1:
            vps, cores, dt = cpu_port_views_per_s(cfg, args.cpu_sample_views)
            line["cpu_baseline"] = {"value": round(vps, 4), "unit": "views/s", "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_sample_views} view(s) of {args.workload} (all colour sets, "
                                              f"fwd+bwd) through the OpenMP C port in oracle/, {dt:.1f} s"}
            if getattr(cpu_port_views_per_s, "scene_stats", None):
                config["scene_stats"] = cpu_port_views_per_s.scene_stats
            try:
                # the baseline as the north_star words it (torch preprocessing + naive torch compositor), next to the
                # much faster C port above
                line["cpu_baseline"]["torch_naive"] = cpu_torch_naive(cfg)
            except Exception as e:
                line["cpu_baseline"]["torch_naive"] = {"unavailable": f"{type(e).__name__}: {e}"}
    if world > 1:
        torch.distributed.destroy_process_group()
    return line


if __name__ == "__main__":
    sys.exit(main())
