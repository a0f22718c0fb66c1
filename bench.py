#!/usr/bin/env python
"""bench.py — throughput of the RoI hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

Headline workload (BASELINE.json configs[3], the one the north_star's >=60 %-of-HBM target is stated on):
  training-mode PyramidROIAlign, batch 16 images x 512 sampled RoIs, 256 channels, P2..P5 of a 1024x1024 image:
  7x7 forward + backward, 14x14 forward + backward, and the 28x28 mask-target crop (168 positives / image).
  One "step" = that whole pass over one batch.  metric = RoIs/s (each RoI = all five ops above).

  value      : device-resident inputs, CUDA-event timed, max over ranks (weak scaling: every rank owns a batch)
  e2e        : same step through the C ABI with HOST (pinned) buffers: H2D of pyramid/boxes/upstream grads and
               D2H of every result inside the timed region, pipelined per image over three streams
  roofline   : dominant kernel, algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference : the reference's own CPU extension (oracle/_ref, built from /root/reference)
               driven the way model.roi_align + CropFunction.backward drive it, on the box's host cores
  also       : the other BASELINE configs (proposal layer images/s, forward-only RoIs/s, detection path images/s)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMAGE = 1024
CHANNELS = 256
STRIDES = (4, 8, 16, 32)
BATCH = 16
ROIS_PER_IMAGE = 512
MASK_POS = 168          # 512 * ROI_POSITIVE_RATIO(0.33)
GT_PER_IMAGE = 8
SEED = 1234 + 3
LEVEL_HW = [(IMAGE // s, IMAGE // s) for s in STRIDES]
PYR_ELEMS_PER_IMAGE = CHANNELS * sum(h * w for h, w in LEVEL_HW)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
def image_inputs(image_id, pyramid=True):
    """The synthetic inputs of ONE image of the headline workload, a function of the image id alone: the GPU arm (rank r
    owns images r * BATCH ... r * BATCH + BATCH - 1) and the CPU reference arm (worker w owns image w % BATCH) build their
    batches from the same images.  NCHW fp32 numpy, the reference's layout."""
    from maskrcnn_b200 import synth
    rng = np.random.default_rng(SEED * 1000 + image_id)
    st = {"fms": [rng.standard_normal((1, CHANNELS, h, w), dtype=np.float32) for h, w in LEVEL_HW] if pyramid else None}
    st["boxes"] = synth.random_rois(ROIS_PER_IMAGE, SEED + 17 * image_id)
    st["mboxes"] = synth.random_rois(MASK_POS, SEED + 1000 + image_id)
    rects = []
    for _ in range(GT_PER_IMAGE):                      # gt masks: filled rectangles
        y, x = rng.integers(0, IMAGE - 64, 2)
        h, w = rng.integers(32, 512, 2)
        rects.append((int(y), int(x), int(h), int(w)))
    st["rects"] = rects
    st["mind"] = rng.integers(0, GT_PER_IMAGE, MASK_POS).astype(np.int32)      # instance of every positive RoI, image-local
    return st


def gt_masks_of(rects):
    gt = np.zeros((len(rects), 1, IMAGE, IMAGE), np.float32)
    for k, (y, x, h, w) in enumerate(rects):
        gt[k, 0, y:y + h, x:x + w] = 1.0
    return gt


class Workload(object):
    """Device-resident buffers + the five launches of one step, straight through the C ABI."""

    def __init__(self, torch, device, batch=BATCH, seed=SEED, crops_channels_last=True, first_image=0):
        from maskrcnn_b200 import _lib
        self.torch, self.L, self.batch = torch, _lib, batch
        self.cl_crops = crops_channels_last
        self.crop_layout = _lib.NHWC if crops_channels_last else _lib.NCHW
        cfmt = torch.channels_last if crops_channels_last else torch.contiguous_format
        g = torch.Generator(device=device)
        g.manual_seed(seed + first_image)
        cl = torch.channels_last
        self.fm = [torch.empty((batch, CHANNELS, h, w), device=device, memory_format=cl) for h, w in LEVEL_HW]
        self.gt = torch.zeros((batch * GT_PER_IMAGE, 1, IMAGE, IMAGE), device=device)
        boxes, mboxes, mind = [], [], []
        for i in range(batch):                      # every image is a function of its id (see image_inputs)
            st = image_inputs(first_image + i)
            for l in range(4):
                self.fm[l][i].copy_(torch.from_numpy(st["fms"][l][0]).to(device))
            for k, (y, x, h, w) in enumerate(st["rects"]):
                self.gt[i * GT_PER_IMAGE + k, 0, y:y + h, x:x + w] = 1.0
            boxes.append(st["boxes"])
            mboxes.append(st["mboxes"])
            mind.append(st["mind"] + i * GT_PER_IMAGE)
        boxes, mboxes, mind = np.concatenate(boxes, 0), np.concatenate(mboxes, 0), np.concatenate(mind, 0).astype(np.int32)
        ind = np.repeat(np.arange(batch, dtype=np.int32), ROIS_PER_IMAGE)
        self.gfm7 = [torch.empty_like(f) for f in self.fm]
        self.gfm14 = [torch.empty_like(f) for f in self.fm]
        self.boxes_np, self.ind_np = boxes, ind
        self.boxes = torch.from_numpy(boxes).to(device)
        self.ind = torch.from_numpy(ind).to(device)
        self.N = len(boxes)
        self.out7 = torch.empty((self.N, CHANNELS, 7, 7), device=device, memory_format=cfmt)
        self.out14 = torch.empty((self.N, CHANNELS, 14, 14), device=device, memory_format=cfmt)
        self.g7 = torch.randn((self.N, CHANNELS, 7, 7), device=device, generator=g).contiguous(memory_format=cfmt)
        self.g14 = torch.randn((self.N, CHANNELS, 14, 14), device=device, generator=g).contiguous(memory_format=cfmt)
        # mask targets: gt masks [batch*G,1,1024,1024] (binary rectangles), crop 28x28 by (image, instance) index
        self.mboxes = torch.from_numpy(mboxes).to(device)
        self.mind = torch.from_numpy(mind).to(device)
        self.mt = torch.empty((len(mboxes), 1, 28, 28), device=device)
        self.Hs = _lib.i4([h for h, _ in LEVEL_HW])
        self.Ws = _lib.i4([w for _, w in LEVEL_HW])
        self.area = float(IMAGE * IMAGE)
        self.offsets = _lib.i32_array([i * ROIS_PER_IMAGE for i in range(batch + 1)])  # host: boxes grouped by image
        self.ws = torch.empty(_lib.lib.mrcnn_pyramid_roi_align_backward_workspace_bytes(self.Hs, self.Ws, batch, self.N, 14),
                              dtype=torch.uint8, device=device)
        # the planned backward: one workspace per head (both plans are alive at once) and a side stream to build them on
        self.ws7 = torch.empty(_lib.lib.mrcnn_pyramid_roi_align_backward_workspace_bytes(self.Hs, self.Ws, batch, self.N, 7),
                               dtype=torch.uint8, device=device)
        self.side = torch.cuda.Stream(device=device, priority=-1)   # its few small CTAs go first whenever an SM has room
        self.launches = 0

    def _s(self):
        return self.torch.cuda.current_stream().cuda_stream

    def fwd(self, pool, out):
        L = self.L
        L.check(L.lib.mrcnn_pyramid_roi_align_forward(L.vp4([f.data_ptr() for f in self.fm]), self.Hs, self.Ws, self.batch,
                                                      CHANNELS, L.NHWC, self.boxes.data_ptr(), self.ind.data_ptr(), self.N, pool,
                                                      self.area, out.data_ptr(), self.crop_layout, None, self._s()))
        self.launches += 1

    def fwd_pair(self):
        """Both heads' forward (7x7 into out7, 14x14 into out14) as ONE launch: mrcnn_pyramid_roi_align_forward_pair."""
        L = self.L
        L.check(L.lib.mrcnn_pyramid_roi_align_forward_pair(L.vp4([f.data_ptr() for f in self.fm]), self.Hs, self.Ws, self.batch, CHANNELS,
                                                           self.boxes.data_ptr(), self.ind.data_ptr(), self.N, self.area,
                                                           self.out7.data_ptr(), self.out14.data_ptr(), self._s()))
        self.launches += 1

    def bwd(self, pool, grad, gfm):
        L = self.L
        L.check(L.lib.mrcnn_pyramid_roi_align_backward(grad.data_ptr(), self.crop_layout, self.Hs, self.Ws, self.batch, CHANNELS,
                                                       self.boxes.data_ptr(), self.ind.data_ptr(), self.N, pool, self.area,
                                                       L.vp4([g.data_ptr() for g in gfm]), L.NHWC, 1, None, L.BWD_AUTO, self.ws.data_ptr(), self.ws.numel(), self._s()))
        self.launches += 4   # bwd_items_kernel<count>, bwd_alloc_kernel, bwd_items_kernel<fill>, roialign_bwd_gather_kernel

    def mask_targets(self):
        L = self.L
        L.check(L.lib.mrcnn_crop_forward(self.gt.data_ptr(), self.gt.shape[0], 1, IMAGE, IMAGE, L.NCHW, self.mboxes.data_ptr(),
                                         self.mind.data_ptr(), self.mt.shape[0], 0.0, 28, 28, self.mt.data_ptr(), L.NCHW, self._s()))
        self.launches += 1

    def plan(self, pool, ws, stream):
        L = self.L
        L.check(L.lib.mrcnn_pyramid_roi_align_backward_plan(self.Hs, self.Ws, self.batch, CHANNELS, self.boxes.data_ptr(), self.ind.data_ptr(),
                                                            self.N, pool, self.area, ws.data_ptr(), ws.numel(), stream.cuda_stream))
        self.launches += 3   # bwd_items_kernel<count>, bwd_alloc_kernel, bwd_items_kernel<fill>

    def bwd_planned(self, pool, grad, gfm, ws):
        L = self.L
        L.check(L.lib.mrcnn_pyramid_roi_align_backward_planned(grad.data_ptr(), self.Hs, self.Ws, self.batch, CHANNELS, self.N, pool,
                                                               L.vp4([g.data_ptr() for g in gfm]), 1, ws.data_ptr(), ws.numel(), self._s()))
        self.launches += 1   # roialign_bwd_gather_kernel

    def step_unplanned(self):
        self.fwd(7, self.out7)
        self.fwd(14, self.out14)
        self.mask_targets()
        self.bwd(14, self.g14, self.gfm14)
        self.bwd(7, self.g7, self.gfm7)

    def step_two_forwards(self):
        """The step with one forward launch per head (round 1's step)."""
        cur = self.torch.cuda.current_stream()
        self.side.wait_stream(cur)
        self.plan(14, self.ws, self.side)
        self.plan(7, self.ws7, self.side)
        ready = self.side.record_event()
        self.fwd(7, self.out7)
        self.fwd(14, self.out14)
        self.mask_targets()
        cur.wait_event(ready)
        self.bwd_planned(14, self.g14, self.gfm14, self.ws)
        self.bwd_planned(7, self.g7, self.gfm7, self.ws7)

    def step(self):
        """One training step of the RoI path.  The work-item queues of the two gather backwards depend on the boxes only:
        they are built on a side stream while the forwards run (what ops.pyramid_roi_align's autograd node does), so
        each backward on the main stream is the gather launch alone."""
        if not self.cl_crops:
            return self.step_unplanned()
        cur = self.torch.cuda.current_stream()
        self.side.wait_stream(cur)          # the previous step's gathers have finished reading the plans
        self.plan(14, self.ws, self.side)
        self.plan(7, self.ws7, self.side)
        ready = self.side.record_event()
        self.fwd_pair()                     # 7x7 and 14x14 crops of the same RoIs: one launch, the footprint read once
        self.mask_targets()
        cur.wait_event(ready)
        self.bwd_planned(14, self.g14, self.gfm14, self.ws)
        self.bwd_planned(7, self.g7, self.gfm7, self.ws7)

    # ---- per-kernel CUDA-event timings (each op alone, back to back, inputs >> L2) ----
    def time_op(self, fn, iters=20, warm=3):
        torch = self.torch
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters * 1e-3


def capture_step(torch, wl):
    """The step as ONE CUDA graph (the five ops and the side-stream queue building, fork / join included): launch gaps and
    the host's launch work leave the timed region.  Returns a callable, or None when capture is not possible."""
    try:
        for _ in range(2):
            wl.step()
        torch.cuda.synchronize()
        cap = torch.cuda.Stream()
        cap.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        before = wl.launches
        with torch.cuda.graph(graph, stream=cap):
            wl.step()
        per_step = wl.launches - before
        torch.cuda.current_stream().wait_stream(cap)

        def replay():
            graph.replay()
            wl.launches += per_step
        replay.graph = graph
        return replay
    except Exception as e:   # never cost the line: fall back to eager launches
        sys.stderr.write("CUDA graph capture of the step failed (%s: %s); timing eager launches\n" % (type(e).__name__, e))
        return None


def timed_steps(torch, dist, wl, steps, warmup, world, step=None):
    step = step or wl.step
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wl.launches = 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


# ---------------------------------------------------------------------------------------------------
def e2e_run(torch, dist, wl, steps, warmup, world, fused_backward=False):
    """Host buffers in, host buffers out, every step, through the product's host-memory API (maskrcnn_b200.hoststep.HostTrainStep:
    per image H2D -> kernels -> D2H over three streams).  fused_backward: both heads' backward as ONE gather into ONE gradient
    pyramid (what autograd accumulates in a training step), so one pyramid of gradients (1.43 GB) crosses the host link, not two."""
    from maskrcnn_b200.hoststep import HostTrainStep
    hs = HostTrainStep(wl.batch, ROIS_PER_IMAGE, CHANNELS, LEVEL_HW, (IMAGE, IMAGE), fused_backward=fused_backward, device=wl.boxes.device)
    pin = lambda t: torch.empty(t.shape, dtype=t.dtype).pin_memory()  # noqa: E731
    # host side (pinned, the device's physical order): pyramid, boxes, upstream gradients; results
    h_fm = [hs.pinned_like(f) for f in wl.fm]
    for hf, f in zip(h_fm, wl.fm):
        hf.copy_(f.permute(0, 2, 3, 1))
    h_boxes = wl.boxes.cpu().pin_memory()
    h_g7, h_g14 = hs.pinned_like(wl.g7), hs.pinned_like(wl.g14)
    h_g7.copy_(wl.g7.permute(0, 2, 3, 1))
    h_g14.copy_(wl.g14.permute(0, 2, 3, 1))
    h_out7, h_out14 = hs.pinned_like(wl.out7), hs.pinned_like(wl.out14)
    h_gfa = [hs.pinned_like(f) for f in wl.fm]
    h_gfb = None if fused_backward else [hs.pinned_like(f) for f in wl.fm]
    mask = {"d_images": wl.gt, "h_boxes": wl.mboxes.cpu().pin_memory(), "h_index": wl.mind.cpu().pin_memory(), "d_boxes": wl.mboxes,
            "d_index": wl.mind, "d_targets": wl.mt, "h_targets": pin(wl.mt)}
    h2d, d2h = hs.h2d_bytes(mask), hs.d2h_bytes(mask)

    def one_step():
        hs.run(h_fm, h_boxes, h_g7, h_g14, h_out7, h_out14, h_gfa, h_gfb, mask)

    for _ in range(max(1, min(warmup, 2))):
        one_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n = max(2, min(steps, 5))
    t0 = time.perf_counter()
    for _ in range(n):
        one_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = (time.perf_counter() - t0) / n
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return {"value": world * wl.N / dt, "unit": "RoIs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": dt * 1e3, "steps": n, "api": "maskrcnn_b200.hoststep.HostTrainStep.run",
            "note": "PCIe-bound: every input and every result crosses the host link each step; per-image 3-stream pipeline"}


# ---------------------------------------------------------------------------------------------------
# CPU side: the reference's own extension (oracle/_ref) driven like model.roi_align / CropFunction.backward
# ---------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(counter, use_ref):
    """Per-process setup, outside every timed region: import torch, take a worker number, synthesize that worker's image
    (image id = worker % BATCH: the images of the GPU arm's rank-0 batch, see image_inputs).  Imports only the host-side
    helpers of maskrcnn_b200 (synth / roofline): the CUDA library is not mapped into a CPU worker."""
    import torch
    torch.set_num_threads(1)
    from maskrcnn_b200.roofline import roi_levels
    if counter is None:
        worker = 0
    else:
        with counter.get_lock():
            worker = counter.value
            counter.value += 1
    st = image_inputs(worker % BATCH)
    st["use_ref"] = use_ref
    st["gt"] = gt_masks_of(st["rects"])
    lv = roi_levels(st["boxes"], float(IMAGE * IMAGE))
    st["lv"] = lv
    if use_ref:
        from oracle import reference
        st["C"] = reference.ref_C()
        st["tf"] = [torch.from_numpy(f) for f in st["fms"]]
        st["tgt"] = torch.from_numpy(st["gt"])
    _CPU.clear()
    _CPU.update(st)


def _cpu_worker(rois):
    """One image: the first `rois` of its 512 RoIs (all of them unless the run had to be bounded), 7x7 + 14x14 forward and
    backward, + the matching share of the 168 mask-target crops.  Returns seconds."""
    import torch
    st = _CPU
    n = int(rois)
    m = max(1, (MASK_POS * n) // ROIS_PER_IMAGE)
    boxes, lv = st["boxes"][:n], st["lv"][:n]
    if st["use_ref"]:
        from oracle import reference
        C, tf = st["C"], st["tf"]
        tb = torch.from_numpy(boxes)
        sel = [torch.from_numpy(np.nonzero(lv == l)[0]) for l in (2, 3, 4, 5)]
        tmb, tmi = torch.from_numpy(st["mboxes"][:m]), torch.from_numpy(st["mind"][:m])
        t0 = time.perf_counter()
        with reference.quiet_stdout():
            for pool in (7, 14):
                for l in range(4):          # model.py:347-377, one crop per populated level
                    if len(sel[l]) == 0:
                        continue
                    lb = tb[sel[l]]
                    ind = torch.zeros(len(lb), dtype=torch.int32)
                    crops = torch.zeros_like(tf[l])                      # __init__.py:36
                    C.crop_forward(tf[l], lb, ind, 0.0, pool, pool, crops)
                    g = torch.ones_like(crops)
                    gi = torch.zeros_like(g).resize_(*tf[l].shape)       # __init__.py:52
                    C.crop_backward(g, lb, ind, gi)
            mt = torch.zeros(1)
            C.crop_forward(st["tgt"], tmb, tmi, 0.0, 28, 28, mt)
        return time.perf_counter() - t0
    import oracle
    fms = st["fms"]
    t0 = time.perf_counter()
    for pool in (7, 14):
        out, _ = oracle.pyramid_roi_align_fwd(fms, boxes, None, pool, float(IMAGE * IMAGE))
        oracle.pyramid_roi_align_bwd(np.ones_like(out), [f.shape for f in fms], boxes, None, float(IMAGE * IMAGE))
    oracle.crop_forward(st["gt"], st["mboxes"][:m], st["mind"][:m], 28, 28, 0.0)
    return time.perf_counter() - t0


class CpuArm(object):
    """The CPU path on `procs` worker processes (image-parallel; the reference's ops are single-threaded).
    Process start-up, imports and input synthesis happen once, outside the timed region."""

    def __init__(self, procs):
        from oracle import reference
        self.use_ref = reference.ref_C_available()
        self.kind = "reference" if self.use_ref else "port"
        self.procs = max(1, procs)
        self.pool = None
        if self.procs == 1:
            _cpu_init(None, self.use_ref)
        else:
            import multiprocessing as mp
            ctx = mp.get_context("spawn")
            self.pool = ctx.Pool(self.procs, initializer=_cpu_init, initargs=(ctx.Value("i", 0), self.use_ref))
        self.measure(self.procs)                                         # warm every worker (first touch of its buffers) ...
        self.warm_s = self.measure(self.procs)[1]                        # ... then one full image each: seconds per full step

    def measure(self, images, rois=ROIS_PER_IMAGE):
        """RoIs/s (wall clock over the whole batch of images) and seconds."""
        t0 = time.perf_counter()
        if self.pool is None:
            for _ in range(images):
                _cpu_worker(rois)
        else:
            self.pool.map(_cpu_worker, [rois] * images, chunksize=1)
        dt = time.perf_counter() - t0
        return images * rois / dt, dt

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def cpu_model():
    """The host CPU's model string (stated next to every CPU number), or None."""
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return None


REF_ARM_BUDGET_S = 150.0     # the whole --impl reference run is sized to end within a few minutes


def run_reference_arm(args):
    """bench.py --impl reference: the reference's own CPU extension (oracle/_ref: vision.cpp + cpu/*.cpp compiled from
    /root/reference, unmodified) on all host cores, on the GPU arm's config, metric and images.  Runs EXACTLY --steps timed
    steps after --warmup untimed ones; a step is one image per worker process.  When steps x (seconds per full image)
    would not fit REF_ARM_BUDGET_S, every step processes the first R of each image's 512 RoIs (and the matching share of
    the mask crops) - the `sample` field says which R."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    if os.environ.get("MRCNN_BENCH_REF_PROCS"):        # tests: a small arm (the JSON says how many processes ran)
        procs = max(1, min(procs, int(os.environ["MRCNN_BENCH_REF_PROCS"])))
    per_step_images = procs                       # one image per worker per step
    arm = CpuArm(procs)                           # spawns, imports, synthesizes inputs and warms up (one full image per worker): untimed
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    budget = float(os.environ.get("MRCNN_BENCH_REF_BUDGET_S", REF_ARM_BUDGET_S))
    rois = ROIS_PER_IMAGE
    if (steps + warmup) * arm.warm_s > budget:
        rois = int(max(8, min(ROIS_PER_IMAGE, ROIS_PER_IMAGE * budget / ((steps + warmup) * arm.warm_s))))
    for _ in range(warmup):
        arm.measure(per_step_images, rois)
    vals, secs = [], []
    kind = arm.kind
    t0 = time.perf_counter()
    for _ in range(steps):
        v, dt = arm.measure(per_step_images, rois)
        vals.append(v)
        secs.append(dt)
    total = time.perf_counter() - t0
    arm.close()
    value = per_step_images * rois * steps / total
    sample = "%d images x %d of %d RoIs per step (7x7+14x14 fwd+bwd + mask crops each), image-parallel over %d processes, images " \
             "0..%d of the GPU arm's batch%s" % (per_step_images, rois, ROIS_PER_IMAGE, procs, min(procs, BATCH) - 1,
                                               "" if rois == ROIS_PER_IMAGE else
                                               "; bounded sample: the per-call zero-fill of the gradient maps is spread over fewer RoIs, "
                                               "which lowers RoIs/s (full images: %.0f RoIs/s in the warm-up)" % (procs * ROIS_PER_IMAGE / arm.warm_s))
    line = {"impl": "reference", "metric": "roialign_train_rois_per_s", "value": value, "unit": "RoIs/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": total / steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(),
            "cpu_baseline": {"value": value, "unit": "RoIs/s", "cores": procs, "kind": kind, "sample": sample, "cpu_model": cpu_model()},
            "e2e": {"value": value, "unit": "RoIs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "host_cores": cores, "step_rois_per_s_min_max": [float(min(vals)), float(max(vals))]}
    print(json.dumps(line), flush=True)


def cpu_config_baselines():
    """CPU baselines for the other BASELINE configs, on the box's host cores, rank 0 /
    N = 1 only, each a bounded sample (a few seconds).  kind "reference": the reference's own code runs - its compiled
    extension (oracle/_ref) under its unmodified Python (baseline/_ref/model.py, data.py via tools/refmodel.py):
      configs[1]  MaskRCNN.rpn_refine's steps (model.py:1336-1374) with the canonical 6000 / 1000 limits - the fork's
                  rpn_refine hard-codes 500 (model.py:1345), so the slice / sort lines are restated around the reference's
                  data.boxes_scale / boxes_refine / boxes_clamp_ and maskrcnn.nms
      configs[2]  model.roi_align (model.py:276-393), 1000 RoIs, 7x7 and 14x14
      configs[4]  MaskRCNN.mrn_refine (model.py:1389-1487) + model.roi_align(..., 14) of its detections, per image
    Falls back to the oracle port (kind "port") where the reference's Python did not travel."""
    import torch
    from maskrcnn_b200 import synth
    out = {}
    threads = torch.get_num_threads()
    host = {"cores_native_ops": 1, "torch_threads": threads, "host_cores": os.cpu_count() or 1, "cpu_model": cpu_model()}
    ref = None
    try:
        from oracle import reference
        from tools import refmodel
        if reference.ref_C_available() and refmodel.available():
            ref = refmodel.load(reference.make_maskrcnn_shim())
            quiet = reference.quiet_stdout
    except Exception:
        ref = None
    import types
    anchors = synth.pyramid_anchors((IMAGE, IMAGE))
    # ---- configs[1]
    rcs, rbs = zip(*[synth.rpn_outputs(anchors, 1235 + i) for i in range(2)])
    if ref is not None:
        ta = torch.from_numpy(anchors)
        std = np.array([0.1, 0.1, 0.2, 0.2])

        def rpn_refine_6000(rc, rb):
            scores = rc[:, 1]                                                                  # model.py:1336
            deltas = ref.data.boxes_scale(rb, std)                                             # :1341
            scores, order = scores.sort(descending=True)                                       # :1346
            order, scores = order[:6000], scores[:6000]                                        # :1347-1350
            boxes = ref.data.boxes_refine(ta[order, :], deltas[order, :])                      # :1354
            ref.data.boxes_clamp_(boxes, [0, 0, IMAGE, IMAGE])                                 # :1358 (in place)
            keep = ref.maskrcnn.nms(torch.cat((boxes, scores.unsqueeze(1)), 1), 0.7)[:1000]    # :1364-1366
            return boxes[keep, :] / torch.tensor([IMAGE, IMAGE, IMAGE, IMAGE], dtype=torch.float32)   # :1371-1374
        t0 = time.perf_counter()
        n = 0
        while n < 4 or time.perf_counter() - t0 < 3.0:
            rpn_refine_6000(torch.from_numpy(rcs[n % 2]), torch.from_numpy(rbs[n % 2]))
            n += 1
        dt = (time.perf_counter() - t0) / n
        kind = "reference"
    else:
        import oracle
        t0 = time.perf_counter()
        n = 0
        while n < 4 or time.perf_counter() - t0 < 3.0:
            oracle.proposal_layer(rcs[n % 2], rbs[n % 2], anchors, 6000, 1000, 0.7)
            n += 1
        dt = (time.perf_counter() - t0) / n
        kind = "port"
    out["configs[1]"] = dict(host, value=1.0 / dt, unit="images/s", kind=kind, sample="%d images, 261,888 anchors, 6000 -> NMS 0.7 -> 1000, one at a time" % n)
    # ---- configs[2] and configs[4]
    rng = np.random.default_rng(SEED)
    fms = [rng.standard_normal((1, CHANNELS, h, w), dtype=np.float32) for h, w in LEVEL_HW]
    boxes = synth.random_rois(1000, 1234)
    N, NC, D = 1000, 81, 100
    if ref is not None:
        tf = [torch.from_numpy(f) for f in fms]
        for pool in (7, 14):
            t0 = time.perf_counter()
            with quiet(), torch.no_grad():
                ref.model.roi_align([torch.from_numpy(boxes).unsqueeze(0)] + list(tf), pool, [IMAGE, IMAGE, 3])
            dt = time.perf_counter() - t0
            out["configs[2] %dx%d" % (pool, pool)] = dict(host, value=1000 / dt, unit="RoIs/s", kind="reference",
                                                         sample="model.roi_align, 1000 RoIs x 256 ch, one call, %.2f s" % dt)
        cfg = type("C", (ref.config.CocoInferenceConfig,), {"GPU_COUNT": 0, "DETECTION_MAX_INSTANCES": D})()
        me = types.SimpleNamespace(config=cfg)
        t0 = time.perf_counter()
        n = 0
        while n < 2 or time.perf_counter() - t0 < 3.0:
            pr, de = synth.head_outputs(N, NC, 700 + n)
            rois = torch.from_numpy(synth.random_rois(N, 300 + n)).unsqueeze(0)
            with quiet(), torch.no_grad():
                cls_, sc_, bx_ = ref.model.MaskRCNN.mrn_refine(me, rois, torch.from_numpy(pr), torch.from_numpy(de), (0, 0, IMAGE, IMAGE))
                if cls_ is not None:
                    ref.model.roi_align([bx_.float() / IMAGE] + list(tf), 14, [IMAGE, IMAGE, 3])   # model.py:1188-1189 -> :889
            n += 1
        dt = (time.perf_counter() - t0) / n
        out["configs[4]"] = dict(host, value=1.0 / dt, unit="images/s", kind="reference",
                                 sample="%d images: MaskRCNN.mrn_refine (1000 RoIs, 81 classes, top-100) + model.roi_align(14) of the detections" % n)
    else:
        import oracle
        for pool in (7, 14):
            t0 = time.perf_counter()
            oracle.pyramid_roi_align_fwd(fms, boxes, None, pool, float(IMAGE * IMAGE))
            dt = time.perf_counter() - t0
            out["configs[2] %dx%d" % (pool, pool)] = dict(host, value=1000 / dt, unit="RoIs/s", kind="port", sample="1000 RoIs x 256 ch, one call")
        t0 = time.perf_counter()
        n = 0
        while n < 2 or time.perf_counter() - t0 < 3.0:
            pr, de = synth.head_outputs(N, NC, 700 + n)
            d = oracle.detection_layer(synth.random_rois(N, 300 + n), pr, de, np.array([0, 0, IMAGE, IMAGE], np.float32), 0.0, 0.3, D)
            oracle.pyramid_roi_align_fwd(fms, d[:, :4] / np.float32(IMAGE), None, 14, float(IMAGE * IMAGE))
            n += 1
        dt = (time.perf_counter() - t0) / n
        out["configs[4]"] = dict(host, value=1.0 / dt, unit="images/s", kind="port", sample="%d images" % n)
    return out


def predict_flow(torch):
    """BASELINE configs[0]: predict.py's flow (predict.py:43-60 -> MaskRCNN.detect) on images/car58a54312d.jpg with a
    random-init ResNet-101-FPN: the reference's UNMODIFIED model.py (baseline/_ref) on the B200 with this repo's drop-in,
    (a) package only, (b) patch() + channels_last; and the same flow on the host CPU with the reference's own extension.
    Seconds per image (median of 5 after 2 warm-ups), with the time inside the RoI-path operators measured by CUDA events."""
    from tools import refmodel
    if not refmodel.available() or refmodel.image_path() is None:
        return {"unavailable": "baseline/_ref (the reference's model.py + demo image) did not travel"}
    import maskrcnn as product
    import maskrcnn_b200 as m
    img = refmodel.pil_imread(refmodel.image_path())
    out = {"config": "configs[0]: model.detect on images/car58a54312d.jpg (1920x1200 -> 1024x1024), random-init ResNet-101-FPN, seed 2026"}

    def timed(fn, spans):
        def wrapped(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            spans.append((e0, e1))
            return r
        return wrapped

    for mode in ("package_only", "patched_channels_last"):
        ref = refmodel.load(product)
        spans = []
        if mode == "patched_channels_last":
            m.patch(ref.model, ref.data)
            M = ref.model.MaskRCNN
            ref.model.roi_align = timed(ref.model.roi_align, spans)
            for name in ("rpn_refine", "mrn_refine", "rpn_detect"):
                f = getattr(M, name)
                setattr(M, name, (lambda f: lambda self, *a: timed(lambda *b: f(self, *b), spans)(*a))(f))
            ref.data.full_masks = timed(ref.data.full_masks, spans)
            ref.data.decode_masks = timed(ref.data.decode_masks, spans)
        else:
            refmodel.tolerate_empty_boxes(ref)
            ref.model.roi_align = timed(ref.model.roi_align, spans)
            M = ref.model.MaskRCNN
            for name in ("rpn_refine", "mrn_refine"):
                f = getattr(M, name)
                setattr(M, name, (lambda f: lambda self, *a: timed(lambda *b: f(self, *b), spans)(*a))(f))
        cfg = refmodel.make_config(ref, gpu=True)
        model = refmodel.make_model(ref, cfg, 2026)
        if mode == "patched_channels_last":
            model = model.to(memory_format=torch.channels_last)
        pspans = []
        model.predict = timed(model.predict, pspans)           # detect = resize/mold on the host + predict + decode + .tolist()
        ts, roi, pred = [], [], []
        n_det = None
        try:
            with torch.no_grad():
                for it in range(7):
                    del spans[:]
                    del pspans[:]
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    res = model.detect(img)
                    torch.cuda.synchronize()
                    ts.append(time.perf_counter() - t0)
                    roi.append(sum(a.elapsed_time(b) for a, b in spans) * 1e-3)
                    pred.append(sum(a.elapsed_time(b) for a, b in pspans) * 1e-3)
                    n_det = None if res[0] is None else len(res[0])
            k = int(np.argsort(ts[2:])[len(ts[2:]) // 2]) + 2
            out[mode] = {"s_per_image": ts[k], "s_in_model_predict": pred[k], "s_in_roi_path_ops": roi[k], "detections": n_det,
                         "host_note": "s_per_image is the reference's whole detect(): it ends with mrn_masks.cpu().tolist() of [D,1200,1920] "
                                      "uint8 masks (model.py:1136), ~1 s of Python list building that is neither network nor RoI path; "
                                      "s_in_model_predict is FPN + RPN + heads + RoI path + full_masks on the device (CUDA events)",
                         "note": ("roi_align, rpn_refine, mrn_refine as the reference's own Python over maskrcnn.nms / CropFunction"
                                  if mode == "package_only" else
                                  "fused rpn_detect, rpn_refine, roi_align x2, mrn_refine, full_masks, decode_masks")}
        except Exception as e:   # an extra: it must never cost the line
            out[mode] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        del model
        torch.cuda.empty_cache()
    try:
        from oracle import reference
        if reference.ref_C_available():
            ref = refmodel.load(reference.make_maskrcnn_shim())
            refmodel.tolerate_empty_boxes(ref)
            model = refmodel.make_model(ref, refmodel.make_config(ref, gpu=False), 2026)
            with torch.no_grad(), reference.quiet_stdout():
                t0 = time.perf_counter()
                res = model.detect(img)
                dt = time.perf_counter() - t0
            out["cpu_reference"] = {"s_per_image": dt, "detections": None if res[0] is None else len(res[0]), "kind": "reference",
                                    "torch_threads": torch.get_num_threads(), "host_cores": os.cpu_count(), "cpu_model": cpu_model(),
                                    "sample": "one image, convolutions on all torch threads, c++ext ops single-threaded as shipped"}
    except Exception as e:
        out["cpu_reference"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    return out


def bind_to_gpu_numa_node(local):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the end-to-end
    leg are first-touched on the GPU's own NUMA node (8 ranks sharing one node's memory channels cap the host link)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:   # CUDA_VISIBLE_DEVICES may renumber the devices: go through the PCI address
            import torch
            pr = torch.cuda.get_device_properties(local)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))
    except Exception:  # no NVML / not permitted: keep the inherited affinity
        return None


def ncu_traffic():
    """DRAM bytes per launch of each timed op, from the newest committed ncu summary (profiles/rNN_traffic.json):
    {capture name: MB summed over the kernels the op launches}."""
    import glob
    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r*_traffic.json")))
    if not files:
        return {}
    out = {}
    for key, rec in json.load(open(files[-1])).items():
        cap = key.split(":", 1)[0]
        out[cap] = out.get(cap, 0.0) + rec["dram_bytes"] / 1e6
        if "roialign_bwd_gather_kernel" in key:      # the gather launch alone (the planned backward)
            out[cap + ":gather"] = rec["dram_bytes"] / 1e6
    return out


def mask_target_bytes(wl):
    """Algorithmic bytes of the 28x28 mask-target crop: outputs written once + the UNIQUE mask pixels the taps touch (4 B each)
    + boxes and indices.  (Round 1 credited five floats per output; taps of neighbouring bins coincide for small boxes.)"""
    mb, mi = wl.mboxes.cpu().numpy(), wl.mind.cpu().numpy()
    sm1 = np.float32(IMAGE - 1)
    i = np.arange(28, dtype=np.float32)
    unique = 0
    for k in range(len(mb)):
        taps = []
        for a1, a2 in ((mb[k, 0], mb[k, 2]), (mb[k, 1], mb[k, 3])):
            pos = (a1 * sm1 + i * (((a2 - a1) * sm1) / np.float32(27))).astype(np.float32)
            ok = (pos >= 0) & (pos <= sm1)
            taps.append(len(np.unique(np.concatenate([np.floor(pos[ok]), np.ceil(pos[ok])]))))
        unique += taps[0] * taps[1]
    return wl.mt.numel() * 4 + unique * 4 + len(mb) * 20


def mask_target_sector_bytes(wl):
    """The same crop counted in the unit DRAM moves: the UNIQUE 32-byte sectors the taps touch (a mask row starts on a 4 KB
    boundary, so a pixel x lies in sector x // 8 of its row) + the outputs.  The taps of a 28-bin crop of a box wider than ~220
    pixels are more than 8 pixels apart, so every tap pair drags a whole sector in: ~3x the unique-pixel bytes."""
    mb = wl.mboxes.cpu().numpy()
    sm1 = np.float32(IMAGE - 1)
    i = np.arange(28, dtype=np.float32)
    sectors = 0
    for k in range(len(mb)):
        def taps(a1, a2):
            pos = (a1 * sm1 + i * (((a2 - a1) * sm1) / np.float32(27))).astype(np.float32)
            ok = (pos >= 0) & (pos <= sm1)
            return np.unique(np.concatenate([np.floor(pos[ok]), np.ceil(pos[ok])]))
        rows, cols = taps(mb[k, 0], mb[k, 2]), taps(mb[k, 1], mb[k, 3])
        sectors += len(rows) * len(np.unique(cols // 8))
    return wl.mt.numel() * 4 + sectors * 32 + len(mb) * 20


def workload_config():
    return {"workload": "BASELINE configs[3]: training-mode PyramidROIAlign fwd+bwd, batch %d x %d RoIs x %d ch, P2-P5 of %dx%d, "
                        "7x7 + 14x14 + %d 28x28 mask-target crops/img" % (BATCH, ROIS_PER_IMAGE, CHANNELS, IMAGE, IMAGE, MASK_POS),
            "batch_per_gpu": BATCH, "rois_per_image": ROIS_PER_IMAGE, "channels": CHANNELS,
            "feature_layout": "channels_last (NHWC) pyramid, crops and gradients (logical shapes are the reference's [N,C,h,w])", "parallelism": "image-sharded replicas, no collective",
            "l2": "inputs (1.43 GB pyramid + 2 GB crops per step) are larger than L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------------
def secondary(torch, wl, hbm):
    """The other BASELINE configs, each timed alone with CUDA events (N=1, rank 0)."""
    import maskrcnn_b200 as m
    from maskrcnn_b200 import synth, roofline
    out = {}
    dev = "cuda"
    L = wl.L
    # both heads' backward fused into one gradient pyramid (what autograd accumulates in the reference's training step)
    ws2 = torch.empty(L.lib.mrcnn_pyramid_roi_align_backward_pair_workspace_bytes(wl.Hs, wl.Ws, wl.batch, wl.N, 7, 14), dtype=torch.uint8,
                      device=dev)

    def bwd_pair():
        L.check(L.lib.mrcnn_pyramid_roi_align_backward_pair(wl.g7.data_ptr(), 7, wl.g14.data_ptr(), 14, wl.Hs, wl.Ws, wl.batch, CHANNELS,
                                                            wl.boxes.data_ptr(), wl.ind.data_ptr(), wl.N, wl.area,
                                                            L.vp4([x.data_ptr() for x in wl.gfm14]), 1, ws2.data_ptr(), ws2.numel(), wl._s()))

    def step_fused():
        wl.fwd_pair()
        wl.mask_targets()
        bwd_pair()
    if wl.cl_crops:
        tb = wl.time_op(bwd_pair, iters=20)
        ts = wl.time_op(step_fused, iters=20)
        by = (wl.g7.numel() + wl.g14.numel() + wl.batch * PYR_ELEMS_PER_IMAGE) * 4
        out["train_step_fused_backward"] = {
            "config": "configs[3] with ONE forward launch and ONE backward for both heads: d/dP2..P5 = 7x7 head + 14x14 head, written once "
                      "(mrcnn_pyramid_roi_align_forward_pair + mrcnn_pyramid_roi_align_backward_pair = what ops.pyramid_roi_align_pair runs)", "rois_per_s": wl.N / ts, "ms_per_step": ts * 1e3,
            "backward_pair_ms": tb * 1e3, "backward_pair_algorithmic_MB": by / 1e6, "backward_pair_frac_of_hbm": by / tb / 1e9 / hbm,
            "note": "not the headline: the headline step returns the two heads' gradient pyramids separately, like the reference ops"}
    del ws2
    # detection-target layer (mrn_samples) between the proposal layer and the training RoIAlign
    Bt, Nt, Gt, Tt = 16, 1000, 20, ROIS_PER_IMAGE
    ins = [synth.target_inputs(Nt, Gt, 900 + b, image=IMAGE, n_crowd=b % 2, n_pad=2) for b in range(2)]   # two distinct images, tiled
    stack = lambda k: torch.from_numpy(np.stack([ins[b % 2][k] for b in range(Bt)])).to(dev)  # noqa: E731
    t_rois, t_cls, t_gt, t_masks = stack(0), stack(1), stack(2), stack(3)
    g_ = torch.Generator(device=dev)
    g_.manual_seed(5)
    kp, kn = torch.rand((Bt, Nt), device=dev, generator=g_), torch.rand((Bt, Nt), device=dev, generator=g_)
    f = lambda: m.detection_targets(t_rois, t_cls, t_gt, t_masks, kp, kn, train_rois_per_image=Tt)  # noqa: E731
    t = wl.time_op(f, iters=20)
    take = f()[4].float().mean(0)
    by = Bt * (Nt * 16 + Gt * 20 + Tt * (16 + 4 + 16 + 28 * 28 * 4)) + float(take[0]) * Bt * 28 * 28 * 16
    out["detection_targets"] = {"config": "mrn_samples batched (SURVEY 8f): %d images x %d proposals x %d gt, %dx%d gt masks, %d RoIs + 28x28 "
                                          "mask targets per image, sync-free (random keys)" % (Bt, Nt, Gt, IMAGE, IMAGE, Tt),
                                "images_per_s": Bt / t, "ms_per_batch": t * 1e3, "kept_pos_mean": float(take[0]), "kept_neg_mean": float(take[1]),
                                "algorithmic_GBps": by / t / 1e9, "frac_of_hbm": by / t / 1e9 / hbm,
                                "note": "3 launches (classify, select, emit); latency-bound: ~4 MB per image"}
    del t_masks
    # RPN anchor matching (data.rpn_samples), 261,888 anchors x 20 gt boxes
    import types
    anc64 = torch.from_numpy(synth.pyramid_anchors((IMAGE, IMAGE)).astype(np.float64)).to(dev)
    r_cls, r_gt = synth.rpn_target_inputs(20, 77, image=IMAGE, n_crowd=1)
    r_cls_d, r_gt_d = torch.from_numpy(r_cls).to(dev), torch.from_numpy(r_gt).to(dev)
    rcfg = types.SimpleNamespace(RPN_TRAIN_ANCHORS_PER_IMAGE=256, RPN_BBOX_STD_DEV=np.array([0.1, 0.1, 0.2, 0.2]))
    A_ = anc64.size(0)
    mt_, am_ = torch.empty(A_, dtype=torch.int32, device=dev), torch.empty(A_, dtype=torch.int32, device=dev)
    wsr = torch.empty(L.lib.mrcnn_rpn_match_workspace_bytes(20), dtype=torch.uint8, device=dev)
    fk = lambda: L.check(L.lib.mrcnn_rpn_match(anc64.data_ptr(), A_, r_gt_d.data_ptr(), r_cls_d.data_ptr(), 20, mt_.data_ptr(),  # noqa: E731
                                               am_.data_ptr(), wsr.data_ptr(), wsr.numel(), wl._s()))
    tk = wl.time_op(fk, iters=20)
    np.random.seed(0)
    t0 = time.perf_counter()
    for _ in range(10):
        m.rpn_samples(anc64, r_cls_d, r_gt_d, rcfg)
    torch.cuda.synchronize()
    td = (time.perf_counter() - t0) / 10
    out["rpn_anchor_matching"] = {"config": "data.rpn_samples (SURVEY 8f): %d anchors x 20 gt boxes, 256 anchors kept" % A_,
                                  "match_kernels_us": tk * 1e6, "match_algorithmic_GBps": (A_ * 40 + 20 * 20) / tk / 1e9,
                                  "dropin_images_per_s": 1.0 / td, "dropin_ms": td * 1e3,
                                  "note": "drop-in = reference semantics incl. three count read-backs and np.random draws on the host"}
    # mask paste-back (data.full_masks) - 8 images x 100 detections -> 800 boolean 1024x1024 masks
    Dm = 800
    m_cls, m_boxes, m_masks = synth.mask_head_outputs(100, 81, 41, image=IMAGE)
    m_cls_d = torch.from_numpy(np.tile(m_cls, Dm // 100)).to(dev)
    m_boxes_d = torch.from_numpy(np.tile(m_boxes, (Dm // 100, 1))).to(dev)
    m_masks_d = torch.from_numpy(m_masks).to(dev).repeat(Dm // 100, 1, 1, 1)
    m_out = torch.empty((Dm, IMAGE, IMAGE), dtype=torch.bool, device=dev)
    m_ws = torch.empty(L.lib.mrcnn_full_masks_workspace_bytes(Dm, 28, 28, IMAGE, IMAGE), dtype=torch.uint8, device=dev)
    fpaste = lambda: L.check(L.lib.mrcnn_full_masks(m_cls_d.data_ptr(), m_boxes_d.data_ptr(), m_masks_d.data_ptr(), Dm, 81, 28, 28, IMAGE,  # noqa: E731
                                                    IMAGE, m_out.data_ptr(), m_ws.data_ptr(), m_ws.numel(), wl._s()))
    t = wl.time_op(fpaste, iters=20)
    t_fill = wl.time_op(lambda: m_out.zero_(), iters=20)   # write-only ceiling: a plain fill of the same buffer
    by = Dm * (IMAGE * IMAGE + 28 * 28 * 4 + 24)
    out["full_masks"] = {"config": "data.full_masks (SURVEY 8f): %d detections (8 images x 100) x 81 classes x 28x28 -> bool [%d,%d,%d], two launches"
                                   % (Dm, Dm, IMAGE, IMAGE), "detections_per_s": Dm / t, "images_per_s": Dm / 100 / t, "ms": t * 1e3,
                         "algorithmic_MB": by / 1e6, "algorithmic_GBps": by / t / 1e9, "frac_of_hbm": by / t / 1e9 / hbm,
                         "box_area_fraction": float(((m_boxes[:, 2] - m_boxes[:, 0]) * (m_boxes[:, 3] - m_boxes[:, 1])).mean() / IMAGE / IMAGE),
                         "fill_of_the_same_buffer_ms": t_fill * 1e3, "fill_GBps": Dm * IMAGE * IMAGE / t_fill / 1e9,
                         "frac_of_write_only_fill": t_fill / t,
                         "note": "write-bound: H*W bytes per detection written once with 128-bit streaming stores; a write-only stream "
                                 "tops out well below the copy bandwidth (see the plain fill timed beside it)"}
    # ... and its downstream half, data.decode_masks: the predict.py frame (1920 x 1200 -> scale 1024 / 1920, window rows
    # 192..832): 100 of the pasted masks cropped to the window and resized to 1200 x 1920, uint8
    Dd, dch, dcw, dnh, dnw = 100, 640, IMAGE, 1200, 1920
    d_out = torch.empty((Dd, dnh, dnw), dtype=torch.uint8, device=dev)
    d_ws = torch.empty(L.lib.mrcnn_decode_masks_workspace_bytes(dch, dcw, dnh, dnw), dtype=torch.uint8, device=dev)
    fdec = lambda: L.check(L.lib.mrcnn_decode_masks(m_out.data_ptr(), 1, Dd, IMAGE, IMAGE, (IMAGE - dch) // 2, 0, dch, dcw, dnh, dnw,  # noqa: E731
                                                    d_out.data_ptr(), d_ws.data_ptr(), d_ws.numel(), wl._s()))
    fpaste()
    t = wl.time_op(fdec, iters=20)
    by = Dd * (dch * dcw + dnh * dnw)
    out["decode_masks"] = {"config": "data.decode_masks (SURVEY 8f): %d masks, window %dx%d of 1024x1024 -> uint8 [%d,%d,%d] (the predict.py "
                                     "frame), two launches" % (Dd, dch, dcw, Dd, dnh, dnw), "masks_per_s": Dd / t, "ms": t * 1e3,
                           "algorithmic_MB": by / 1e6, "algorithmic_GBps": by / t / 1e9, "frac_of_hbm": by / t / 1e9 / hbm,
                           "note": "window read once + output written once; Pillow's two-pass 8-bit resample with the horizontal pass "
                                   "staged in shared memory"}
    del m_out, m_masks_d, m_ws, d_out, d_ws
    # RPN head output plumbing (rpn_detect) - conv outputs of P2..P6 -> [B,A,2] / [B,A,4] / fg [B,A], batch 8
    Bp = 8
    g_ = torch.Generator(device=dev)
    g_.manual_seed(9)
    sides = [IMAGE // s_ for s_ in (4, 8, 16, 32, 64)]
    cls_l = [torch.randn((Bp, 6, s_, s_), device=dev, generator=g_) for s_ in sides]
    box_l = [torch.randn((Bp, 12, s_, s_), device=dev, generator=g_) for s_ in sides]
    with torch.no_grad():
        import ctypes
        A_all = 3 * sum(s_ * s_ for s_ in sides)
        po = [torch.empty((Bp, A_all, k_), device=dev) for k_ in (2, 2, 4, 1)]
        lp = (ctypes.c_void_p * 5)(*[x.data_ptr() for x in cls_l])
        bp = (ctypes.c_void_p * 5)(*[x.data_ptr() for x in box_l])
        hs = (ctypes.c_int * 5)(*sides)
        fpack = lambda: L.check(L.lib.mrcnn_rpn_pack(lp, bp, hs, hs, 5, Bp, 3, L.NCHW, po[0].data_ptr(), po[1].data_ptr(),  # noqa: E731
                                                     po[2].data_ptr(), po[3].data_ptr(), wl._s()))
        t = wl.time_op(fpack, iters=20)
        t_api = wl.time_op(lambda: m.rpn_pack(cls_l, box_l), iters=20)

        def torch_chain():   # the reference's launches (model.py:624-641, :1294-1304) on the same device, for scale
            lg = [x.permute(0, 2, 3, 1).contiguous().view(Bp, -1, 2) for x in cls_l]
            pr = [torch.softmax(x, 2) for x in lg]
            bb = [x.permute(0, 2, 3, 1).contiguous().view(Bp, -1, 4) for x in box_l]
            return torch.cat(lg, 1), torch.cat(pr, 1), torch.cat(bb, 1)
        t_ref = wl.time_op(torch_chain, iters=20)
    by = Bp * A_all * (24 + 36)   # 6 floats read, 2 + 2 + 4 + 1 floats written per anchor
    out["rpn_pack"] = {"config": "MaskRCNN.rpn_detect plumbing (SURVEY 8f): P2..P6 conv outputs of %d images -> logits/probs [B,%d,2], deltas "
                                 "[B,%d,4], fg [B,%d]; one launch" % (Bp, A_all, A_all, A_all),
                       "images_per_s": Bp / t, "us": t * 1e6, "algorithmic_MB": by / 1e6, "algorithmic_GBps": by / t / 1e9,
                       "frac_of_hbm": by / t / 1e9 / hbm, "through_ops_rpn_pack_us": t_api * 1e6, "same_ops_stock_pytorch_us": t_ref * 1e6,
                       "note": "stock PyTorch chain = 5 x (2 permute copies + softmax) + 3 cat on the same GPU"}
    anchors_d = torch.from_numpy(synth.pyramid_anchors((IMAGE, IMAGE))).to(dev)
    rc_, rb_ = zip(*[synth.rpn_outputs(synth.pyramid_anchors((IMAGE, IMAGE)), 500 + i) for i in range(2)])
    rc_d = torch.from_numpy(np.stack([rc_[i % 2] for i in range(Bp)])).to(dev)
    rb_d = torch.from_numpy(np.stack([rb_[i % 2] for i in range(Bp)])).to(dev)
    fg_d = rc_d[:, :, 1].contiguous()
    t2 = wl.time_op(lambda: m.proposal_layer(rc_d, rb_d, anchors_d, 6000, 1000, 0.7), iters=20)
    t1 = wl.time_op(lambda: m.proposal_layer(fg_d, rb_d, anchors_d, 6000, 1000, 0.7), iters=20)
    out["proposal_layer_fg_scores"] = {"config": "configs[1] fed with fg probabilities [B,A] (rpn_pack's fg output) instead of rpn_class [B,A,2]",
                                       "ms_per_batch_fg": t1 * 1e3, "ms_per_batch_pair": t2 * 1e3, "images_per_s_fg": Bp / t1}
    del cls_l, box_l, rc_d, rb_d, po
    # configs[2]: forward only, 1000 RoIs x 256 ch on one image
    boxes_np = synth.random_rois(1000, 1234)
    boxes = torch.from_numpy(boxes_np).to(dev)
    fm1 = [f_[:1] for f_ in wl.fm]
    for pool in (7, 14):
        inputs1 = [boxes.unsqueeze(0)] + list(fm1)
        g = lambda: m.roi_align(inputs1, pool, [IMAGE, IMAGE, 3])  # noqa: E731   (the reference's call, model.py:276)
        t = wl.time_op(g, iters=50)
        g_nchw = lambda: m.pyramid_roi_align(fm1, boxes, None, pool, (IMAGE, IMAGE, 3), out_channels_last=False)  # noqa: E731
        t_nchw = wl.time_op(g_nchw, iters=50)
        U, _ = roofline.unique_taps(boxes_np, None, pool, (IMAGE, IMAGE), LEVEL_HW, 1)
        by = roofline.roialign_fwd_bytes(1000, CHANNELS, pool, U)
        out["roialign_fwd_%dx%d" % (pool, pool)] = {"config": "configs[2]: 1000 RoIs x 256 ch, one image (warm L2: 89 MB pyramid fits)",
                                                    "rois_per_s": 1000 / t, "us": t * 1e6, "algorithmic_MB": by / 1e6,
                                                    "algorithmic_GBps": by / t / 1e9, "frac_of_hbm": by / t / 1e9 / hbm,
                                                    "through_ops_nchw_crops": {"us": t_nchw * 1e6, "rois_per_s": 1000 / t_nchw,
                                                                               "frac_of_hbm": by / t_nchw / 1e9 / hbm},
                                                    "note": "through the drop-in ops.roi_align(inputs, pool, image_shape) on a channels-last "
                                                            "pyramid: the crops follow the pyramid's memory format (channels-last, what the "
                                                            "heads' cuDNN convolutions take natively); `through_ops_nchw_crops` forces "
                                                            "NCHW-contiguous crops (out_channels_last=False), `direct_abi` is the bare C call"}
        # the same launch straight through the C ABI into a preallocated output (what wl.fwd does for the headline), for
        # both crop layouts: takes the host-side call overhead out of a ~20 us kernel.  Guarded: an extra.
        try:
            La = wl.L
            ptrs = La.vp4([f_.data_ptr() for f_ in fm1])
            direct = {}
            for lname, lay, mf in (("nchw_crops", La.NCHW, torch.contiguous_format), ("channels_last_crops", La.NHWC, torch.channels_last)):
                o_ = torch.empty((1000, CHANNELS, pool, pool), device=dev, memory_format=mf)

                def launch(o_=o_, lay=lay, pool=pool):
                    La.check(La.lib.mrcnn_pyramid_roi_align_forward(ptrs, wl.Hs, wl.Ws, 1, CHANNELS, La.NHWC, boxes.data_ptr(), None, 1000,
                                                                    pool, wl.area, o_.data_ptr(), lay, None, wl._s()))
                td = wl.time_op(launch, iters=50)
                direct[lname] = {"us": td * 1e6, "rois_per_s": 1000 / td, "algorithmic_GBps": by / td / 1e9, "frac_of_hbm": by / td / 1e9 / hbm}
            out["roialign_fwd_%dx%d" % (pool, pool)]["direct_abi"] = direct
        except Exception as e:
            out["roialign_fwd_%dx%d" % (pool, pool)]["direct_abi"] = {"error": "%s: %s" % (type(e).__name__, e)}
    # the same training step with NCHW-contiguous crops and upstream gradients (the reference's physical layout;
    # the kernels then transpose through shared memory)
    o7, o14 = wl.out7.contiguous(), wl.out14.contiguous()
    g7, g14 = wl.g7.contiguous(), wl.g14.contiguous()
    L = wl.L
    def step_nchw():
        for pool, o in ((7, o7), (14, o14)):
            L.check(L.lib.mrcnn_pyramid_roi_align_forward(L.vp4([f.data_ptr() for f in wl.fm]), wl.Hs, wl.Ws, wl.batch, CHANNELS, L.NHWC,
                                                          wl.boxes.data_ptr(), wl.ind.data_ptr(), wl.N, pool, wl.area, o.data_ptr(), L.NCHW,
                                                          None, wl._s()))
        wl.mask_targets()
        for pool, g, gf in ((14, g14, wl.gfm14), (7, g7, wl.gfm7)):
            L.check(L.lib.mrcnn_pyramid_roi_align_backward(g.data_ptr(), L.NCHW, wl.Hs, wl.Ws, wl.batch, CHANNELS, wl.boxes.data_ptr(),
                                                           wl.ind.data_ptr(), wl.N, pool, wl.area, L.vp4([x.data_ptr() for x in gf]),
                                                           L.NHWC, 1, None, L.BWD_AUTO, wl.ws.data_ptr(), wl.ws.numel(), wl._s()))
    t = wl.time_op(step_nchw, iters=20)
    out["train_step_nchw_crops"] = {"config": "configs[3] with NCHW-contiguous crops and gradients (smem-transposed)", "rois_per_s": wl.N / t,
                                    "ms_per_step": t * 1e3}
    del o7, o14, g7, g14
    # NCHW feature pyramids - what an UNMODIFIED model.py produces (no channels_last): the literal drop-in's layout.
    try:
        fm_nchw = [f_.contiguous() for f_ in wl.fm]                       # [16,256,H,W] NCHW copies of the same pyramid
        nchw = {}
        for pool in (7, 14):
            fm1n = [f_[:1] for f_ in fm_nchw]
            U, _ = roofline.unique_taps(boxes_np, None, pool, (IMAGE, IMAGE), LEVEL_HW, 1)
            by = roofline.roialign_fwd_bytes(1000, CHANNELS, pool, U)
            o_ = torch.empty((1000, CHANNELS, pool, pool), device=dev)
            pt = L.vp4([f_.data_ptr() for f_ in fm1n])
            td = wl.time_op(lambda: L.check(L.lib.mrcnn_pyramid_roi_align_forward(pt, wl.Hs, wl.Ws, 1, CHANNELS, L.NCHW, boxes.data_ptr(), None,
                                                                                  1000, pool, wl.area, o_.data_ptr(), L.NCHW, None, wl._s())), iters=50)
            ta = wl.time_op(lambda: m.pyramid_roi_align(fm1n, boxes, None, pool, (IMAGE, IMAGE, 3)), iters=50)
            nchw["fwd_%dx%d_1000_rois" % (pool, pool)] = {"direct_abi_us": td * 1e6, "through_ops_us": ta * 1e6, "algorithmic_MB": by / 1e6,
                                                          "frac_of_hbm_direct": by / td / 1e9 / hbm, "frac_of_hbm_ops": by / ta / 1e9 / hbm}
        # configs[3] on the NCHW pyramid: NCHW crops and gradients too (the reference's layouts end to end)
        o7, o14 = torch.empty((wl.N, CHANNELS, 7, 7), device=dev), torch.empty((wl.N, CHANNELS, 14, 14), device=dev)
        g7, g14 = wl.g7.contiguous(), wl.g14.contiguous()
        gfn = [torch.empty_like(f_) for f_ in fm_nchw]
        pf, pg = L.vp4([f_.data_ptr() for f_ in fm_nchw]), L.vp4([f_.data_ptr() for f_ in gfn])

        def fwd_n(pool, o):
            L.check(L.lib.mrcnn_pyramid_roi_align_forward(pf, wl.Hs, wl.Ws, wl.batch, CHANNELS, L.NCHW, wl.boxes.data_ptr(), wl.ind.data_ptr(),
                                                          wl.N, pool, wl.area, o.data_ptr(), L.NCHW, None, wl._s()))

        ws_n = torch.empty(L.lib.mrcnn_pyramid_roi_align_backward_workspace_bytes_ex(wl.Hs, wl.Ws, wl.batch, CHANNELS, wl.N, 14, L.NCHW),
                           dtype=torch.uint8, device=dev)   # queues + the transposed copy of the NCHW upstream gradients

        def bwd_n(pool, g):
            L.check(L.lib.mrcnn_pyramid_roi_align_backward(g.data_ptr(), L.NCHW, wl.Hs, wl.Ws, wl.batch, CHANNELS, wl.boxes.data_ptr(),
                                                           wl.ind.data_ptr(), wl.N, pool, wl.area, pg, L.NCHW, 1, None, L.BWD_AUTO,
                                                           ws_n.data_ptr(), ws_n.numel(), wl._s()))

        def step_n():
            fwd_n(7, o7)
            fwd_n(14, o14)
            wl.mask_targets()
            bwd_n(14, g14)
            bwd_n(7, g7)
        parts = {"fwd7": lambda: fwd_n(7, o7), "fwd14": lambda: fwd_n(14, o14), "bwd7": lambda: bwd_n(7, g7), "bwd14": lambda: bwd_n(14, g14)}
        U7, _ = roofline.unique_taps(wl.boxes_np, wl.ind_np, 7, (IMAGE, IMAGE), LEVEL_HW, wl.batch)
        U14, _ = roofline.unique_taps(wl.boxes_np, wl.ind_np, 14, (IMAGE, IMAGE), LEVEL_HW, wl.batch)
        pyr = wl.batch * PYR_ELEMS_PER_IMAGE
        alg = {"fwd7": roofline.roialign_fwd_bytes(wl.N, CHANNELS, 7, U7), "fwd14": roofline.roialign_fwd_bytes(wl.N, CHANNELS, 14, U14),
               "bwd7": roofline.roialign_bwd_bytes(wl.N, CHANNELS, 7, pyr), "bwd14": roofline.roialign_bwd_bytes(wl.N, CHANNELS, 14, pyr)}
        for k_, fn in parts.items():
            tp = wl.time_op(fn, iters=10)
            nchw["train_" + k_] = {"ms": tp * 1e3, "algorithmic_MB": alg[k_] / 1e6, "frac_of_hbm": alg[k_] / tp / 1e9 / hbm}
        tstep = wl.time_op(step_n, iters=10)
        nchw["train_step_nchw_pyramid"] = {"ms_per_step": tstep * 1e3, "rois_per_s": wl.N / tstep,
                                           "step_frac_of_hbm": (sum(alg.values()) + wl.mt.numel() * 4) / tstep / 1e9 / hbm,
                                           "config": "configs[3] with the pyramid, crops and gradients all NCHW-contiguous (same algorithmic bytes)"}
        out["nchw_pyramid"] = nchw
        del fm_nchw, o7, o14, g7, g14, gfn, ws_n
    except Exception as e:
        out["nchw_pyramid"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    # the standalone drop-in nms (c++ext/maskrcnn/__init__.py:21-22) at the proposal layer's size: 6000 boxes, IoU 0.7
    try:
        rng_ = np.random.default_rng(11)      # clustered boxes (every box has a jittered twin) in descending score order, like rpn_refine's call
        b_ = synth.random_rois(6000, 11, image=float(IMAGE), min_size=16, max_size=500) * float(IMAGE)
        b_[3000:] = b_[:3000] + rng_.uniform(-8, 8, (3000, 4)).astype(np.float32)
        d5 = np.concatenate([b_, np.sort(synth.unique_scores(6000, 11))[::-1][:, None]], 1).astype(np.float32)
        d5 = torch.from_numpy(d5).to(dev)
        keep = torch.empty(6000, dtype=torch.int64, device=dev)
        cnt = torch.empty(1, dtype=torch.int32, device=dev)
        wsn = torch.empty(L.lib.mrcnn_nms_workspace_bytes(6000), dtype=torch.uint8, device=dev)
        tk = wl.time_op(lambda: L.check(L.lib.mrcnn_nms(d5.data_ptr(), 6000, 0.7, keep.data_ptr(), cnt.data_ptr(), wsn.data_ptr(), wsn.numel(), wl._s())),
                        iters=50)
        tn = wl.time_op(lambda: m.nms(d5, 0.7), iters=50)
        d500 = d5[:500].contiguous()      # the size this fork's rpn_refine calls it with (model.py:1345 hard-codes 500)
        t500 = wl.time_op(lambda: m.nms(d500, 0.7), iters=50)
        out["nms_standalone"] = {"config": "maskrcnn.nms on 6000 clustered boxes in score order, IoU 0.7", "kernels_us": tk * 1e6,
                                 "dropin_us_incl_count_readback": tn * 1e6, "dropin_us_500_boxes": t500 * 1e6, "kept": int(cnt.item()),
                                 "algorithmic_GBps": (6000 * 20 + int(cnt.item()) * 8) / tk / 1e9}
    except Exception as e:
        out["nms_standalone"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    # configs[4]: detection layer + mask RoIAlign on the detections, 64 images
    B, N, NC = 64, 1000, 81
    rois = torch.from_numpy(np.stack([synth.random_rois(N, 300 + i) for i in range(B)])).to(dev)
    g_ = torch.Generator(device=dev)
    g_.manual_seed(7)
    probs = torch.softmax(3.0 * torch.randn((B, N, NC), device=dev, generator=g_), -1)
    deltas = 0.1 * torch.randn((B, N, NC, 4), device=dev, generator=g_)
    win = torch.tensor([[0., 0., IMAGE, IMAGE]], device=dev).repeat(B, 1)
    fm64 = wl.fm  # 16 images of pyramid; detections of image i read pyramid i % 16
    def det_path():
        dets, counts = m.detection_layer(rois, probs, deltas, win, 0.0, 0.3, 100)
        b = (dets[:, :, :4] / float(IMAGE)).reshape(-1, 4)
        ind = (torch.arange(B, device=dev, dtype=torch.int32) % wl.batch).repeat_interleave(100)
        return m.pyramid_roi_align(fm64, b, ind, 14, (IMAGE, IMAGE, 3)), counts
    t = wl.time_op(det_path, iters=20)
    t_det = wl.time_op(lambda: m.detection_layer(rois, probs, deltas, win, 0.0, 0.3, 100), iters=20)
    m.set_detection_nms("mask")
    try:
        t_det_mask = wl.time_op(lambda: m.detection_layer(rois, probs, deltas, win, 0.0, 0.3, 100), iters=20)
    finally:
        m.set_detection_nms("auto")
    out["detection_path"] = {"detection_layer_only_ms": t_det * 1e3, "detection_layer_only_ms_mask_nms": t_det_mask * 1e3,"config": "configs[4]: detection layer (81 classes, 1000 RoIs, NMS 0.3, top-100) + 14x14 mask RoIAlign, 64 images",
                             "images_per_s": B / t, "ms_per_64_images": t * 1e3, "detection_layer_only_images_per_s": B / t_det,
                             "frac_of_hbm_detection_layer": B * roofline.detection_bytes(N, NC, 100) / t_det / 1e9 / hbm}
    return out


def sharded_detection(torch, dist, wl, world, rank, hbm):
    """BASELINE configs[4]: 64 images sharded over the ranks (strong scaling): detection layer + 14x14 mask RoIAlign on
    the detections, then ONE all-gather of the padded detections (NCCL over NVLink).  Max over ranks."""
    import maskrcnn_b200 as m
    from maskrcnn_b200 import synth, dist as mdist
    dev = "cuda"
    TOTAL, N, NC, D = 64, 1000, 81, 100
    b, e = mdist.shard_range(TOTAL, rank, world)
    Bl = e - b

    def image_heads(i):          # every per-image input is a function of the image id alone: G-invariant results
        pr, de = synth.head_outputs(N, NC, 700 + i)
        return synth.random_rois(N, 300 + i), pr, de
    per = [image_heads(i) for i in range(b, e)]
    rois = torch.from_numpy(np.stack([p_[0] for p_ in per])).to(dev)
    probs = torch.from_numpy(np.stack([p_[1] for p_ in per])).to(dev)
    deltas = torch.from_numpy(np.stack([p_[2] for p_ in per])).to(dev)
    win = torch.tensor([[0., 0., IMAGE, IMAGE]], device=dev).repeat(Bl, 1)
    ind = (torch.arange(b, e, device=dev, dtype=torch.int32) % wl.batch).repeat_interleave(D)   # image i reads pyramid i % 16

    ex = mdist.DetectionExchange(TOTAL, D)
    side_c = torch.cuda.Stream()
    L = wl.L
    masks_in = torch.empty((Bl * D, CHANNELS, 14, 14), device=dev, memory_format=torch.channels_last)
    fmp = L.vp4([f.data_ptr() for f in wl.fm])

    def run():
        """Three launches: detection layer fused with the sending half of the exchange (it also writes the mask head's RoIs),
        the 14x14 RoIAlign of the detections, the collecting half."""
        ex.run(rois, probs, deltas, win, 0.0, 0.3, ind_offset=b, ind_mod=wl.batch)
        cur = torch.cuda.current_stream()
        side_c.wait_stream(cur)
        with torch.cuda.stream(side_c):          # the collecting half runs beside the RoIAlign: it only needs the peers' flags
            all_dets, all_counts = ex.collect()
        L.check(L.lib.mrcnn_pyramid_roi_align_forward(fmp, wl.Hs, wl.Ws, wl.batch, CHANNELS, L.NHWC, ex.mask_boxes.data_ptr(),
                                                      ex.mask_box_ind.data_ptr(), Bl * D, 14, wl.area, masks_in.data_ptr(), L.NHWC, None, wl._s()))
        cur.wait_stream(side_c)
        return masks_in, all_dets, all_counts

    def run_nccl():
        """Round 1's path for comparison: detection layer, torch glue, RoIAlign through ops, torch.cat + all_gather_into_tensor."""
        dets, counts = m.detection_layer(rois, probs, deltas, win, 0.0, 0.3, D)
        mi = m.pyramid_roi_align(wl.fm, (dets[:, :, :4] / float(IMAGE)).reshape(-1, 4), ind, 14, (IMAGE, IMAGE, 3))
        all_dets, all_counts = mdist.gather_detections(dets, counts, n_images=TOTAL)
        return mi, all_dets, all_counts

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # the three launches as one CUDA graph per rank (every rank replays the same number of times)
    step, mode = run, "eager launches"
    try:
        cap = torch.cuda.Stream()
        cap.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=cap):
            run()
        torch.cuda.current_stream().wait_stream(cap)
        step, mode = graph.replay, "one CUDA graph replay (3 kernels)"
    except Exception as e_:
        sys.stderr.write("detection path: graph capture failed (%s: %s)\n" % (type(e_).__name__, e_))
    ok_flag = torch.tensor([1 if step is not run else 0], device=dev)
    if world > 1:   # all ranks must agree (a rank replaying a graph while another launches eagerly would still be correct, but keep it uniform)
        dist.all_reduce(ok_flag, op=dist.ReduceOp.MIN)
    if int(ok_flag.item()) == 0:
        step, mode = run, "eager launches"
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # every iteration is timed on its own (events on the launching stream) and the MEDIAN is reported: one host hiccup (the
    # pool's boxes share their cores) in a 20-iteration window otherwise decides the number; mean and max are kept beside it
    iters = 20
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    evs[0].record()
    for i in range(iters):
        step()
        evs[i + 1].record()
    torch.cuda.synchronize()
    out = (masks_in, ex.dets_all, ex.counts_all)
    per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(iters))
    ms, ms_mean, ms_max = per[iters // 2], sum(per) / iters, per[-1]
    if world > 1:
        dist.barrier()

    # the rank-local part and the exchange timed apart
    def local_part():
        m.detection_layer(rois, probs, deltas, win, 0.0, 0.3, D)
        L.check(L.lib.mrcnn_pyramid_roi_align_forward(fmp, wl.Hs, wl.Ws, wl.batch, CHANNELS, L.NHWC, ex.mask_boxes.data_ptr(),
                                                      ex.mask_box_ind.data_ptr(), Bl * D, 14, wl.area, masks_in.data_ptr(), L.NHWC, None, wl._s()))
    ms_local = wl.time_op(local_part, iters=iters) * 1e3

    def exchange_only():         # what the exchange adds to the rank-local part: the fused epilogue's peer stores + the collect kernel
        ex.run(rois, probs, deltas, win, 0.0, 0.3, ind_offset=b, ind_mod=wl.batch)
        ex.collect()
    t_ex = wl.time_op(exchange_only, iters=iters) * 1e3
    t_det = wl.time_op(lambda: m.detection_layer(rois, probs, deltas, win, 0.0, 0.3, D), iters=iters) * 1e3
    ms_gather = max(t_ex - t_det, 0.0)
    if world > 1:
        dist.barrier()
    ms_nccl = wl.time_op(run_nccl, iters=iters) * 1e3
    dets_l, counts_l = m.detection_layer(rois, probs, deltas, win, 0.0, 0.3, D)
    ms_nccl_gather = wl.time_op(lambda: mdist.gather_detections(dets_l, counts_l, n_images=TOTAL), iters=iters) * 1e3
    nccl_dets, nccl_counts = mdist.gather_detections(dets_l, counts_l, n_images=TOTAL)
    same_as_nccl = bool(torch.equal(nccl_dets, ex.dets_all)) and bool(torch.equal(nccl_counts, ex.counts_all))
    # G-invariance: the gathered [64, D, 6] must equal, bit for bit and in image order, what ONE GPU computes
    # image by image.  Rank 0 recomputes all 64 images one call at a time (batch 1, no sharding, no collective) and compares.
    g_invariant = None
    if rank == 0:
        all_dets, all_counts = out[1], out[2]
        ok = tuple(all_dets.shape) == (TOTAL, D, 6)
        one = torch.tensor([[0., 0., IMAGE, IMAGE]], device=dev)
        for i in range(TOTAL):
            r_, p_, d_ = image_heads(i)
            di, ci = m.detection_layer(torch.from_numpy(r_)[None].to(dev), torch.from_numpy(p_)[None].to(dev),
                                       torch.from_numpy(d_)[None].to(dev), one, 0.0, 0.3, D)
            ok = ok and bool(torch.equal(di[0], all_dets[i])) and int(ci[0]) == int(all_counts[i])
        g_invariant = bool(ok)
    if world > 1:
        t = torch.tensor([ms, ms_mean, ms_max, ms_local, ms_gather, ms_nccl, ms_nccl_gather, 0.0 if same_as_nccl else 1.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_mean, ms_max, ms_local, ms_gather, ms_nccl, ms_nccl_gather, bad = (float(v) for v in t.tolist())
        same_as_nccl = bad == 0.0
    return {"config": "configs[4]: 64 images sharded over %d GPU(s): detection layer + 14x14 mask RoIAlign + one all-gather" % world,
            "images_per_s": TOTAL / (ms * 1e-3), "ms_per_64_images": ms, "ms_per_64_images_mean": ms_mean, "ms_per_64_images_max": ms_max,
            "statistic": "median of %d iterations, max over ranks" % iters, "scaling": "strong", "launch": mode,
            "ms_rank_local_part": ms_local, "ms_all_gather": ms_gather,
            "exchange": "fused: the detection kernel stores every image's packed row into all ranks' receive buffers (NVLink peer stores, "
                        "%d B per image per peer) and raises a flag; mrcnn_detection_collect waits per image and copies out.  ms_all_gather = "
                        "(fused detection + collect) - (plain detection layer), eager, mean of %d" % ((D * 6 + 1) * 4, iters),
            "nvlink_bytes_per_rank": int((world - 1) * Bl * (D * 6 + 1) * 4),
            "nccl_path": {"ms_per_64_images": ms_nccl, "ms_all_gather": ms_nccl_gather, "identical_results": same_as_nccl,
                          "note": "round 1's path: detection layer + torch glue + ops RoIAlign + torch.cat / all_gather_into_tensor, eager, mean"},
            "gathered_images": int(out[1].shape[0]), "mean_detections": float(out[2].float().mean().item()),
            "g_invariant": g_invariant,
            "g_invariant_check": "rank 0: gathered detections of all 64 images == the same images computed one by one on one GPU "
                                 "(torch.equal on [D,6] rows and counts); inputs are functions of the image id"}


def rpn_nms(torch, dist, wl, world, rank, hbm):
    """BASELINE configs[1] on every rank (weak scaling, no collective): proposal layer, 261,888 anchors, top-6000 ->
    NMS 0.7 -> 1000, batch 8 per GPU.  CUDA events, max over ranks."""
    import maskrcnn_b200 as m
    from maskrcnn_b200 import synth, roofline
    dev = "cuda"
    anchors = synth.pyramid_anchors((IMAGE, IMAGE))
    rcs, rbs = zip(*[synth.rpn_outputs(anchors, 1235 + 8 * rank + i) for i in range(8)])
    rc, rb, an = torch.from_numpy(np.stack(rcs)).to(dev), torch.from_numpy(np.stack(rbs)).to(dev), torch.from_numpy(anchors).to(dev)
    f = lambda: m.proposal_layer(rc, rb, an, 6000, 1000, 0.7)  # noqa: E731
    t = wl.time_op(f, iters=20)
    _, counts = f()
    # the same layer with the N x N mask + sweep NMS, and both on inputs whose top boxes converge on a few objects (heavy
    # suppression from the first box on, as a trained RPN produces): reported next to the headline, rank-local
    variants = {}
    rcs2, rbs2 = zip(*[synth.rpn_outputs(anchors, 1235 + 8 * rank + i, converge=0.9) for i in range(2)])
    rc2 = torch.from_numpy(np.stack([rcs2[i % 2] for i in range(8)])).to(dev)
    rb2 = torch.from_numpy(np.stack([rbs2[i % 2] for i in range(8)])).to(dev)
    try:
        for algo in ("hybrid", "lazy", "mask"):
            m.set_proposal_nms(algo)
            variants["ms_per_batch_%s_nms" % algo] = wl.time_op(f, iters=20) * 1e3
            g = lambda: m.proposal_layer(rc2, rb2, an, 6000, 1000, 0.7)  # noqa: E731
            variants["converged_inputs_ms_per_batch_%s_nms" % algo] = wl.time_op(g, iters=20) * 1e3
            variants["converged_inputs_kept_mean"] = float(g()[1].float().mean().item())
    finally:
        m.set_proposal_nms("auto")
    # one batch of 8 occupies 64 of the SMs (an 8-CTA cluster per image): two independent batches of 8 on two streams, as
    # a server that pipelines batches would issue them.  Reported beside the headline, never instead of it; rank-local.
    try:
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        rc_b, rb_b = rc.flip(0).contiguous(), rb.flip(0).contiguous()
        lanes = ((s1, rc, rb), (s2, rc_b, rb_b))

        def in_flight(iters):
            cur = torch.cuda.current_stream()
            for s, _, _ in lanes:
                s.wait_stream(cur)
            for _ in range(iters):
                for s, c, b in lanes:
                    with torch.cuda.stream(s):
                        m.proposal_layer(c, b, an, 6000, 1000, 0.7)
            for s, _, _ in lanes:
                cur.wait_stream(s)
        in_flight(3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        in_flight(20)
        e1.record()
        torch.cuda.synchronize()
        t2 = e0.elapsed_time(e1) / 20 * 1e-3
        variants["two_batches_in_flight"] = {"ms_per_two_batches": t2 * 1e3, "images_per_s_per_gpu": 16 / t2,
                                             "note": "two streams, 8 images each, no join between iterations"}
    except Exception as e:  # an extra: it must never cost the line
        variants["two_batches_in_flight"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if world > 1:
        tt = torch.tensor([t], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt.item())
    by = 8 * roofline.proposal_bytes(len(anchors), 6000, 1000)
    return {"config": "configs[1]: 261,888 anchors, top-6000 -> NMS 0.7 -> 1000, batch 8 per GPU", "images_per_s": world * 8 / t,
            "ms_per_batch": t * 1e3, "kept_mean": float(counts.float().mean().item()), "scaling": "weak",
            "algorithmic_GBps_per_gpu": by / t / 1e9, "frac_of_hbm": by / t / 1e9 / hbm,
            "nms_variants": variants,
            "note": "latency-bound (multi-pass select, then NMS); the default NMS is the hybrid one - the first 1.25 post_nms boxes of all "
                    "images resolved at once by a grid-wide fixed-point iteration over (20 chunks x 8 images) CTAs, the lazy cluster "
                    "kernel only for images still short of survivors; 4 launches per batch (select, prefix mask, fixed point, lazy "
                    "tail).  The lazy kernel alone and the N x N mask + sweep path are timed beside it"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="timed steps only (use under ncu)")
    ap.add_argument("--eager", action="store_true",
                    help="time eagerly launched steps instead of replaying the step's CUDA graph (the default)")
    ap.add_argument("--graph-step", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        # stdout carries ONE JSON line: NCCL prints its "NCCL version ..." banner with a bare printf when the first
        # communicator comes up, so file descriptor 1 points at stderr until that has happened
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    hbm, peak_src = hbm_peak()
    wl = Workload(torch, torch.device("cuda", local), first_image=rank * BATCH)

    step_fn = None if args.eager else capture_step(torch, wl)
    with ClockSampler(local) as cs:
        ms = timed_steps(torch, dist, wl, args.steps, args.warmup, world, step_fn)
    launches = wl.launches
    clocks = cs.summary()
    per_step = ms / args.steps
    value = world * wl.N / (per_step * 1e-3)

    line = {"metric": "roialign_train_rois_per_s", "value": value, "unit": "RoIs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(), "clocks": clocks, "gpu_launches": launches}
    line["config"]["launch"] = ("one CUDA graph replay per step (10 kernels: the two heads' forward as one launch, mask-target crop, "
                                "2 x 3 queue kernels on a side stream, 2 gathers)" if step_fn is not None else "eager launches")

    if not args.no_extras:
        from maskrcnn_b200 import roofline
        # per-kernel breakdown, each op alone (burst), CUDA events on the launching stream
        U7, _ = roofline.unique_taps(wl.boxes_np, wl.ind_np, 7, (IMAGE, IMAGE), LEVEL_HW, wl.batch)
        U14, _ = roofline.unique_taps(wl.boxes_np, wl.ind_np, 14, (IMAGE, IMAGE), LEVEL_HW, wl.batch)
        pyr = wl.batch * PYR_ELEMS_PER_IMAGE
        def plans():
            wl.plan(14, wl.ws, torch.cuda.current_stream())
            wl.plan(7, wl.ws7, torch.cuda.current_stream())
        plan_name = "bwd_items x2+bwd_alloc for both heads (side stream, overlaps the forward)"
        ops = {
            "roialign_fwd_nhwc_pair_kernel<7+14,nhwc>": (wl.fwd_pair, roofline.roialign_fwd_bytes(wl.N, CHANNELS, 7, U7) +
                                                         roofline.roialign_fwd_bytes(wl.N, CHANNELS, 14, U14)),
            "roialign_bwd_gather_kernel<7,nhwc>": (lambda: wl.bwd_planned(7, wl.g7, wl.gfm7, wl.ws7), roofline.roialign_bwd_bytes(wl.N, CHANNELS, 7, pyr)),
            "roialign_bwd_gather_kernel<14,nhwc>": (lambda: wl.bwd_planned(14, wl.g14, wl.gfm14, wl.ws), roofline.roialign_bwd_bytes(wl.N, CHANNELS, 14, pyr)),
            "crop_plane_fwd_kernel<28x28 mask targets>": (wl.mask_targets, mask_target_bytes(wl)),
            plan_name: (plans, 2 * wl.N * 20),
        }
        captures = dict(zip(ops, ("fwd_pair_nhwc", "bwd7_nhwc:gather", "bwd14_nhwc:gather", None, None)))
        traffic = ncu_traffic()
        kern = {}
        torch.cuda.synchronize()
        plans()                                              # the gathers below read these queues
        for name, (fn, by) in ops.items():
            t = wl.time_op(fn)
            kern[name] = {"ms": t * 1e3, "algorithmic_MB": by / 1e6, "GBps": by / t / 1e9, "frac": by / t / 1e9 / hbm,
                          "ncu_dram_MB": traffic.get(captures[name])}
        kern[plan_name]["note"] = "not on the main stream: runs beside the forward kernels; its bytes (boxes) are not credited"
        mt_name = "crop_plane_fwd_kernel<28x28 mask targets>"
        sec = mask_target_sector_bytes(wl)
        kern[mt_name].update({"ncu_dram_MB": traffic.get("mask_targets"), "unique_sector_MB": sec / 1e6,
                              "frac_in_sectors": sec / (kern[mt_name]["ms"] * 1e-3) / 1e9 / hbm,
                              "note": "algorithmic_MB counts 4 bytes per unique mask pixel; DRAM moves 32-byte sectors and the taps of a "
                                      "large box are further apart than that, so the kernel's floor is unique_sector_MB (ncu_dram_MB is "
                                      "what it moved: profiles/r02_mask_targets_ncu_summary.txt)"})
        # the same launch with the footprint both heads share counted ONCE (what the kernel actually has to read): the conservative figure
        U_both, _ = roofline.unique_taps(wl.boxes_np, wl.ind_np, (7, 14), (IMAGE, IMAGE), LEVEL_HW, wl.batch)
        by_once = wl.N * CHANNELS * (49 + 196) * 4 + 4 * CHANNELS * U_both + wl.N * 20
        pk = kern["roialign_fwd_nhwc_pair_kernel<7+14,nhwc>"]
        pk["footprint_once_MB"] = by_once / 1e6
        pk["frac_footprint_once"] = by_once / (pk["ms"] * 1e-3) / 1e9 / hbm
        kern["roialign_fwd_nhwc_pair_kernel<7+14,nhwc>"]["note"] = (
            "both heads in one launch; algorithmic bytes = the two heads' bytes as SURVEY 8(d) defines them (each head's unique taps "
            "counted), while the launch reads the shared footprint once - its DRAM traffic is below that sum; footprint_once_MB / "
            "frac_footprint_once count the shared footprint once")
        # the single-head forward kernels, timed alone for comparison (not launches of the step)
        single = {}
        for pool, o_, U_ in ((7, wl.out7, U7), (14, wl.out14, U14)):
            t_ = wl.time_op(lambda: wl.fwd(pool, o_))
            by_ = roofline.roialign_fwd_bytes(wl.N, CHANNELS, pool, U_)
            single["roialign_fwd_nhwc_col_kernel<%d,nhwc>" % pool] = {"ms": t_ * 1e3, "algorithmic_MB": by_ / 1e6, "GBps": by_ / t_ / 1e9,
                                                                     "frac": by_ / t_ / 1e9 / hbm,
                                                                     "ncu_dram_MB": traffic.get("fwd%d_nhwc" % pool)}
        t_two = wl.time_op(wl.step_two_forwards, iters=20)
        total = sum(k["ms"] for n, k in kern.items() if n != plan_name)
        for k in kern.values():
            k["share_of_step"] = k["ms"] / total
        # the same step with each backward building its own queues on the main stream (mrcnn_pyramid_roi_align_backward)
        t_eager = wl.time_op(wl.step, iters=20)
        line["eager_step"] = {"ms_per_step": t_eager * 1e3, "rois_per_s": wl.N / t_eager, "note": "the same step launched kernel by kernel from Python"}
        t_unplanned = wl.time_op(wl.step_unplanned, iters=10)
        t_bwd14 = wl.time_op(lambda: wl.bwd(14, wl.g14, wl.gfm14))
        t_bwd7 = wl.time_op(lambda: wl.bwd(7, wl.g7, wl.gfm7))
        line["unplanned_step"] = {"ms_per_step": t_unplanned * 1e3, "rois_per_s": wl.N / t_unplanned,
                                  "bwd14_with_own_queues_ms": t_bwd14 * 1e3, "bwd7_with_own_queues_ms": t_bwd7 * 1e3,
                                  "note": "queues built inside each backward call (4 launches per backward) instead of beside the forward"}
        top = max((n for n in kern if n != plan_name), key=lambda n: kern[n]["ms"])
        line["roofline"] = {"bound": "hbm", "kernel": top, "achieved": kern[top]["GBps"], "peak": hbm, "unit": "GB/s",
                            "frac": kern[top]["frac"],
                            "traffic": (kern[top]["ncu_dram_MB"] * 1e6 if kern[top]["ncu_dram_MB"] else None),
                            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum over the op's kernels, one launch each, "
                                              "ncu --set full captures summarised in profiles/*_traffic.json",
                            "peak_source": peak_src,
                            "step_frac": sum(k["algorithmic_MB"] for k in kern.values()) / 1e3 / (per_step * 1e-3) / hbm,
                            "step_frac_footprint_once": sum(k.get("footprint_once_MB", k["algorithmic_MB"]) for k in kern.values()) / 1e3 /
                                                        (per_step * 1e-3) / hbm,
                            "kernels": kern, "single_head_forward_kernels": single,
                            "step_with_one_forward_per_head": {"ms_per_step": t_two * 1e3, "rois_per_s": wl.N / t_two}}
        line["e2e"] = e2e_run(torch, dist, wl, args.steps, args.warmup, world)
        if wl.cl_crops:
            try:
                line["e2e_fused_backward"] = e2e_run(torch, dist, wl, args.steps, args.warmup, world, fused_backward=True)
                line["e2e_fused_backward"]["note"] = ("not the headline: the two heads' gradient pyramids leave the device summed (what a training step "
                                                      "accumulates), so 1.43 GB less crosses the host link per step")
            except Exception as e_:
                line["e2e_fused_backward"] = {"error": "%s: %s" % (type(e_).__name__, str(e_)[:300])}
        line["e2e"]["host_cpus_bound"] = (len(numa) if numa else None)
        line["detection_path_sharded"] = sharded_detection(torch, dist, wl, world, rank, hbm)
        line["rpn_nms"] = rpn_nms(torch, dist, wl, world, rank, hbm)
        if rank == 0 and world == 1:
            cores = os.cpu_count() or 1
            arm = CpuArm(1)                                  # warms up on one full image
            v, dt = arm.measure(8)
            line["cpu_baseline"] = {"value": v, "unit": "RoIs/s", "cores": 1, "kind": arm.kind, "host_cores": cores, "cpu_model": cpu_model(),
                                    "sample": "8 passes over image 0 of the GPU arm's batch x %d RoIs (same five ops), single thread as the reference ships, %.1f s" % (ROIS_PER_IMAGE, dt)}
            line["also"] = secondary(torch, wl, hbm)
            try:
                line["also"]["cpu_baselines_other_configs"] = cpu_config_baselines()
            except Exception as e:   # extras must never cost the line
                line["also"]["cpu_baselines_other_configs"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            try:
                line["also"]["predict_flow"] = predict_flow(torch)
            except Exception as e:
                line["also"]["predict_flow"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
