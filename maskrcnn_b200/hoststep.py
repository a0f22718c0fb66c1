"""The RoIAlign training step for callers whose tensors live in HOST memory (the `e2e` leg of bench.py; a data-loader or a CPU
framework driving this library): every input crosses the host link on the way in, every result on the way out, and the three
phases of consecutive images overlap.

    image i:   H2D (pyramid slice, boxes, upstream gradients)  ->  kernels  ->  D2H (crops, mask targets, gradient pyramids)
    streams:   copy-in                                             compute       copy-out          (events chain image i's phases)

The step is the one bench.py times on the device (BASELINE configs[3]): PyramidROIAlign forward at 7x7 and 14x14 of the same RoIs
(model.py:778 / :889), the 28x28 mask-target crops (model.py:501-502) and the two backwards - or, with fused_backward, ONE
backward that leaves the SUM of the two heads' gradient pyramids (what autograd accumulates), so one pyramid instead of two goes
back over the link.  All tensors channels-last; host tensors pinned and in the device's physical order, so that every copy is one
plain memcpy.  PyTorch provides pinned memory, streams and events; the work is the C ABI's."""
import torch

from . import _lib

__all__ = ["HostTrainStep"]


class HostTrainStep(object):
    """Device staging buffers + streams for `batch` images of `rois_per_image` RoIs on a pyramid of `channels` channels with
    levels `level_hw` = [(H2, W2), ..., (H5, W5)] of an `image_hw` image."""

    def __init__(self, batch, rois_per_image, channels, level_hw, image_hw, fused_backward=False, device=None):
        self.B, self.R, self.C = int(batch), int(rois_per_image), int(channels)
        if self.C % 4:
            raise ValueError("channels must be a multiple of 4 (channels-last vector kernels)")
        self.level_hw = [(int(h), int(w)) for h, w in level_hw]
        self.area = float(image_hw[0] * image_hw[1])
        self.fused = bool(fused_backward)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        dev, cl = self.device, torch.channels_last
        N = self.B * self.R
        self.N = N
        e = lambda *s: torch.empty(s, device=dev).contiguous(memory_format=cl)  # noqa: E731
        self.fm = [e(self.B, self.C, h, w) for h, w in self.level_hw]
        self.gfm_a = [e(self.B, self.C, h, w) for h, w in self.level_hw]             # 14x14 head (or the sum when fused)
        self.gfm_b = None if self.fused else [e(self.B, self.C, h, w) for h, w in self.level_hw]
        self.boxes = torch.empty((N, 4), device=dev)
        self.out7, self.out14 = e(N, self.C, 7, 7), e(N, self.C, 14, 14)
        self.g7, self.g14 = e(N, self.C, 7, 7), e(N, self.C, 14, 14)
        self.Hs, self.Ws = _lib.i4([h for h, _ in self.level_hw]), _lib.i4([w for _, w in self.level_hw])
        L = _lib.lib
        # one image at a time: per-image workspaces, reused (the compute stream orders their uses)
        self.ws14 = torch.empty(L.mrcnn_pyramid_roi_align_backward_workspace_bytes(self.Hs, self.Ws, 1, self.R, 14), dtype=torch.uint8, device=dev)
        self.ws7 = torch.empty(L.mrcnn_pyramid_roi_align_backward_workspace_bytes(self.Hs, self.Ws, 1, self.R, 7), dtype=torch.uint8, device=dev)
        self.ws_pair = torch.empty(L.mrcnn_pyramid_roi_align_backward_pair_workspace_bytes(self.Hs, self.Ws, 1, self.R, 7, 14), dtype=torch.uint8,
                                   device=dev) if self.fused else None
        self.s_in, self.s_run, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.launches = 0
        self._phys = lambda t: t.permute(0, 2, 3, 1)      # logical [N,C,h,w] channels-last -> the physical [N,h,w,C] view

    # host-side helpers ------------------------------------------------------------------------------------------------
    def pinned_like(self, t):
        """A pinned host tensor with the PHYSICAL shape of the channels-last device tensor t ([N,h,w,C])."""
        return torch.empty(self._phys(t).shape, dtype=t.dtype).pin_memory()

    def h2d_bytes(self, mask=None):
        b = sum(f.numel() for f in self.fm) * 4 + self.boxes.numel() * 4 + (self.g7.numel() + self.g14.numel()) * 4
        if mask is not None:
            b += (mask["h_boxes"].numel() + mask["h_index"].numel()) * 4
        return b

    def d2h_bytes(self, mask=None):
        b = (self.out7.numel() + self.out14.numel()) * 4 + (1 if self.fused else 2) * sum(f.numel() for f in self.fm) * 4
        if mask is not None:
            b += mask["h_targets"].numel() * 4
        return b

    # the step -----------------------------------------------------------------------------------------------------------
    def run(self, h_fm, h_boxes, h_g7, h_g14, h_out7, h_out14, h_gfm_a, h_gfm_b=None, mask=None):
        """One step.  Pinned host tensors, physical (channels-last) order: h_fm[l] [B,H,W,C]; h_boxes [B*R,4] normalised, grouped by
        image; h_g7 / h_g14 upstream gradients [N,7,7,C] / [N,14,14,C]; results into h_out7 / h_out14 (same shapes), h_gfm_a[l] (and
        h_gfm_b[l] unless fused) [B,H,W,C].  mask (optional): dict(d_images=[G,1,H,W] device tensor of instance masks, h_boxes
        [P*B,4], h_index int32 [P*B] (row of d_images), d_boxes / d_index / d_targets device staging, h_targets [P*B,1,28,28]) for the
        28x28 mask-target crops, P per image.  Returns after everything has landed in host memory."""
        L, B, R, C = _lib, self.B, self.R, self.C
        lib = _lib.lib
        fm_p = [self._phys(f) for f in self.fm]
        ga_p = [self._phys(f) for f in self.gfm_a]
        gb_p = None if self.fused else [self._phys(f) for f in self.gfm_b]
        d_g7, d_g14, d_o7, d_o14 = self._phys(self.g7), self._phys(self.g14), self._phys(self.out7), self._phys(self.out14)
        P = 0 if mask is None else mask["h_boxes"].shape[0] // B
        ev_in = [torch.cuda.Event() for _ in range(B)]
        ev_run = [torch.cuda.Event() for _ in range(B)]
        for i in range(B):
            rs = slice(i * R, (i + 1) * R)
            ms = slice(i * P, (i + 1) * P)
            with torch.cuda.stream(self.s_in):
                for l in range(4):
                    fm_p[l][i].copy_(h_fm[l][i], non_blocking=True)
                self.boxes[rs].copy_(h_boxes[rs], non_blocking=True)
                d_g7[rs].copy_(h_g7[rs], non_blocking=True)
                d_g14[rs].copy_(h_g14[rs], non_blocking=True)
                if mask is not None:
                    mask["d_boxes"][ms].copy_(mask["h_boxes"][ms], non_blocking=True)
                    mask["d_index"][ms].copy_(mask["h_index"][ms], non_blocking=True)
                ev_in[i].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(ev_in[i])
                st = self.s_run.cuda_stream
                fmp = L.vp4([f[i].data_ptr() for f in self.fm])
                bp = self.boxes[rs].data_ptr()
                L.check(lib.mrcnn_pyramid_roi_align_forward_pair(fmp, self.Hs, self.Ws, 1, C, bp, None, R, self.area, self.out7[rs].data_ptr(),
                                                                 self.out14[rs].data_ptr(), st))
                self.launches += 1
                if mask is not None:
                    gt = mask["d_images"]
                    L.check(lib.mrcnn_crop_forward(gt.data_ptr(), gt.shape[0], 1, gt.shape[2], gt.shape[3], L.NCHW, mask["d_boxes"][ms].data_ptr(),
                                                   mask["d_index"][ms].data_ptr(), P, 0.0, 28, 28, mask["d_targets"][ms].data_ptr(), L.NCHW, st))
                    self.launches += 1
                ga = L.vp4([g[i].data_ptr() for g in self.gfm_a])
                if self.fused:
                    L.check(lib.mrcnn_pyramid_roi_align_backward_pair(self.g7[rs].data_ptr(), 7, self.g14[rs].data_ptr(), 14, self.Hs, self.Ws, 1, C,
                                                                      bp, None, R, self.area, ga, 1, self.ws_pair.data_ptr(), self.ws_pair.numel(), st))
                    self.launches += 6          # count x2, alloc, fill x2, gather
                else:
                    gb = L.vp4([g[i].data_ptr() for g in self.gfm_b])
                    L.check(lib.mrcnn_pyramid_roi_align_backward(self.g14[rs].data_ptr(), L.NHWC, self.Hs, self.Ws, 1, C, bp, None, R, 14, self.area,
                                                                 ga, L.NHWC, 1, None, L.BWD_AUTO, self.ws14.data_ptr(), self.ws14.numel(), st))
                    L.check(lib.mrcnn_pyramid_roi_align_backward(self.g7[rs].data_ptr(), L.NHWC, self.Hs, self.Ws, 1, C, bp, None, R, 7, self.area,
                                                                 gb, L.NHWC, 1, None, L.BWD_AUTO, self.ws7.data_ptr(), self.ws7.numel(), st))
                    self.launches += 8          # 2 x (count, alloc, fill, gather)
                ev_run[i].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_run[i])
                h_out7[rs].copy_(d_o7[rs], non_blocking=True)
                h_out14[rs].copy_(d_o14[rs], non_blocking=True)
                if mask is not None:
                    mask["h_targets"][ms].copy_(mask["d_targets"][ms], non_blocking=True)
                for l in range(4):
                    h_gfm_a[l][i].copy_(ga_p[l][i], non_blocking=True)
                    if not self.fused:
                        h_gfm_b[l][i].copy_(gb_p[l][i], non_blocking=True)
        for s in (self.s_in, self.s_run, self.s_out):
            s.synchronize()
