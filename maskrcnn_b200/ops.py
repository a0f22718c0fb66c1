"""Host-side mirror of the reference's operator interface for the RoI hot path, on top of the C ABI
(include/mrcnn_b200.h).  PyTorch is used for device memory, streams and autograd plumbing only.

Reference interface being mirrored (paths relative to /root/reference):
    c++ext/maskrcnn/__init__.py:21-22   nms(dets, threshold)
    c++ext/maskrcnn/__init__.py:25-57   CropFunction(crop_height, crop_width, extrapolation_value=0)(image, boxes, box_ind)
    model.py:276-393                    roi_align(inputs, pool_size, image_shape)
    model.py:1307-1382                  MaskRCNN.rpn_refine(self, rpn_class, rpn_bbox)
    model.py:1389-1487                  MaskRCNN.mrn_refine(self, rpn_rois, probs, deltas, window)
    model.py:396-576                    mrn_samples(rpn_rois, gt_class_ids, gt_boxes, gt_masks, config)
    data.py:449-591                     rpn_samples(anchors, gt_class_ids, gt_boxes, config)
    data.py:287-314                     full_masks(class_id, boxes, masks, height, width)
    model.py:1294-1304 (+ :624-641)     MaskRCNN.rpn_detect(self, rpn_feature_maps)

There is no CPU implementation: CPU tensors raise TypeError, a missing library raises ImportError.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import NCHW, NHWC, check, lib

__all__ = ["nms", "CropFunction", "crop_and_resize", "pyramid_roi_align", "roi_align", "proposal_layer",
           "rpn_refine", "detection_layer", "mrn_refine", "detection_targets", "mrn_samples", "pyramid_roi_align_backward_pair", "pyramid_roi_align_pair", "rpn_samples", "full_masks", "decode_masks", "rpn_pack", "rpn_detect", "set_proposal_nms", "set_detection_nms", "check_device_errors",
           "set_backward_algorithm", "set_backward_planning", "set_deterministic"]


# Backward of the channels-last PyramidROIAlign: "auto" (MRCNN_BWD_AUTO: the row-owner gather whenever the call is
# eligible, else the scatter), "gather" (insist on the gather where eligible: no atomics on gradient data, no
# zero-fill pass, every gradient pixel written once) or "scatter" (clear + column-aggregated vector reductions).
BACKWARD_ALGORITHM = "auto"


def set_backward_algorithm(name):
    global BACKWARD_ALGORITHM
    if name not in ("auto", "gather", "scatter"):
        raise ValueError("backward algorithm must be 'auto', 'gather' or 'scatter'")
    BACKWARD_ALGORITHM = name


BACKWARD_PLANNING = True
_PLAN_STREAMS = {}


def set_backward_planning(enabled):
    """True (default): when a feature map requires grad, pyramid_roi_align's FORWARD also builds the work-item queues of
    its gather backward (they depend on the boxes only) on a side stream, concurrently with the forward kernel, and the
    backward is the one gather launch (mrcnn_pyramid_roi_align_backward_plan / _planned).  False: the backward builds
    them itself.  Results are identical."""
    global BACKWARD_PLANNING
    BACKWARD_PLANNING = bool(enabled)


def set_deterministic(enabled):
    """True: every gather-backward plan ends with a pass that orders each unit's work items, so the gradients of the gather
    backward are bit-reproducible from call to call (the default plan fills the queues through atomic cursors: same values to
    ~1e-7 relative, last bits depend on scheduling).  Process-wide; set it before the first backward (workspaces grow by a
    second item array).  The scatter backward (NCHW 14x14 by default, `set_backward_algorithm("scatter")`) sums through
    unordered reductions and is not covered."""
    check(lib.mrcnn_set_deterministic(1 if enabled else 0))


def _plan_stream(device):
    key = torch.device(device).index
    if key not in _PLAN_STREAMS:
        _PLAN_STREAMS[key] = torch.cuda.Stream(device=device, priority=-1)   # small kernels: let them slip in beside the forward
    return _PLAN_STREAMS[key]


def set_proposal_nms(name):
    """NMS inside the proposal layer: "auto" (hybrid when post_nms <= 2048), "mask" (IoU-bitmask tiles + sweep), "lazy"
    (chunks of 64 boxes against the survivors so far, stops at the post_nms-th survivor) or "hybrid" (the first 1.25 post_nms
    boxes resolved at once by a grid-wide fixed-point iteration, the lazy kernel for what is left).  Identical results."""
    algos = {"auto": 0, "mask": 1, "lazy": 2, "hybrid": 3}
    if name not in algos:
        raise ValueError("proposal NMS must be 'auto', 'mask', 'lazy' or 'hybrid'")
    check(lib.mrcnn_set_proposal_nms(algos[name]))


def set_detection_nms(name):
    """NMS inside the detection layer: "auto" (lazy when max_instances <= 1024), "mask" (class-aware IoU words + sweep) or
    "lazy" (chunks of 64 boxes against the same-class survivors so far, stops at the max_instances-th survivor).
    Identical results."""
    algos = {"auto": 0, "mask": 1, "lazy": 2}
    if name not in algos:
        raise ValueError("detection NMS must be 'auto', 'mask' or 'lazy'")
    check(lib.mrcnn_set_detection_nms(algos[name]))


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t, name, dtype=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise TypeError("%s must be a CUDA tensor: maskrcnn_b200 has no CPU path (got device=%s)" % (name, t.device))
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must have dtype %s (got %s)" % (name, dtype, t.dtype))
    return t


def _layout4(t):
    """(tensor, layout) with the tensor dense in NCHW or channels-last memory order."""
    if t.dim() != 4:
        raise ValueError("expected a 4-D tensor, got %d-D" % t.dim())
    if t.is_contiguous():
        # [N,C,1,1] is dense in both orders: call it channels-last so that it can take the vectorised kernels
        if t.size(2) == 1 and t.size(3) == 1 and t.size(1) % 4 == 0 and t.is_contiguous(memory_format=torch.channels_last):
            return t, NHWC
        return t, NCHW
    if t.is_contiguous(memory_format=torch.channels_last):
        return t, NHWC
    return t.contiguous(), NCHW


def _empty4(shape, layout, like):
    mf = torch.channels_last if layout == NHWC else torch.contiguous_format
    return torch.empty(shape, dtype=torch.float32, device=like.device, memory_format=mf)


def _ptr(t):
    return t.data_ptr() if t is not None else None


class _NoGuard(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NO_GUARD = _NoGuard()


def _on(device):
    """Device guard for a launch: torch.cuda.device(...) costs several microseconds per use, more than the argument checks of
    a small call; when the tensor already lives on the current device (the single-GPU-per-process case) nothing is switched."""
    if device.type != "cuda" or device.index == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)


_GEOM = {}


def _geometry(fms):
    """(Hs, Ws) ctypes arrays of a pyramid, cached by shape: a batch-1 inference call issues a ~20 us kernel, so building two
    ctypes arrays per call is measurable."""
    key = (fms[0].shape[2], fms[0].shape[3], fms[1].shape[2], fms[1].shape[3], fms[2].shape[2], fms[2].shape[3], fms[3].shape[2],
           fms[3].shape[3])
    g = _GEOM.get(key)
    if g is None:
        g = _GEOM[key] = (_lib.i4(key[0::2]), _lib.i4(key[1::2]))
    return g


def check_device_errors():
    """Synchronises and raises if a kernel saw a box_index outside [0, batch) since the last call
    (the reference exit(-1)s, cpu/crop_cpu.cpp:47-50)."""
    check(lib.mrcnn_poll_device_errors(_stream()))


# ------------------------------------------------------------------------------------------------
# nms
# ------------------------------------------------------------------------------------------------
def nms(dets, threshold):
    """c++ext/maskrcnn/__init__.py:21-22.  dets [N,5] = (y1,x1,y2,x2,score) -> int64 [K] ascending indices.
    One 4-byte device->host read (K) is the only synchronisation."""
    _require_cuda(dets, "dets", torch.float32)
    if dets.numel() == 0:
        # nms.h:20-21: the reference returns an empty CPU tensor here; a CUDA one indexes equally well
        return torch.empty(0, dtype=torch.int64, device=dets.device)
    if dets.dim() != 2 or dets.size(1) != 5:
        raise ValueError("dets must be [N,5]")
    dets = dets.contiguous()
    n = dets.size(0)
    with torch.cuda.device(dets.device):
        keep = torch.empty(n, dtype=torch.int64, device=dets.device)
        count = torch.empty(1, dtype=torch.int32, device=dets.device)
        ws_bytes = lib.mrcnn_nms_workspace_bytes(n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dets.device)
        check(lib.mrcnn_nms(dets.data_ptr(), n, float(threshold), keep.data_ptr(), count.data_ptr(), ws.data_ptr(),
                            ws_bytes, _stream()))
        k = int(count.item())
    return keep[:k]


# ------------------------------------------------------------------------------------------------
# crop_and_resize
# ------------------------------------------------------------------------------------------------
def _crop_forward(image, boxes, box_ind, ch, cw, ev):
    image, il = _layout4(image)
    B, C, H, W = image.shape
    N = boxes.size(0)
    out = _empty4((N, C, ch, cw), il, image)
    if N:
        with torch.cuda.device(image.device):
            check(lib.mrcnn_crop_forward(image.data_ptr(), B, C, H, W, il, boxes.data_ptr(), box_ind.data_ptr(), N,
                                         float(ev), ch, cw, out.data_ptr(), il, _stream()))
    return out


def _crop_backward(grad, boxes, box_ind, im_size, image_layout):
    grad, gl = _layout4(grad)
    B, C, H, W = im_size
    N, _, ch, cw = grad.shape
    gimg = _empty4((B, C, H, W), image_layout, grad)
    with torch.cuda.device(grad.device):
        check(lib.mrcnn_crop_backward(_ptr(grad) if N else None, gl, _ptr(boxes) if N else None,
                                      _ptr(box_ind) if N else None, N, ch, cw, gimg.data_ptr(), B, C, H, W,
                                      image_layout, 1, _stream()))
    return gimg


class _CropAndResize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, boxes, box_ind, ch, cw, ev):
        boxes = boxes.contiguous()
        box_ind = box_ind.contiguous()
        _, il = _layout4(image)
        ctx.im_size = tuple(image.shape)
        ctx.image_layout = il
        ctx.save_for_backward(boxes, box_ind)
        return _crop_forward(image, boxes, box_ind, ch, cw, ev)

    @staticmethod
    def backward(ctx, grad_output):
        boxes, box_ind = ctx.saved_tensors
        return _crop_backward(grad_output, boxes, box_ind, ctx.im_size, ctx.image_layout), None, None, None, None, None


def crop_and_resize(image, boxes, box_ind, crop_height, crop_width, extrapolation_value=0.0):
    _require_cuda(image, "image", torch.float32)
    _require_cuda(boxes, "boxes", torch.float32)
    _require_cuda(box_ind, "box_ind", torch.int32)  # __init__.py:34-35
    if boxes.dim() != 2 or boxes.size(1) != 4 or box_ind.dim() != 1 or box_ind.size(0) != boxes.size(0):
        raise ValueError("boxes must be [N,4] and box_ind [N]")
    if not (torch.is_grad_enabled() and image.requires_grad):   # inference: the same launch without the autograd node
        return _crop_forward(image, boxes.contiguous(), box_ind.contiguous(), int(crop_height), int(crop_width), float(extrapolation_value))
    return _CropAndResize.apply(image, boxes, box_ind, int(crop_height), int(crop_width), float(extrapolation_value))


class CropFunction(object):
    """Drop-in for c++ext/maskrcnn/__init__.py:25-57: construct with the crop size, call with
    (image [B,C,H,W] fp32, boxes [N,4] normalised fp32, box_ind [N] int32) -> crops [N,C,h,w].
    Differentiable w.r.t. image only.  (The reference class is a legacy autograd Function that torch 2.x
    refuses to run; instances of this class are plain callables over a static Function.)"""

    def __init__(self, crop_height, crop_width, extrapolation_value=0):
        self.crop_height = crop_height
        self.crop_width = crop_width
        self.extrapolation_value = extrapolation_value

    def __call__(self, image, boxes, box_ind):
        return crop_and_resize(image, boxes, box_ind, self.crop_height, self.crop_width, self.extrapolation_value)

    forward = __call__


# ------------------------------------------------------------------------------------------------
# PyramidROIAlign
# ------------------------------------------------------------------------------------------------
def _check_pyramid_args(feature_maps, boxes, box_ind):
    """Shapes the kernels rely on without being able to see them: one (B, C) for the four levels, one index per box.
    Written for the host-bound small calls (batch-1 inference issues a ~20 us kernel): the passing case touches each tensor
    once; anything unexpected goes through _require_cuda for the message."""
    if len(feature_maps) != 4:
        raise ValueError("expected the four pyramid levels P2..P5")
    f32 = torch.float32
    B = C = dev = None
    for i, f in enumerate(feature_maps):
        if not (isinstance(f, torch.Tensor) and f.is_cuda and f.dtype is f32):
            _require_cuda(f, "feature_maps[%d]" % i, f32)
        if f.dim() != 4:
            raise ValueError("feature_maps must be four [B,C,H_l,W_l] tensors with the same B and C on one device")
        if i == 0:
            B, C, dev = f.size(0), f.size(1), f.get_device()
        elif f.size(0) != B or f.size(1) != C or f.get_device() != dev:
            raise ValueError("feature_maps must be four [B,C,H_l,W_l] tensors with the same B and C on one device")
    if not (isinstance(boxes, torch.Tensor) and boxes.is_cuda and boxes.dtype is f32):
        _require_cuda(boxes, "boxes", f32)
    if boxes.dim() != 2 or boxes.size(1) != 4:
        raise ValueError("boxes must be [N,4]")
    if box_ind is not None:
        if not (isinstance(box_ind, torch.Tensor) and box_ind.is_cuda and box_ind.dtype is torch.int32):
            _require_cuda(box_ind, "box_ind", torch.int32)
        if box_ind.dim() != 1 or box_ind.size(0) != boxes.size(0):
            raise ValueError("box_ind must be int32 [N], one image index per box")


def _pyramid_layout(fms):
    if len(fms) != 4:
        raise ValueError("expected the four pyramid levels P2..P5")
    cl = torch.channels_last
    # the two common cases first: every level channels-last (and not also NCHW-dense), or every level NCHW-contiguous
    if all(f.is_contiguous(memory_format=cl) and not f.is_contiguous() for f in fms):
        return list(fms), NHWC
    if all(f.is_contiguous() and not (f.size(2) == 1 and f.size(3) == 1) for f in fms):
        return list(fms), NCHW
    outs, layouts = zip(*[_layout4(f) for f in fms])
    outs = list(outs)
    if len(set(layouts)) != 1:  # mixed: settle on channels-last (the fast path)
        outs = [f.contiguous(memory_format=torch.channels_last) for f in outs]
        return outs, NHWC
    # a [B,1,H,W] or [B,C,1,1] tensor is dense in both orders; prefer channels-last when any level says so
    return outs, layouts[0]


class _PyramidRoiAlign(torch.autograd.Function):
    @staticmethod
    def forward(ctx, boxes, box_ind, pool, image_area, out_layout, offsets, p2, p3, p4, p5):
        fms, fl = _pyramid_layout([p2, p3, p4, p5])
        B, C = fms[0].shape[:2]
        N = boxes.size(0)
        ol = fl if out_layout is None else out_layout
        out = _empty4((N, C, pool, pool), ol, fms[0])
        Hs = [f.shape[2] for f in fms]
        Ws = [f.shape[3] for f in fms]
        ctx.save_for_backward(boxes, box_ind)
        ctx.meta = (Hs, Ws, B, C, fl, pool, float(image_area), offsets)
        ctx.plan = None
        if (BACKWARD_PLANNING and BACKWARD_ALGORITHM != "scatter" and any(ctx.needs_input_grad[6:]) and fl == NHWC and ol == NHWC
                and C % 4 == 0 and N > 0 and N * pool * pool * C < 2 ** 31 and not offsets):
            # the backward's item queues need the boxes only: build them on a side stream NEXT TO the forward kernel.  The side
            # stream forks here, before the forward is enqueued, so it waits for the boxes (and the previous user of the
            # workspace block) and not for the forward itself.
            with torch.cuda.device(out.device):
                cur, side = torch.cuda.current_stream(), _plan_stream(out.device)
                ws_bytes = lib.mrcnn_pyramid_roi_align_backward_workspace_bytes(_lib.i4(Hs), _lib.i4(Ws), B, N, pool)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=out.device)
                side.wait_stream(cur)
                check(lib.mrcnn_pyramid_roi_align_backward_plan(_lib.i4(Hs), _lib.i4(Ws), B, C, boxes.data_ptr(), _ptr(box_ind), N,
                                                                pool, float(image_area), ws.data_ptr(), ws_bytes, side.cuda_stream))
                for t in (ws, boxes) + ((box_ind,) if box_ind is not None else ()):
                    t.record_stream(side)
                ctx.plan = (ws, ws_bytes, side.record_event())
        if N:
            with torch.cuda.device(out.device):
                check(lib.mrcnn_pyramid_roi_align_forward(_lib.vp4([f.data_ptr() for f in fms]), _lib.i4(Hs), _lib.i4(Ws),
                                                          B, C, fl, boxes.data_ptr(), _ptr(box_ind), N, pool,
                                                          float(image_area), out.data_ptr(), ol, None, _stream()))
        return out

    @staticmethod
    def backward(ctx, grad):
        boxes, box_ind = ctx.saved_tensors
        Hs, Ws, B, C, fl, pool, image_area, offsets = ctx.meta
        grad, gl = _layout4(grad)
        N = boxes.size(0)
        gfm = [_empty4((B, C, h, w), fl, grad) for h, w in zip(Hs, Ws)]
        if ctx.plan is not None and gl == NHWC and grad.data_ptr() % 16 == 0:
            ws, ws_bytes, ready = ctx.plan
            with torch.cuda.device(grad.device):
                torch.cuda.current_stream().wait_event(ready)
                check(lib.mrcnn_pyramid_roi_align_backward_planned(grad.data_ptr(), _lib.i4(Hs), _lib.i4(Ws), B, C, N, pool,
                                                                   _lib.vp4([g.data_ptr() for g in gfm]), 1, ws.data_ptr(), ws_bytes,
                                                                   _stream()))
            return (None, None, None, None, None, None) + tuple(gfm)
        with torch.cuda.device(grad.device):
            # the gather serves every layout combination (NCHW gradient maps sector-wise, NCHW upstream gradients through a
            # transposed copy in its workspace); the scatter remains for C % 4 != 0, huge RoI sets and the image-by-image mode
            gather_ok = C % 4 == 0 and N > 0 and N * pool * pool * C < 2 ** 31 and not offsets
            algo = {"auto": _lib.BWD_AUTO, "gather": _lib.BWD_GATHER if gather_ok else _lib.BWD_SCATTER,
                    "scatter": _lib.BWD_SCATTER}[BACKWARD_ALGORITHM]
            ws_bytes, ws = 0, None
            if gather_ok and algo != _lib.BWD_SCATTER:
                ws_bytes = lib.mrcnn_pyramid_roi_align_backward_workspace_bytes_ex(_lib.i4(Hs), _lib.i4(Ws), B, C, N, pool, gl)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=grad.device)
            check(lib.mrcnn_pyramid_roi_align_backward(_ptr(grad) if N else None, gl, _lib.i4(Hs), _lib.i4(Ws), B, C,
                                                       _ptr(boxes) if N else None, _ptr(box_ind), N, pool, image_area,
                                                       _lib.vp4([g.data_ptr() for g in gfm]), fl, 1,
                                                       ctypes.cast(_lib.i32_array(offsets), ctypes.c_void_p) if offsets else None,
                                                       algo, _ptr(ws), ws_bytes, _stream()))
        return (None, None, None, None, None, None) + tuple(gfm)


def pyramid_roi_align(feature_maps, boxes, box_ind, pool_size, image_shape, out_channels_last=None,
                      rois_per_image=None):
    """Batched PyramidROIAlign.  feature_maps: [P2,P3,P4,P5], each [B,C,Hl,Wl] (NCHW or channels-last; the
    channels-last layout runs the 128-bit vectorised kernels).  boxes [N,4] normalised (not differentiated,
    model.py:358), box_ind [N] int32 image index or None (all image 0).  Returns [N,C,pool,pool] in box order;
    memory format follows the feature maps unless out_channels_last is given.
    rois_per_image: optional host-side list of B counts (or one int) stating that boxes are grouped by image in
    order; box_ind is then derived from the counts and must not be passed as well; lets the backward clear + scatter image
    by image (L2-resident)."""
    feature_maps = list(feature_maps)
    _check_pyramid_args(feature_maps, boxes, box_ind)
    if boxes.requires_grad:
        boxes = boxes.detach()
    if not boxes.is_contiguous():
        boxes = boxes.contiguous()
    if box_ind is not None and not box_ind.is_contiguous():
        box_ind = box_ind.contiguous()
    image_area = float(image_shape[0] * image_shape[1])  # model.py:331
    ol = None if out_channels_last is None else (NHWC if out_channels_last else NCHW)
    offsets = None
    if rois_per_image is not None:
        B = feature_maps[0].size(0)
        counts = [int(rois_per_image)] * B if np.isscalar(rois_per_image) else [int(v) for v in rois_per_image]
        if len(counts) != B or sum(counts) != boxes.size(0):
            raise ValueError("rois_per_image must list one count per image and sum to the number of boxes")
        offsets = tuple(int(v) for v in np.concatenate([[0], np.cumsum(counts)]))
        if box_ind is not None:
            # the image-by-image backward scatters box i into the image the counts say, the forward reads the image box_ind
            # says: one source of truth, or a mismatch would silently put gradients into another image than the forward read
            raise ValueError("give either box_ind or rois_per_image (which implies box_ind = repeat(arange(B), counts)), not both")
        if B > 1:
            box_ind = torch.repeat_interleave(torch.arange(B, dtype=torch.int32, device=boxes.device),
                                              torch.tensor(counts, device=boxes.device))
    if not (torch.is_grad_enabled() and any(f.requires_grad for f in feature_maps)):
        # nothing to differentiate (inference, torch.no_grad): the same launch without the autograd node around it
        return _pyramid_forward_nograd(boxes, box_ind, int(pool_size), image_area, ol, feature_maps)
    return _PyramidRoiAlign.apply(boxes, box_ind, int(pool_size), image_area, ol, offsets, *feature_maps)


def _pyramid_forward_nograd(boxes, box_ind, pool, image_area, out_layout, feature_maps):
    """_PyramidRoiAlign.forward without the bookkeeping of a backward that will not run."""
    fms, fl = _pyramid_layout(feature_maps)
    f0 = fms[0]
    B, C = f0.size(0), f0.size(1)
    N = boxes.size(0)
    ol = fl if out_layout is None else out_layout
    out = torch.empty((N, C, pool, pool), dtype=torch.float32, device=f0.device,
                      memory_format=torch.channels_last if ol == NHWC else torch.contiguous_format)
    if N:
        Hs, Ws = _geometry(fms)
        with _on(f0.device):
            rc = lib.mrcnn_pyramid_roi_align_forward(_lib._vp4(fms[0].data_ptr(), fms[1].data_ptr(), fms[2].data_ptr(), fms[3].data_ptr()),
                                                     Hs, Ws, B, C, fl, boxes.data_ptr(), None if box_ind is None else box_ind.data_ptr(),
                                                     N, pool, image_area, out.data_ptr(), ol, None, _stream())
        if rc:
            check(rc)
    return out


def roi_align(inputs, pool_size, image_shape):
    """Drop-in for model.py:276-393: inputs = [boxes [1,N,4], P2 [1,C,H,W], P3, P4, P5] -> [N,C,pool,pool].
    One fused launch instead of ~30 (per-level nonzero/gather/crop + cat + sort + gather)."""
    boxes = inputs[0]
    if boxes.dim() == 3:
        if boxes.size(0) != 1:
            raise ValueError("roi_align drop-in is batch-1 like the reference (model.py:296); use pyramid_roi_align")
        boxes = boxes[0]
    fms = [f if f.dim() == 4 else f.unsqueeze(0) for f in inputs[1:5]]
    return pyramid_roi_align(fms, boxes, None, pool_size, image_shape)


# ------------------------------------------------------------------------------------------------
# proposal layer
# ------------------------------------------------------------------------------------------------
_PROPOSAL_WS = {}   # (B, A, pre_nms) -> workspace bytes
_STD4 = {}          # std tuple -> ctypes float[4]


def proposal_layer(rpn_class, rpn_bbox, anchors, pre_nms_limit, post_nms_limit, nms_threshold,
                   std=(0.1, 0.1, 0.2, 0.2), image_hw=(1024, 1024)):
    """Batched proposal layer.  rpn_class [B,A,2] - or the foreground probabilities alone, [B,A], as rpn_pack returns
    them - rpn_bbox [B,A,4], anchors [A,4] px -> (rois [B,post,4] normalised & zero padded, counts int32 [B]).  No host
    synchronisation."""
    _require_cuda(rpn_class, "rpn_class", torch.float32)
    _require_cuda(rpn_bbox, "rpn_bbox", torch.float32)
    _require_cuda(anchors, "anchors", torch.float32)
    fg_only = rpn_class.dim() == 2
    if (not fg_only and (rpn_class.dim() != 3 or rpn_class.size(2) != 2)) or rpn_bbox.shape != rpn_class.shape[:2] + (4,):
        raise ValueError("rpn_class must be [B,A,2] (or fg scores [B,A]) and rpn_bbox [B,A,4]")
    B, A = rpn_class.shape[:2]
    if anchors.shape != (A, 4):
        raise ValueError("anchors must be [A,4]")
    rpn_class, rpn_bbox, anchors = rpn_class.contiguous(), rpn_bbox.contiguous(), anchors.contiguous()
    post = int(post_nms_limit)
    with _on(rpn_class.device):   # the layer is ~65 us of kernels for a batch of 8: the Python around the call is kept short
        rois = torch.empty((B, post, 4), dtype=torch.float32, device=rpn_class.device)
        counts = torch.empty(B, dtype=torch.int32, device=rpn_class.device)
        key = (B, A, int(pre_nms_limit))
        ws_bytes = _PROPOSAL_WS.get(key)
        if ws_bytes is None:
            ws_bytes = _PROPOSAL_WS[key] = lib.mrcnn_proposal_workspace_bytes(*key)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=rpn_class.device)
        std = tuple(float(v) for v in std)
        std4 = _STD4.get(std)
        if std4 is None:
            std4 = _STD4[std] = _lib.f4(np.float32(std))
        fn = lib.mrcnn_proposal_layer_fg if fg_only else lib.mrcnn_proposal_layer
        check(fn(rpn_class.data_ptr(), rpn_bbox.data_ptr(), anchors.data_ptr(), B, A,
                 int(pre_nms_limit), post, float(nms_threshold), std4,
                 float(image_hw[0]), float(image_hw[1]), rois.data_ptr(), counts.data_ptr(),
                 ws.data_ptr(), ws_bytes, _stream()))
    return rois, counts


# ------------------------------------------------------------------------------------------------
# RPN head output plumbing
# ------------------------------------------------------------------------------------------------
def _rpn_levels(logits, bboxes):
    if len(logits) != len(bboxes) or not 1 <= len(logits) <= 8:
        raise ValueError("rpn_pack takes 1..8 pyramid levels, one class and one bbox conv output each")
    B = logits[0].size(0)
    K, rem = divmod(logits[0].size(1), 2)
    layouts = set()
    ls, bs = [], []
    for lg, bx in zip(logits, bboxes):
        _require_cuda(lg, "rpn class logits", torch.float32)
        _require_cuda(bx, "rpn bbox", torch.float32)
        if lg.dim() != 4 or rem or lg.size(0) != B or lg.size(1) != 2 * K or bx.shape != (B, 4 * K) + tuple(lg.shape[2:]):
            raise ValueError("per level: class logits [B,2K,H,W] and bbox [B,4K,H,W]")
        lg, la = _layout4(lg)
        bx, lb = _layout4(bx)
        ls.append(lg)
        bs.append(bx)
        layouts.update((la, lb))
    if len(layouts) != 1:   # mixed memory formats: fall back to the conv default for all of them
        ls = [t.contiguous() for t in ls]
        bs = [t.contiguous() for t in bs]
        layouts = {NCHW}
    return ls, bs, B, K, layouts.pop()


def _ptr_array(ts):
    return (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


def _int_array(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


class _RpnPack(torch.autograd.Function):
    @staticmethod
    def forward(ctx, n_levels, want_class, *tensors):
        ls, bs, B, K, layout = _rpn_levels(tensors[:n_levels], tensors[n_levels:])
        Hs, Ws = [t.size(2) for t in ls], [t.size(3) for t in ls]
        A = K * sum(h * w for h, w in zip(Hs, Ws))
        dev = ls[0].device
        o_logits = torch.empty((B, A, 2), dtype=torch.float32, device=dev)
        o_class = torch.empty((B, A, 2), dtype=torch.float32, device=dev) if want_class else None
        o_bbox = torch.empty((B, A, 4), dtype=torch.float32, device=dev)
        o_fg = torch.empty((B, A), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.mrcnn_rpn_pack(_ptr_array(ls), _ptr_array(bs), _int_array(Hs), _int_array(Ws), n_levels, B, K, layout,
                                     o_logits.data_ptr(), _ptr(o_class), o_bbox.data_ptr(), o_fg.data_ptr(), _stream()))
        ctx.geom = (Hs, Ws, B, K, layout, n_levels)
        ctx.set_materialize_grads(False)   # an output the loss does not read costs no unpack pass
        if o_class is None:
            o_class = o_fg.new_empty(0)
        ctx.mark_non_differentiable(o_class, o_fg)   # the probabilities only feed the proposal layer (no gradient path)
        return o_logits, o_class, o_bbox, o_fg

    @staticmethod
    def backward(ctx, g_logits, _g_class, g_bbox, _g_fg):
        Hs, Ws, B, K, layout, n = ctx.geom
        need_l = any(ctx.needs_input_grad[2:2 + n]) and g_logits is not None
        need_b = any(ctx.needs_input_grad[2 + n:]) and g_bbox is not None
        if not need_l and not need_b:
            return (None, None) + (None,) * (2 * n)
        dev = (g_logits if need_l else g_bbox).device
        mf = torch.channels_last if layout == NHWC else torch.contiguous_format
        gl = [torch.empty((B, 2 * K, h, w), dtype=torch.float32, device=dev, memory_format=mf) for h, w in zip(Hs, Ws)] if need_l else None
        gb = [torch.empty((B, 4 * K, h, w), dtype=torch.float32, device=dev, memory_format=mf) for h, w in zip(Hs, Ws)] if need_b else None
        with torch.cuda.device(dev):
            check(lib.mrcnn_rpn_unpack(g_logits.contiguous().data_ptr() if need_l else None,
                                       g_bbox.contiguous().data_ptr() if need_b else None, _int_array(Hs), _int_array(Ws), n, B, K,
                                       layout, _ptr_array(gl) if need_l else None, _ptr_array(gb) if need_b else None, _stream()))
        none = [None] * n
        return (None, None) + tuple(gl if need_l else none) + tuple(gb if need_b else none)


def rpn_pack(class_logits, bboxes, want_class=True):
    """One launch for everything between the RPN head's two 1x1 convolutions and the proposal layer: class_logits[l]
    [B,2K,H_l,W_l] and bboxes[l] [B,4K,H_l,W_l] (conv outputs of every pyramid level, NCHW or channels-last) ->
    (rpn_class_logits [B,A,2], rpn_class [B,A,2] = softmax, rpn_bbox [B,A,4], fg [B,A] = rpn_class[..., 1]) in the
    reference's anchor order (model.py:624-641 per level, torch.cat over levels :1294-1304).  Differentiable w.r.t. the
    conv outputs through rpn_class_logits and rpn_bbox (what the RPN losses read)."""
    class_logits, bboxes = list(class_logits), list(bboxes)
    out = _RpnPack.apply(len(class_logits), bool(want_class), *class_logits, *bboxes)
    return out[0], (out[1] if want_class else None), out[2], out[3]


def rpn_detect(self, rpn_feature_maps):
    """Drop-in for MaskRCNN.rpn_detect (model.py:1294-1304): runs the RPN head's three convolutions per level on stock
    PyTorch / cuDNN (self.rpn: padding, conv_shared, relu, conv_class, conv_bbox - model.py:600-641) and replaces the
    per-level permute / contiguous / view / softmax and the three torch.cat by one rpn_pack launch.  Returns
    (rpn_class_logits, rpn_class, rpn_bbox) like the reference.

    DIFFERENCE FROM THE REFERENCE: rpn_class (the softmax) is returned without a gradient path - requires_grad is False.  In
    the reference it is an ordinary differentiable tensor, but nothing differentiates it: the RPN losses read
    rpn_class_logits and rpn_bbox (model.py:652-718) and the proposal layer detaches.  A loss built on rpn_class would get no
    gradient here; build it on rpn_class_logits (torch.softmax(rpn_class_logits, 2) is differentiable and equal)."""
    rpn = self.rpn
    logits, bboxes = [], []
    for p in rpn_feature_maps:
        x = rpn.relu(rpn.conv_shared(rpn.padding(p)))
        logits.append(rpn.conv_class(x))
        bboxes.append(rpn.conv_bbox(x))
    rpn_class_logits, rpn_class, rpn_bbox, _ = rpn_pack(logits, bboxes)
    return rpn_class_logits, rpn_class, rpn_bbox


def rpn_refine(self, rpn_class, rpn_bbox, pre_nms_limit=None):
    """Drop-in for MaskRCNN.rpn_refine (model.py:1307-1382): returns [1,K,4] normalised proposals.
    pre_nms_limit defaults to the fork's hard-coded 500 (model.py:1345)."""
    cfg = self.config
    pre = min(500 if pre_nms_limit is None else pre_nms_limit, self.anchors.size(0))
    h, w = (int(v) for v in cfg.IMAGE_SHAPE[:2])
    rois, counts = proposal_layer(rpn_class, rpn_bbox, self.anchors, pre, cfg.RPN_NMS_MAX_ROIS_NUM,
                                  cfg.RPN_NMS_THRESHOLD, cfg.RPN_BBOX_STD_DEV, (h, w))
    k = int(counts[0].item())  # the reference's output is not padded -> one 4-byte read
    return rois[:1, :k]


# ------------------------------------------------------------------------------------------------
# detection layer
# ------------------------------------------------------------------------------------------------
def detection_layer(rois, probs, deltas, windows, min_confidence, nms_threshold, max_instances,
                    std=(0.1, 0.1, 0.2, 0.2), image_hw=(1024, 1024), return_index=False):
    """Batched detection layer.  rois [B,N,4] normalised, probs [B,N,NC], deltas [B,N,NC,4], windows [B,4] px ->
    (dets [B,D,6] = (y1,x1,y2,x2,score,class) score-descending zero padded, counts int32 [B][, index int32 [B,D]])."""
    for t, name in ((rois, "rois"), (probs, "probs"), (deltas, "deltas"), (windows, "windows")):
        _require_cuda(t, name, torch.float32)
    if rois.dim() != 3 or rois.size(2) != 4:
        raise ValueError("rois must be [B,N,4]")
    B, N = rois.shape[:2]
    NC = probs.size(-1)
    if probs.shape != (B, N, NC) or deltas.shape != (B, N, NC, 4) or windows.shape != (B, 4):
        raise ValueError("probs must be [B,N,NC], deltas [B,N,NC,4], windows [B,4]")
    rois, probs, deltas, windows = rois.contiguous(), probs.contiguous(), deltas.contiguous(), windows.contiguous()
    D = int(max_instances)
    with torch.cuda.device(rois.device):
        dets = torch.empty((B, D, 6), dtype=torch.float32, device=rois.device)
        counts = torch.empty(B, dtype=torch.int32, device=rois.device)
        index = torch.empty((B, D), dtype=torch.int32, device=rois.device) if return_index else None
        ws_bytes = lib.mrcnn_detection_workspace_bytes(B, N)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=rois.device)
        check(lib.mrcnn_detection_layer(rois.data_ptr(), probs.data_ptr(), deltas.data_ptr(), windows.data_ptr(), B, N, NC,
                                        float(min_confidence or 0.0), float(nms_threshold), D, _lib.f4(np.float32(std)),
                                        float(image_hw[0]), float(image_hw[1]), dets.data_ptr(), counts.data_ptr(),
                                        _ptr(index), ws.data_ptr(), ws_bytes, _stream()))
    if return_index:
        return dets, counts, index
    return dets, counts


def mrn_refine(self, rpn_rois, probs, deltas, window):
    """Drop-in for MaskRCNN.mrn_refine (model.py:1389-1487): returns (class_ids [1,D] int64, scores [1,D],
    boxes [1,D,4]) or (None, None, None) when nothing survives (:1445-1447)."""
    cfg = self.config
    dev = probs.device
    win = torch.as_tensor(np.asarray(window, dtype=np.float32).reshape(1, 4)).to(dev) if not isinstance(window, torch.Tensor) \
        else window.to(device=dev, dtype=torch.float32).reshape(1, 4)
    h, w = (int(v) for v in cfg.IMAGE_SHAPE[:2])
    rois = rpn_rois if rpn_rois.dim() == 3 else rpn_rois.unsqueeze(0)
    dets, counts = detection_layer(rois, probs.unsqueeze(0), deltas.unsqueeze(0), win, cfg.DETECTION_MIN_CONFIDENCE,
                                   cfg.DETECTION_NMS_THRESHOLD, cfg.DETECTION_MAX_INSTANCES,
                                   np.asarray(cfg.RPN_BBOX_STD_DEV, dtype=np.float32).reshape(4), (h, w))
    d = int(counts[0].item())
    if d < 1:
        return None, None, None
    dets = dets[0, :d]
    return dets[:, 5].long().unsqueeze(0), dets[:, 4].unsqueeze(0), dets[:, :4].unsqueeze(0)


def pyramid_roi_align_backward_pair(grad_a, grad_b, feature_shapes, boxes, box_ind, image_shape, out=None, accumulate=False):
    """Both heads' RoIAlign backward in one pass: grad_a [N,C,pa,pa] and grad_b [N,C,pb,pb] (channels-last), the same
    RoIs -> the SUM of the two gradient pyramids, [B,C,H_l,W_l] channels-last per level (what autograd accumulates after
    the reference's per-head, per-level CropFunction.backward calls, model.py:778 + :889).  feature_shapes = the four
    (B, C, H, W); `out` = four preallocated channels-last tensors (optional); accumulate adds to them."""
    _require_cuda(grad_a, "grad_a", torch.float32)
    _require_cuda(grad_b, "grad_b", torch.float32)
    _require_cuda(boxes, "boxes", torch.float32)
    if not (grad_a.is_contiguous(memory_format=torch.channels_last) and grad_b.is_contiguous(memory_format=torch.channels_last)):
        raise ValueError("pyramid_roi_align_backward_pair needs channels-last gradients")
    N, C, pa = grad_a.shape[:3]
    pb = grad_b.size(2)
    if boxes.shape != (N, 4) or grad_b.shape[:2] != (N, C) or len(feature_shapes) != 4 or N == 0:
        raise ValueError("grad_a [N,C,pa,pa], grad_b [N,C,pb,pb], boxes [N,4] with N > 0, four feature shapes")
    B = int(feature_shapes[0][0])
    Hs, Ws = [int(s[2]) for s in feature_shapes], [int(s[3]) for s in feature_shapes]
    boxes = boxes.contiguous()
    if box_ind is not None:
        box_ind = _require_cuda(box_ind, "box_ind", torch.int32).contiguous()
        if box_ind.shape != (N,):
            raise ValueError("box_ind must be int32 [N], one image index per box")
    if out is not None:
        if len(out) != 4:
            raise ValueError("out must hold the four gradient levels")
        for g, hh, ww in zip(out, Hs, Ws):
            _require_cuda(g, "out", torch.float32)
            if g.shape != (B, C, hh, ww) or not g.is_contiguous(memory_format=torch.channels_last):
                raise ValueError("out[l] must be a channels-last [B,C,H_l,W_l] tensor matching feature_shapes")
    h, w = float(image_shape[0]), float(image_shape[1])
    with torch.cuda.device(grad_a.device):
        gfm = out if out is not None else [_empty4((B, C, hh, ww), NHWC, grad_a) for hh, ww in zip(Hs, Ws)]
        ws_bytes = lib.mrcnn_pyramid_roi_align_backward_pair_workspace_bytes(_lib.i4(Hs), _lib.i4(Ws), B, N, pa, pb)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=grad_a.device)
        check(lib.mrcnn_pyramid_roi_align_backward_pair(grad_a.data_ptr(), pa, grad_b.data_ptr(), pb, _lib.i4(Hs), _lib.i4(Ws), B, C,
                                                        boxes.data_ptr(), _ptr(box_ind), N, h * w,
                                                        _lib.vp4([g.data_ptr() for g in gfm]), 0 if accumulate else 1,
                                                        ws.data_ptr(), ws_bytes, _stream()))
    return gfm


class _PyramidRoiAlignPair(torch.autograd.Function):
    """Both heads' crops from one autograd node, so that their backward is ONE fused gather."""

    @staticmethod
    def forward(ctx, boxes, box_ind, pool_a, pool_b, image_area, p2, p3, p4, p5):
        fms, fl = _pyramid_layout([p2, p3, p4, p5])
        if fl != NHWC:
            raise ValueError("pyramid_roi_align_pair needs channels-last feature maps")
        B, C = fms[0].shape[:2]
        N = boxes.size(0)
        Hs = [f.shape[2] for f in fms]
        Ws = [f.shape[3] for f in fms]
        outs = []
        if {pool_a, pool_b} == {7, 14} and C % 4 == 0 and N > 0:
            # the two heads of Mask R-CNN: ONE launch (each CTA pools its RoI at 14x14, then at 7x7 from the footprint it just read)
            o7, o14 = _empty4((N, C, 7, 7), NHWC, fms[0]), _empty4((N, C, 14, 14), NHWC, fms[0])
            with torch.cuda.device(fms[0].device):
                check(lib.mrcnn_pyramid_roi_align_forward_pair(_lib.vp4([f.data_ptr() for f in fms]), _lib.i4(Hs), _lib.i4(Ws), B, C,
                                                               boxes.data_ptr(), _ptr(box_ind), N, float(image_area), o7.data_ptr(),
                                                               o14.data_ptr(), _stream()))
            outs = [o7, o14] if pool_a == 7 else [o14, o7]
        with torch.cuda.device(fms[0].device):
            for pool in (() if outs else (pool_a, pool_b)):
                out = _empty4((N, C, pool, pool), NHWC, fms[0])
                check(lib.mrcnn_pyramid_roi_align_forward(_lib.vp4([f.data_ptr() for f in fms]), _lib.i4(Hs), _lib.i4(Ws), B, C, fl,
                                                          boxes.data_ptr(), _ptr(box_ind), N, pool, float(image_area),
                                                          out.data_ptr(), NHWC, None, _stream()))
                outs.append(out)
        ctx.save_for_backward(boxes, box_ind)
        ctx.meta = ([tuple(f.shape) for f in fms], float(image_area), pool_a, pool_b)
        return tuple(outs)

    @staticmethod
    def backward(ctx, grad_a, grad_b):
        boxes, box_ind = ctx.saved_tensors
        shapes, image_area, pool_a, pool_b = ctx.meta
        N, C = boxes.size(0), shapes[0][1]
        cl = lambda g, p: (torch.zeros((N, C, p, p), device=boxes.device).contiguous(memory_format=torch.channels_last)  # noqa: E731
                           if g is None else g.contiguous(memory_format=torch.channels_last))
        gfm = pyramid_roi_align_backward_pair(cl(grad_a, pool_a), cl(grad_b, pool_b), shapes, boxes, box_ind, (image_area, 1.0))
        return (None, None, None, None, None) + tuple(gfm)


def pyramid_roi_align_pair(feature_maps, boxes, box_ind, pool_sizes, image_shape):
    """PyramidROIAlign of the same RoIs at two pool sizes (box head 7x7 + mask head 14x14, model.py:778 / :889) as ONE
    autograd node: returns (crops_a, crops_b), channels-last, and backpropagates both heads with the fused gather
    (one gradient-pyramid write instead of two plus autograd's add).  Channels-last feature maps, N > 0, C % 4 == 0."""
    feature_maps = list(feature_maps)
    _check_pyramid_args(feature_maps, boxes, box_ind)
    if boxes.size(0) == 0 or len(pool_sizes) != 2:
        raise ValueError("boxes must be [N,4] with N > 0 and pool_sizes a pair")
    boxes = boxes.detach().contiguous()
    if box_ind is not None:
        box_ind = box_ind.contiguous()
    image_area = float(image_shape[0] * image_shape[1])
    return _PyramidRoiAlignPair.apply(boxes, box_ind, int(pool_sizes[0]), int(pool_sizes[1]), image_area, *feature_maps)


# ------------------------------------------------------------------------------------------------
# detection-target layer
# ------------------------------------------------------------------------------------------------
def _target_classify(rois, gt_class_ids, gt_boxes, want_iou=False):
    B, N = rois.shape[:2]
    G = gt_boxes.size(1)
    dev = rois.device
    pos = torch.empty((B, max(N, 1)), dtype=torch.int32, device=dev)
    neg = torch.empty((B, max(N, 1)), dtype=torch.int32, device=dev)
    assign = torch.empty((B, max(N, 1)), dtype=torch.int32, device=dev)
    iou = torch.empty((B, max(N, 1)), dtype=torch.float32, device=dev) if want_iou else None
    counts = torch.empty((B, 2), dtype=torch.int32, device=dev)
    check(lib.mrcnn_target_classify(_ptr(rois) if N else None, _ptr(gt_boxes) if G else None, _ptr(gt_class_ids) if G else None,
                                    B, N, G, pos.data_ptr(), neg.data_ptr(), assign.data_ptr(), _ptr(iou), counts.data_ptr(),
                                    _stream()))
    return pos, neg, assign, iou, counts


def _target_emit(rois, gt_class_ids, gt_boxes, gt_masks, pos, neg, perm_pos, perm_neg, take, assign, std, mask_shape, T):
    B, N = rois.shape[:2]
    G, H, W = gt_masks.shape[1:]
    dev = rois.device
    mh, mw = int(mask_shape[0]), int(mask_shape[1])
    o_rois = torch.empty((B, T, 4), dtype=torch.float32, device=dev)
    o_cls = torch.empty((B, T), dtype=torch.int32, device=dev)
    o_d = torch.empty((B, T, 4), dtype=torch.float32, device=dev)
    o_m = torch.empty((B, T, mh, mw), dtype=torch.float32, device=dev)
    check(lib.mrcnn_target_emit(rois.data_ptr(), gt_boxes.data_ptr(), gt_class_ids.data_ptr(), gt_masks.data_ptr(), B, N, G, H, W,
                                pos.data_ptr(), neg.data_ptr(), _ptr(perm_pos), _ptr(perm_neg), take.data_ptr(), assign.data_ptr(),
                                _lib.f4(np.float32(std)), mh, mw, T, o_rois.data_ptr(), o_cls.data_ptr(), o_d.data_ptr(),
                                o_m.data_ptr(), _stream()))
    return o_rois, o_cls, o_d, o_m


def _negatives_for(pos_kept, ratio):
    """model.py:518-519 in Python doubles, like the reference."""
    r = 1.0 / ratio
    return int(r * pos_kept - pos_kept)


def detection_targets(rois, gt_class_ids, gt_boxes, gt_masks, keys_pos, keys_neg, train_rois_per_image=512,
                      roi_positive_ratio=0.33, std=(0.1, 0.1, 0.2, 0.2), mask_shape=(28, 28)):
    """Batched, sync-free detection-target layer.  rois [B,N,4] and gt_boxes [B,G,4] normalised, gt_class_ids int32
    [B,G], gt_masks fp32 [B,G,H,W]; keys_pos / keys_neg fp32 [B,N]: random keys that stand for the reference's two
    torch.randperm draws (perm = stable argsort of the first P / Q keys).  Returns zero-padded
    (rois [B,T,4], class_ids int32 [B,T], deltas [B,T,4], masks [B,T,mh,mw], take int32 [B,2] = kept positives /
    negatives) with T = train_rois_per_image; positives first, then negatives (model.py:524-541)."""
    for t, name in ((rois, "rois"), (gt_boxes, "gt_boxes"), (gt_masks, "gt_masks"), (keys_pos, "keys_pos"), (keys_neg, "keys_neg")):
        _require_cuda(t, name, torch.float32)
    _require_cuda(gt_class_ids, "gt_class_ids", torch.int32)
    B, N = rois.shape[:2]
    if rois.dim() != 3 or rois.size(2) != 4 or gt_boxes.dim() != 3 or gt_masks.dim() != 4 or keys_pos.shape != (B, N) \
            or keys_neg.shape != (B, N) or gt_class_ids.shape != gt_boxes.shape[:2] or gt_masks.shape[:2] != gt_boxes.shape[:2]:
        raise ValueError("rois [B,N,4], gt_class_ids [B,G], gt_boxes [B,G,4], gt_masks [B,G,H,W], keys [B,N]")
    rois, gt_boxes, gt_masks, gt_class_ids = rois.contiguous(), gt_boxes.contiguous(), gt_masks.contiguous(), gt_class_ids.contiguous()
    keys_pos, keys_neg = keys_pos.contiguous(), keys_neg.contiguous()
    T = int(train_rois_per_image)
    pos_cap = int(T * roi_positive_ratio)                                    # model.py:466-467
    with torch.cuda.device(rois.device):
        table = torch.tensor([_negatives_for(p, roi_positive_ratio) for p in range(pos_cap + 1)], dtype=torch.int32).to(rois.device)
        pos, neg, assign, _, counts = _target_classify(rois, gt_class_ids, gt_boxes)
        perm_pos, perm_neg = torch.empty_like(pos), torch.empty_like(neg)
        take = torch.empty((B, 2), dtype=torch.int32, device=rois.device)
        check(lib.mrcnn_target_select(counts.data_ptr(), keys_pos.data_ptr(), keys_neg.data_ptr(), table.data_ptr(), B, N, pos_cap,
                                      perm_pos.data_ptr(), perm_neg.data_ptr(), take.data_ptr(), _stream()))
        out = _target_emit(rois, gt_class_ids, gt_boxes, gt_masks, pos, neg, perm_pos, perm_neg, take, assign, std, mask_shape, T)
    return out + (take,)


def mrn_samples(rpn_rois, gt_class_ids, gt_boxes, gt_masks, config):
    """Drop-in for mrn_samples (model.py:396-576), batch 1 like the reference: rpn_rois [1,N,4], gt_class_ids [1,G],
    gt_boxes [1,G,4], gt_masks [1,G,H,W] -> (rois [T,4], class ids int32 [T], deltas [T,4], masks [T,mh,mw]).
    The two permutations are drawn with torch.randperm on the CPU generator in the reference's order (:468, :520), so
    results are identical to the reference's under the same torch seed; like the reference it reads the positive /
    negative counts back once."""
    dev = rpn_rois.device
    rois = rpn_rois.float().contiguous()
    cls = gt_class_ids.to(torch.int32).contiguous()
    gtb = gt_boxes.float().contiguous()
    masks = gt_masks.float().contiguous()
    _require_cuda(rois, "rpn_rois", torch.float32)
    if rois.dim() != 3 or rois.size(0) != 1 or masks.dim() != 4:
        raise ValueError("mrn_samples supports batch 1 like the reference: rpn_rois [1,N,4], gt_masks [1,G,H,W]")
    N = rois.size(1)
    ratio = float(config.ROI_POSITIVE_RATIO)
    empty = (torch.empty(0, device=dev), torch.empty(0, dtype=torch.int32, device=dev), torch.empty(0, device=dev),
             torch.empty(0, device=dev))                                    # model.py:563-574
    if N == 0 or gtb.size(1) == 0:
        return empty
    with torch.cuda.device(dev):
        pos, neg, assign, _, counts = _target_classify(rois, cls, gtb)
        P, Q = (int(v) for v in counts[0].tolist())                          # the reference syncs here too (nonzero)
        if P == 0:
            return empty
        perm_pos = torch.randperm(P)[:int(config.TRAIN_ROIS_PER_IMAGE * ratio)]          # :466-471
        tp = int(perm_pos.numel())
        tn, perm_neg = 0, None
        if Q > 0:
            perm_neg = torch.randperm(Q)[:_negatives_for(tp, ratio)]                         # :518-522
            tn = int(perm_neg.numel())
        pp = torch.zeros((1, N), dtype=torch.int32)
        pp[0, :tp] = perm_pos.to(torch.int32)
        pn = torch.zeros((1, N), dtype=torch.int32)
        if tn:
            pn[0, :tn] = perm_neg.to(torch.int32)
        take = torch.tensor([[tp, tn]], dtype=torch.int32)
        o_rois, o_cls, o_d, o_m = _target_emit(rois, cls, gtb, masks, pos, neg, pp.to(dev), pn.to(dev), take.to(dev), assign,
                                               np.asarray(config.BBOX_STD_DEV, dtype=np.float32).reshape(4), config.MASK_SHAPE,
                                               tp + tn)
    return o_rois[0], o_cls[0], o_d[0], o_m[0]


# ------------------------------------------------------------------------------------------------
# RPN anchor matching
# ------------------------------------------------------------------------------------------------
def _compact_equal(values, target):
    n = values.numel()
    ids = torch.empty(n, dtype=torch.int32, device=values.device)
    count = torch.empty(1, dtype=torch.int32, device=values.device)
    ws_bytes = lib.mrcnn_compact_equal_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=values.device)
    check(lib.mrcnn_compact_equal(values.data_ptr(), n, int(target), ids.data_ptr(), count.data_ptr(), ws.data_ptr(), ws_bytes, _stream()))
    return ids, count


def rpn_samples(anchors, gt_class_ids, gt_boxes, config, device=None, return_tensors=False):
    """Drop-in for data.rpn_samples (data.py:449-591): anchors [A,4] float64 px, gt_class_ids [G] int32, gt_boxes [G,4]
    int32 px (numpy arrays like the reference's, or CUDA tensors - pass the anchors as a CUDA float64 tensor to avoid
    re-uploading them for every sample) -> (rpn_match int32 [A], rpn_bbox float64 [RPN_TRAIN_ANCHORS_PER_IMAGE, 4]),
    numpy like the reference's unless return_tensors.  The two subsampling draws use np.random.permutation, which is what
    the reference's np.random.choice(ids, extra, replace=False) draws, so a seeded run returns the reference's result."""
    dev = torch.device(device) if device is not None else (anchors.device if isinstance(anchors, torch.Tensor) else torch.device("cuda"))
    if dev.type != "cuda":
        raise TypeError("rpn_samples runs on a CUDA device: maskrcnn_b200 has no CPU path")

    def to_dev(x, dtype):
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        return t.to(device=dev, dtype=dtype).contiguous()
    anc, cls, gtb = to_dev(anchors, torch.float64), to_dev(gt_class_ids, torch.int32), to_dev(gt_boxes, torch.int32)
    A, G = anc.size(0), gtb.size(0)
    T = int(config.RPN_TRAIN_ANCHORS_PER_IMAGE)
    std = np.asarray(config.RPN_BBOX_STD_DEV, dtype=np.float64).reshape(4)
    with torch.cuda.device(dev):
        match = torch.empty(A, dtype=torch.int32, device=dev)
        argmax = torch.empty(A, dtype=torch.int32, device=dev)
        ws_bytes = lib.mrcnn_rpn_match_workspace_bytes(G)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        check(lib.mrcnn_rpn_match(anc.data_ptr(), A, gtb.data_ptr(), cls.data_ptr(), G, match.data_ptr(), argmax.data_ptr(),
                                  ws.data_ptr(), ws_bytes, _stream()))
        # subsample (data.py:538-553); the counts are read back like the reference's np.where does implicitly
        ids, count = _compact_equal(match, 1)
        n_pos = int(count.item())
        extra = n_pos - T // 2
        if extra > 0:
            perm = torch.from_numpy(np.random.permutation(n_pos)[:extra].astype(np.int32)).to(dev)
            check(lib.mrcnn_scatter_fill(match.data_ptr(), ids.data_ptr(), perm.data_ptr(), extra, 0, _stream()))
            n_pos -= extra
        ids, count = _compact_equal(match, -1)
        n_neg = int(count.item())
        extra = n_neg - (T - n_pos)
        if extra > 0:
            perm = torch.from_numpy(np.random.permutation(n_neg)[:extra].astype(np.int32)).to(dev)
            check(lib.mrcnn_scatter_fill(match.data_ptr(), ids.data_ptr(), perm.data_ptr(), extra, 0, _stream()))
        ids, count = _compact_equal(match, 1)
        bbox = torch.empty((T, 4), dtype=torch.float64, device=dev)
        check(lib.mrcnn_rpn_deltas(anc.data_ptr(), gtb.data_ptr(), argmax.data_ptr(), ids.data_ptr(), count.data_ptr(), T,
                                   (ctypes.c_double * 4)(*[float(v) for v in std]), bbox.data_ptr(), _stream()))
    if return_tensors:
        return match, bbox
    return match.cpu().numpy(), bbox.cpu().numpy()


# ------------------------------------------------------------------------------------------------
# mask paste-back
# ------------------------------------------------------------------------------------------------
def full_masks(class_id, boxes, masks, height, width):
    """Drop-in for data.full_masks (data.py:287-314): class_id [D] integer, boxes [D,4] px (y1,x1,y2,x2), masks
    [D,NC,mh,mw] float -> bool [D,height,width], bit-identical to the reference's PIL route (mask * 255 -> 8 bit ->
    Pillow bilinear resize to the box -> paste -> '> 127').  Two launches for all detections and no host round trip (the
    reference does `.item()`, `.tolist()` and a CPU copy per detection); leading batch dimensions are allowed
    (class_id [...,D], boxes [...,D,4], masks [...,D,NC,mh,mw] -> [...,D,height,width]).  Unlike the reference, which
    raises ValueError from PIL, an empty box gives an all-False mask (zero-padded detection rows)."""
    _require_cuda(masks, "masks")
    _require_cuda(boxes, "boxes")
    _require_cuda(class_id, "class_id")
    if masks.dim() < 4 or boxes.shape != masks.shape[:-3] + (4,) or class_id.shape != masks.shape[:-3]:
        raise ValueError("class_id [...,D], boxes [...,D,4], masks [...,D,NC,mh,mw]")
    lead = tuple(masks.shape[:-3])
    NC, mh, mw = (int(v) for v in masks.shape[-3:])
    D = int(class_id.numel())
    H, W = int(height), int(width)
    cls = class_id.reshape(-1).to(torch.int64).contiguous()
    bx = boxes.reshape(-1, 4).float().contiguous()
    mk = masks.detach().reshape(-1, NC, mh, mw).float().contiguous()
    out = torch.empty((D, H, W), dtype=torch.bool, device=masks.device)
    with torch.cuda.device(masks.device):
        ws_bytes = lib.mrcnn_full_masks_workspace_bytes(D, mh, mw, H, W)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=masks.device)
        check(lib.mrcnn_full_masks(cls.data_ptr(), bx.data_ptr(), mk.data_ptr(), D, NC, mh, mw, H, W, out.data_ptr(), ws.data_ptr(),
                                   ws_bytes, _stream()))
    return out.reshape(lead + (H, W))


def decode_masks(masks, scale, cropbox):
    """Drop-in for data.decode_masks (data.py:265-284): masks [D,H,W] torch.bool (what full_masks returns; -> 0 / 255) or
    torch.uint8 (pixel values kept), scale = the factor the frame was resized by, cropbox = the window (an object with
    height() / width() like data.Box, or a (height, width) pair) -> uint8 [D,nh,nw]: every mask centre-cropped to the
    window and resized to round(h / scale) x round(w / scale) with Pillow's 8-bit bilinear resample, bit-identical to the
    reference's PIL / torchvision route, not thresholded.  Two launches for all masks and no host round trip (the
    reference copies every mask to the CPU and back).  scale == 1 returns `masks` itself (data.py:267-268)."""
    if scale == 1:
        return masks
    _require_cuda(masks, "masks")
    if masks.dim() != 3 or masks.dtype not in (torch.bool, torch.uint8):
        raise ValueError("masks must be [D,H,W] torch.bool or torch.uint8")
    if hasattr(cropbox, "height"):
        ch, cw = int(cropbox.height()), int(cropbox.width())
    else:
        ch, cw = int(cropbox[0]), int(cropbox[1])
    D, H, W = (int(v) for v in masks.shape)
    if ch > H or cw > W:
        raise ValueError("decode_masks: the crop window is larger than the mask (CenterCrop would pad)")
    # torchvision center_crop's origin and data.py:276-277's target size, in Python arithmetic like the reference
    top = int(round((H - ch) / 2.0))
    left = int(round((W - cw) / 2.0))
    nh = int(round(ch * 1.0 / scale))
    nw = int(round(cw * 1.0 / scale))
    if nh <= 0 or nw <= 0 or ch <= 0 or cw <= 0:
        raise ValueError("height and width must be > 0")     # PIL's error for the same input
    m = masks.contiguous()
    out = torch.empty((D, nh, nw), dtype=torch.uint8, device=masks.device)
    with torch.cuda.device(masks.device):
        ws_bytes = lib.mrcnn_decode_masks_workspace_bytes(ch, cw, nh, nw)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=masks.device)
        check(lib.mrcnn_decode_masks(m.data_ptr(), 1 if masks.dtype == torch.bool else 0, D, H, W, top, left, ch, cw, nh, nw,
                                     out.data_ptr(), ws.data_ptr(), ws_bytes, _stream()))
    return out
