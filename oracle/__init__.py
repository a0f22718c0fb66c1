"""TEST INFRASTRUCTURE ONLY — ctypes/numpy front-end of the plain-C oracle (oracle/mrcnn_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (maskrcnn_b200/) never does.

All arrays are numpy, C-contiguous, fp32 / int32 / int64, NCHW — the reference's layout.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "mrcnn_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.orc_nms.restype = ctypes.c_int64
        L.orc_nms.argtypes = [_f32p, ctypes.c_int64, ctypes.c_float, _i64p]
        L.orc_crop_forward.restype = ctypes.c_int
        L.orc_crop_forward.argtypes = [_f32p] + [ctypes.c_int] * 4 + [_f32p, _i32p, ctypes.c_int,
                                                                    ctypes.c_float, ctypes.c_int,
                                                                    ctypes.c_int, _f32p]
        L.orc_crop_backward.restype = ctypes.c_int
        L.orc_crop_backward.argtypes = [_f32p, _f32p, _i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        _f32p] + [ctypes.c_int] * 4
        L.orc_roi_level.restype = ctypes.c_int
        L.orc_roi_level.argtypes = [_f32p, ctypes.c_float]
        pp = ctypes.POINTER(_f32p)
        ip = ctypes.POINTER(ctypes.c_int)
        L.orc_pyramid_roi_align_fwd.restype = ctypes.c_int
        L.orc_pyramid_roi_align_fwd.argtypes = [pp, ip, ip, ctypes.c_int, ctypes.c_int, _f32p, _i32p,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_float, _f32p, _i32p]
        L.orc_pyramid_roi_align_bwd.restype = ctypes.c_int
        L.orc_pyramid_roi_align_bwd.argtypes = [_f32p, ip, ip, ctypes.c_int, ctypes.c_int, _f32p, _i32p,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_float, pp]
        L.orc_proposal_layer.restype = ctypes.c_int64
        L.orc_proposal_layer.argtypes = [_f32p, _f32p, _f32p, ctypes.c_int64, ctypes.c_int64,
                                         ctypes.c_int64, ctypes.c_float, _f32p, ctypes.c_float,
                                         ctypes.c_float, _f32p, _f32p, _i64p]
        _f64p = ctypes.POINTER(ctypes.c_double)
        L.orc_rpn_match.restype = None
        L.orc_rpn_match.argtypes = [_f64p, ctypes.c_int64, _i32p, _i32p, ctypes.c_int64, _i32p, _i32p]
        L.orc_rpn_deltas.restype = None
        L.orc_rpn_deltas.argtypes = [_f64p, _i32p, _i32p, _i64p, ctypes.c_int64, _f64p, _f64p]
        L.orc_target_classify.restype = None
        L.orc_target_classify.argtypes = [_f32p, ctypes.c_int64, _f32p, _i32p, ctypes.c_int64, _i64p, _i64p, _i32p, _f32p, _i64p]
        L.orc_target_emit.restype = None
        L.orc_target_emit.argtypes = [_f32p, _f32p, _i32p, _f32p, ctypes.c_int, ctypes.c_int, _i64p, ctypes.c_int64, _i64p,
                                      ctypes.c_int64, _i32p, _f32p, ctypes.c_int, ctypes.c_int, _f32p, _i32p, _f32p, _f32p]
        L.orc_detection_layer.restype = ctypes.c_int64
        L.orc_detection_layer.argtypes = [_f32p, _f32p, _f32p, ctypes.c_int64, ctypes.c_int64, _f32p,
                                          ctypes.c_float, ctypes.c_float, ctypes.c_int64, _f32p,
                                          ctypes.c_float, ctypes.c_float, _f32p, _i64p]
        L.orc_full_masks.restype = ctypes.c_int
        L.orc_full_masks.argtypes = [_i64p, _f32p, _f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, ctypes.POINTER(ctypes.c_uint8)]
        L.orc_decode_masks.restype = ctypes.c_int
        L.orc_decode_masks.argtypes = [ctypes.POINTER(ctypes.c_uint8), ctypes.c_int64] + [ctypes.c_int] * 8 + [ctypes.POINTER(ctypes.c_uint8)]
        L.orc_rpn_pack.restype = None
        L.orc_rpn_pack.argtypes = [pp, pp, ip, ip, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, _f32p, _f32p]
        L.orc_boxes_refine.restype = None
        L.orc_boxes_refine.argtypes = [_f32p, _f32p, ctypes.c_int64, _f32p]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=_f32p):
    return a.ctypes.data_as(t)


class OracleError(RuntimeError):
    pass


def nms(dets, threshold):
    """c++ext/maskrcnn/__init__.py:21-22 -> nms_cpu.cpp:11-70.  dets [N,5] -> int64 [K] ascending."""
    dets = _f32(dets).reshape(-1, 5)
    n = dets.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    k = lib().orc_nms(_p(dets), n, float(threshold), _p(keep, _i64p))
    return keep[:k].copy()


def crop_forward(image, boxes, box_index, crop_h, crop_w, extrapolation_value=0.0):
    """crop_cpu.cpp:119-164.  image [B,C,H,W] -> crops [N,C,ch,cw]."""
    image = _f32(image)
    boxes = _f32(boxes).reshape(-1, 4)
    box_index = np.ascontiguousarray(box_index, dtype=np.int32)
    B, C, H, W = image.shape
    N = boxes.shape[0]
    out = np.zeros((N, C, crop_h, crop_w), dtype=np.float32)
    rc = lib().orc_crop_forward(_p(image), B, C, H, W, _p(boxes), _p(box_index, _i32p), N,
                                float(extrapolation_value), crop_h, crop_w, _p(out))
    if rc:
        raise OracleError("box_index out of range")
    return out


def crop_backward(grads, boxes, box_index, image_shape):
    """crop_cpu.cpp:167-265.  grads [N,C,ch,cw] -> grads_image [B,C,H,W]."""
    grads = _f32(grads)
    boxes = _f32(boxes).reshape(-1, 4)
    box_index = np.ascontiguousarray(box_index, dtype=np.int32)
    B, C, H, W = image_shape
    N, _, ch, cw = grads.shape
    gi = np.empty((B, C, H, W), dtype=np.float32)
    rc = lib().orc_crop_backward(_p(grads), _p(boxes), _p(box_index, _i32p), N, ch, cw, _p(gi), B, C, H, W)
    if rc:
        raise OracleError("box_index out of range")
    return gi


def roi_levels(boxes, image_area):
    boxes = _f32(boxes).reshape(-1, 4)
    return np.array([lib().orc_roi_level(_p(boxes[i:i + 1]), float(image_area)) for i in range(len(boxes))],
                    dtype=np.int32)


def _pyr_args(fms):
    fms = [_f32(f) for f in fms]
    assert len(fms) == 4
    B, C = fms[0].shape[:2]
    H = (ctypes.c_int * 4)(*[f.shape[2] for f in fms])
    W = (ctypes.c_int * 4)(*[f.shape[3] for f in fms])
    ptrs = (_f32p * 4)(*[_p(f) for f in fms])
    return fms, B, C, H, W, ptrs


def pyramid_roi_align_fwd(fms, boxes, box_ind, pool, image_area):
    """model.py:276-393 (batched over box_ind).  fms: 4 arrays [B,C,Hl,Wl] -> ([N,C,p,p], levels[N])."""
    fms, B, C, H, W, ptrs = _pyr_args(fms)
    boxes = _f32(boxes).reshape(-1, 4)
    N = boxes.shape[0]
    box_ind = np.zeros(N, np.int32) if box_ind is None else np.ascontiguousarray(box_ind, dtype=np.int32)
    out = np.zeros((N, C, pool, pool), dtype=np.float32)
    lv = np.zeros(N, dtype=np.int32)
    rc = lib().orc_pyramid_roi_align_fwd(ptrs, H, W, B, C, _p(boxes), _p(box_ind, _i32p), N, pool,
                                         float(image_area), _p(out), _p(lv, _i32p))
    if rc:
        raise OracleError("box_index out of range")
    return out, lv


def pyramid_roi_align_bwd(grads, shapes, boxes, box_ind, image_area):
    """Adjoint of pyramid_roi_align_fwd.  shapes: 4 tuples (B,C,Hl,Wl) -> list of 4 grad arrays."""
    grads = _f32(grads)
    boxes = _f32(boxes).reshape(-1, 4)
    N, C, pool, _ = grads.shape
    B = shapes[0][0]
    box_ind = np.zeros(N, np.int32) if box_ind is None else np.ascontiguousarray(box_ind, dtype=np.int32)
    gf = [np.empty(s, dtype=np.float32) for s in shapes]
    H = (ctypes.c_int * 4)(*[s[2] for s in shapes])
    W = (ctypes.c_int * 4)(*[s[3] for s in shapes])
    ptrs = (_f32p * 4)(*[_p(g) for g in gf])
    rc = lib().orc_pyramid_roi_align_bwd(_p(grads), H, W, B, C, _p(boxes), _p(box_ind, _i32p), N, pool,
                                         float(image_area), ptrs)
    if rc:
        raise OracleError("box_index out of range")
    return gf


def proposal_layer(rpn_class, rpn_bbox, anchors, pre_nms_limit, post_nms_limit, nms_threshold,
                   std=(0.1, 0.1, 0.2, 0.2), height=1024.0, width=1024.0, return_intermediate=False):
    """model.py:1307-1382, one image.  rpn_class [A,2], rpn_bbox [A,4], anchors [A,4] -> rois [K,4]."""
    rpn_class = _f32(rpn_class).reshape(-1, 2)
    rpn_bbox = _f32(rpn_bbox).reshape(-1, 4)
    anchors = _f32(anchors).reshape(-1, 4)
    A = rpn_class.shape[0]
    pre = min(int(pre_nms_limit), A)
    rois = np.zeros((max(int(post_nms_limit), 1), 4), dtype=np.float32)
    dets = np.zeros((max(pre, 1), 5), dtype=np.float32)
    order = np.zeros(max(pre, 1), dtype=np.int64)
    stdv = _f32(std)
    k = lib().orc_proposal_layer(_p(rpn_class), _p(rpn_bbox), _p(anchors), A, pre, int(post_nms_limit),
                                 float(nms_threshold), _p(stdv), float(height), float(width), _p(rois),
                                 _p(dets), _p(order, _i64p))
    if return_intermediate:
        return rois[:k].copy(), dets[:pre].copy(), order[:pre].copy()
    return rois[:k].copy()


def detection_layer(rois, probs, deltas, window, min_conf, nms_threshold, max_inst,
                    std=(0.1, 0.1, 0.2, 0.2), height=1024.0, width=1024.0, return_index=False):
    """model.py:1389-1487, one image -> [D,6] (y1,x1,y2,x2,score,class), score-descending."""
    rois = _f32(rois).reshape(-1, 4)
    probs = _f32(probs)
    N, NC = probs.shape
    deltas = _f32(deltas).reshape(N, NC, 4)
    window = _f32(window)
    out = np.zeros((max(int(max_inst), 1), 6), dtype=np.float32)
    idx = np.zeros(max(int(max_inst), 1), dtype=np.int64)
    stdv = _f32(std)
    d = lib().orc_detection_layer(_p(rois), _p(probs), _p(deltas), N, NC, _p(window), float(min_conf or 0.0),
                                  float(nms_threshold), int(max_inst), _p(stdv), float(height), float(width),
                                  _p(out), _p(idx, _i64p))
    if return_index:
        return out[:d].copy(), idx[:d].copy()
    return out[:d].copy()


def boxes_refine(boxes, deltas):
    """data.py:124-148."""
    boxes = _f32(boxes).reshape(-1, 4)
    deltas = _f32(deltas).reshape(-1, 4)
    out = np.empty_like(boxes)
    lib().orc_boxes_refine(_p(boxes), _p(deltas), boxes.shape[0], _p(out))
    return out


def target_classify(rois, gt_class_ids, gt_boxes):
    """model.py:431-463, :513-516 -> (positive RoI indices, negative RoI indices, assigned gt row per RoI, max IoU per RoI)."""
    rois, gt_boxes = _f32(rois), _f32(gt_boxes)
    cls = np.ascontiguousarray(gt_class_ids, np.int32)
    n, g = len(rois), len(gt_boxes)
    pos, neg = np.zeros(max(n, 1), np.int64), np.zeros(max(n, 1), np.int64)
    assign, iou = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float32)
    counts = np.zeros(2, np.int64)
    lib().orc_target_classify(_p(rois), n, _p(gt_boxes), _p(cls, _i32p), g, _p(pos, _i64p), _p(neg, _i64p), _p(assign, _i32p),
                              _p(iou), _p(counts, _i64p))
    return pos[:counts[0]].copy(), neg[:counts[1]].copy(), assign[:n], iou[:n]


def target_emit(rois, gt_class_ids, gt_boxes, gt_masks, sel_pos, sel_neg, assign, std_dev, mask_shape):
    """model.py:474-541 for already chosen positives / negatives -> (rois, class ids, deltas, masks)."""
    rois, gt_boxes, gt_masks = _f32(rois), _f32(gt_boxes), _f32(gt_masks)
    cls = np.ascontiguousarray(gt_class_ids, np.int32)
    sp, sn = np.ascontiguousarray(sel_pos, np.int64), np.ascontiguousarray(sel_neg, np.int64)
    asg = np.ascontiguousarray(assign, np.int32)
    std = _f32(std_dev)
    mh, mw = int(mask_shape[0]), int(mask_shape[1])
    t = len(sp) + len(sn)
    o_rois, o_cls = np.zeros((t, 4), np.float32), np.zeros(t, np.int32)
    o_d, o_m = np.zeros((t, 4), np.float32), np.zeros((t, mh, mw), np.float32)
    h, w = gt_masks.shape[-2:]
    lib().orc_target_emit(_p(rois), _p(gt_boxes), _p(cls, _i32p), _p(gt_masks), h, w, _p(sp, _i64p), len(sp), _p(sn, _i64p), len(sn),
                          _p(asg, _i32p), _p(std), mh, mw, _p(o_rois), _p(o_cls, _i32p), _p(o_d), _p(o_m))
    return o_rois, o_cls, o_d, o_m


def target_counts(n_pos, n_neg, train_rois_per_image, roi_positive_ratio):
    """How many positives / negatives mrn_samples keeps (model.py:466-471, :518-522), in Python doubles like the reference."""
    pc = min(int(train_rois_per_image * roi_positive_ratio), n_pos)
    if pc <= 0:
        return 0, 0                       # model.py:517: negatives only when positive_count > 0
    r = 1.0 / roi_positive_ratio
    nc = min(int(r * pc - pc), n_neg)
    return pc, max(nc, 0)


def mrn_samples(rois, gt_class_ids, gt_boxes, gt_masks, train_rois_per_image, roi_positive_ratio, std_dev, mask_shape, randperm):
    """model.py:396-576 for one image.  `randperm(n)` stands for the reference's torch.randperm draws, called in the
    reference's order: positives first (:468), then negatives (:520)."""
    pos, neg, assign, _ = target_classify(rois, gt_class_ids, gt_boxes)
    pc, _ = target_counts(len(pos), len(neg), train_rois_per_image, roi_positive_ratio)
    sel_pos = np.zeros(0, np.int64)
    if len(pos) > 0:
        sel_pos = pos[np.asarray(randperm(len(pos)), np.int64)[:int(train_rois_per_image * roi_positive_ratio)]]
    sel_neg = np.zeros(0, np.int64)
    if len(neg) > 0 and len(sel_pos) > 0:
        r = 1.0 / roi_positive_ratio
        sel_neg = neg[np.asarray(randperm(len(neg)), np.int64)[:int(r * len(sel_pos) - len(sel_pos))]]
    return target_emit(rois, gt_class_ids, gt_boxes, gt_masks, sel_pos, sel_neg, assign, std_dev, mask_shape)


def rpn_samples(anchors, gt_class_ids, gt_boxes, train_anchors_per_image, std_dev, permutation):
    """data.py:449-591.  anchors float64 [A,4] px, gt_boxes int32 [G,4] px.  `permutation(n)` stands for the draws inside
    the reference's two np.random.choice(ids, extra, replace=False) calls (= ids[permutation(len(ids))[:extra]]), in the
    reference's order: positives (:543) then negatives (:552).  Returns (rpn_match int32 [A], rpn_bbox float64 [T,4])."""
    f64p = ctypes.POINTER(ctypes.c_double)
    anchors = np.ascontiguousarray(anchors, np.float64)
    gtb = np.ascontiguousarray(gt_boxes, np.int32)
    cls = np.ascontiguousarray(gt_class_ids, np.int32)
    A, G = len(anchors), len(gtb)
    match, argmax = np.zeros(A, np.int32), np.zeros(A, np.int32)
    lib().orc_rpn_match(anchors.ctypes.data_as(f64p), A, _p(gtb, _i32p), _p(cls, _i32p), G, _p(match, _i32p), _p(argmax, _i32p))
    T = int(train_anchors_per_image)
    ids = np.where(match == 1)[0]
    extra = len(ids) - T // 2
    if extra > 0:
        match[ids[np.asarray(permutation(len(ids)))[:extra]]] = 0
    ids = np.where(match == -1)[0]
    extra = len(ids) - (T - int(np.sum(match == 1)))
    if extra > 0:
        match[ids[np.asarray(permutation(len(ids)))[:extra]]] = 0
    ids = np.ascontiguousarray(np.where(match == 1)[0], np.int64)
    bbox = np.zeros((T, 4), np.float64)
    std = np.ascontiguousarray(std_dev, np.float64)
    lib().orc_rpn_deltas(anchors.ctypes.data_as(f64p), _p(gtb, _i32p), _p(argmax, _i32p), _p(ids, _i64p), len(ids),
                         std.ctypes.data_as(f64p), bbox.ctypes.data_as(f64p))
    return match, bbox


def full_masks(class_ids, boxes, masks, height, width):
    """data.full_masks (data.py:287-314): class_ids [D], boxes [D,4] px, masks [D,NC,mh,mw] -> bool [D,height,width].
    Raises ValueError where PIL does (empty box); boxes that leave the image are pasted clipped, as the reference's Pad does."""
    cls = np.ascontiguousarray(class_ids, np.int64)
    boxes, masks = _f32(boxes), _f32(masks)
    d, nc, mh, mw = masks.shape
    out = np.zeros((d, int(height), int(width)), np.uint8)
    rc = lib().orc_full_masks(_p(cls, _i64p), _p(boxes), _p(masks), d, nc, mh, mw, int(height), int(width),
                              out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    if rc != 0:
        raise ValueError("full_masks: detection %d has an empty box" % (-rc - 1))
    return out.astype(bool)


def decode_geometry(height, width, scale, crop_hw):
    """The integer geometry of data.decode_masks (data.py:265-284), computed with Python's own arithmetic as the reference
    does: torchvision CenterCrop's origin int(round((H - ch) / 2.0)) (transforms/functional.py center_crop; Python's
    round is half-to-even) and the target size round(ch * 1.0 / scale), round(cw * 1.0 / scale) (data.py:276-277).
    Returns (top, left, ch, cw, nh, nw)."""
    ch, cw = int(crop_hw[0]), int(crop_hw[1])
    if ch > int(height) or cw > int(width):
        raise ValueError("decode_masks: the crop window is larger than the mask (CenterCrop would pad)")
    top = int(round((int(height) - ch) / 2.0))
    left = int(round((int(width) - cw) / 2.0))
    nh = round(ch * 1.0 / scale)
    nw = round(cw * 1.0 / scale)
    return top, left, ch, cw, int(nh), int(nw)


def decode_masks(masks, scale, crop_hw):
    """data.decode_masks (data.py:265-284): masks [D,H,W] bool (-> 0 / 255, PIL mode '1' -> 'L') or uint8 (kept as is),
    scale = the resize factor encode_image applied, crop_hw = (window height, window width) -> uint8 [D,nh,nw] (8-bit
    bilinear, not thresholded).  scale == 1 returns the input, as the reference does (data.py:267-268)."""
    if scale == 1:
        return masks
    m = np.ascontiguousarray(masks)
    if m.dtype == np.bool_:
        m = m.astype(np.uint8) * np.uint8(255)
    elif m.dtype != np.uint8:
        raise TypeError("decode_masks: masks must be bool or uint8")
    d, H, W = m.shape
    top, left, ch, cw, nh, nw = decode_geometry(H, W, scale, crop_hw)
    if nh <= 0 or nw <= 0:
        raise ValueError("height and width must be > 0")   # PIL's message
    out = np.zeros((d, nh, nw), np.uint8)
    u8p = ctypes.POINTER(ctypes.c_uint8)
    rc = lib().orc_decode_masks(m.ctypes.data_as(u8p), d, H, W, top, left, ch, cw, nh, nw, out.ctypes.data_as(u8p))
    if rc != 0:
        raise ValueError("decode_masks: bad geometry")
    return out


def rpn_pack(class_logits, bboxes):
    """model.py:624-641 + :1294-1304: per-level conv outputs class_logits[l] [B,2K,H,W], bboxes[l] [B,4K,H,W] (NCHW) ->
    (rpn_class_logits [B,A,2], rpn_class [B,A,2], rpn_bbox [B,A,4])."""
    ls, bs = [_f32(t) for t in class_logits], [_f32(t) for t in bboxes]
    n = len(ls)
    B, K = ls[0].shape[0], ls[0].shape[1] // 2
    Hs = (ctypes.c_int * n)(*[t.shape[2] for t in ls])
    Ws = (ctypes.c_int * n)(*[t.shape[3] for t in ls])
    A = K * sum(t.shape[2] * t.shape[3] for t in ls)
    o_l, o_c, o_b = np.zeros((B, A, 2), np.float32), np.zeros((B, A, 2), np.float32), np.zeros((B, A, 4), np.float32)
    lp = (_f32p * n)(*[_p(t) for t in ls])
    bp = (_f32p * n)(*[_p(t) for t in bs])
    lib().orc_rpn_pack(lp, bp, Hs, Ws, n, B, K, _p(o_l), _p(o_c), _p(o_b))
    return o_l, o_c, o_b
