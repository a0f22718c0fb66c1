"""Seeded synthetic inputs for the RoI hot path and the anchor generator.

Host-side numpy only; shared by tests/ and bench.py so that the oracle, the golden generator and the
CUDA path all see byte-identical inputs.
"""
import math

import numpy as np

RPN_ANCHOR_SCALES = (32, 64, 128, 256, 512)   # config.py:60
RPN_ANCHOR_RATIOS = (0.5, 1, 2)               # config.py:64
BACKBONE_STRIDES = (4, 8, 16, 32, 64)         # config.py:54
RPN_BBOX_STD_DEV = (0.1, 0.1, 0.2, 0.2)       # config.py:114


def level_anchors(scale, ratios, shape, feature_stride, anchor_stride=1):
    """Anchors of one pyramid level, float64 [h*w*len(ratios), (y1,x1,y2,x2)] in pixels.

    Same enumeration order as utils.py:116-220 (position-major, ratio-minor)."""
    ratios = np.asarray(ratios, dtype=np.float64)
    heights = scale / np.sqrt(ratios)
    widths = scale * np.sqrt(ratios)
    ys = np.arange(0, shape[0], anchor_stride, dtype=np.float64) * feature_stride
    xs = np.arange(0, shape[1], anchor_stride, dtype=np.float64) * feature_stride
    cy = np.repeat(ys, len(xs))[:, None].repeat(len(ratios), 1)          # [P,R]
    cx = np.tile(xs, len(ys))[:, None].repeat(len(ratios), 1)
    hh = np.broadcast_to(heights, cy.shape)
    ww = np.broadcast_to(widths, cx.shape)
    out = np.stack([cy - 0.5 * hh, cx - 0.5 * ww, cy + 0.5 * hh, cx + 0.5 * ww], axis=2)
    return out.reshape(-1, 4)


def pyramid_anchors(image_hw=(1024, 1024), scales=RPN_ANCHOR_SCALES, ratios=RPN_ANCHOR_RATIOS,
                    strides=BACKBONE_STRIDES, anchor_stride=1):
    """All anchors of the pyramid as float32 [A,4] (utils.py:223-291; model.py:991-995 `.float()`).
    1024x1024 -> A = 261,888."""
    out = []
    for s, st in zip(scales, strides):
        shape = (int(math.ceil(image_hw[0] / st)), int(math.ceil(image_hw[1] / st)))
        out.append(level_anchors(s, ratios, shape, st, anchor_stride))
    return np.concatenate(out, axis=0).astype(np.float32)


def random_rois(n, seed, image=1024.0, min_size=16.0, max_size=724.0):
    """[n,4] normalised (y1,x1,y2,x2): sqrt(h*w) log-uniform in [min,max] px, aspect log-uniform in
    [0.5,2], box inside the image — populates all four pyramid levels."""
    rng = np.random.default_rng(seed)
    size = np.exp(rng.uniform(np.log(min_size), np.log(max_size), n))
    aspect = np.exp(rng.uniform(np.log(0.5), np.log(2.0), n))
    h = np.minimum(size * np.sqrt(aspect), image)
    w = np.minimum(size / np.sqrt(aspect), image)
    y1 = rng.uniform(0, 1, n) * (image - h)
    x1 = rng.uniform(0, 1, n) * (image - w)
    b = np.stack([y1, x1, y1 + h, x1 + w], axis=1) / image
    return b.astype(np.float32)


def unique_scores(n, seed, lo=0.0, hi=1.0):
    """A random permutation of n distinct fp32 values in (lo,hi) — no sort ties."""
    rng = np.random.default_rng(seed)
    v = np.linspace(lo, hi, n + 2, dtype=np.float64)[1:-1].astype(np.float32)
    assert len(np.unique(v)) == n, "fp32 cannot hold that many distinct values in the range"
    return v[rng.permutation(n)]


def rpn_outputs(anchors, seed, image=1024.0, n_clusters=40, delta_sigma=0.5, converge=0.0):
    """(rpn_class [A,2], rpn_bbox [A,4]) for one image.  `converge` in (0, 1]: anchors close to an 'object' (and within a
    factor two of its size) regress towards the object's box, as a trained RPN's do, so that the top of the ranking is
    full of near-duplicates and NMS suppresses most of it from the first boxes on (0 = independent random deltas).

    Foreground scores are unique; the top of the ranking is concentrated on anchors close to a few
    'object' centres so that, as in a real RPN, most of the top-k proposals overlap and are
    suppressed (uniformly random high scorers would suppress only ~24 %)."""
    rng = np.random.default_rng(seed)
    A = anchors.shape[0]
    cy = 0.5 * (anchors[:, 0] + anchors[:, 2])
    cx = 0.5 * (anchors[:, 1] + anchors[:, 3])
    centres = rng.uniform(0.1 * image, 0.9 * image, (n_clusters, 2))
    sig = rng.uniform(24.0, 96.0, n_clusters)
    affinity = np.zeros(A, dtype=np.float64)
    for (oy, ox), s in zip(centres, sig):
        affinity = np.maximum(affinity, np.exp(-((cy - oy) ** 2 + (cx - ox) ** 2) / (2 * s * s)))
    noise = rng.uniform(0, 1, A)
    rpn_bbox = (rng.standard_normal((A, 4)) * delta_sigma).astype(np.float32)
    hit = np.zeros(A, dtype=bool)
    if converge > 0.0:
        ah, aw = anchors[:, 2] - anchors[:, 0], anchors[:, 3] - anchors[:, 1]
        best = np.zeros(A, dtype=np.float64)
        target = np.zeros((A, 4), dtype=np.float64)
        for (oy, ox), s in zip(centres, sig):
            aff = np.exp(-((cy - oy) ** 2 + (cx - ox) ** 2) / (2 * s * s))
            side = 3.0 * s
            near = (aff > best) & (np.abs(np.log(side / ah)) < 0.7) & (np.abs(np.log(side / aw)) < 0.7)
            t = np.stack([(oy - cy) / ah, (ox - cx) / aw, np.log(side / ah), np.log(side / aw)], 1) / np.asarray(RPN_BBOX_STD_DEV)
            target[near] = t[near]
            best[near] = aff[near]
        hit = best > 0.3
        mix = (converge * target + (1.0 - converge) * rpn_bbox + rng.standard_normal((A, 4)) * 0.05).astype(np.float32)
        rpn_bbox[hit] = mix[hit]
    # a trained RPN scores the anchors that match an object in position AND size highest
    rank_key = affinity + 0.35 * noise + (0.6 * hit if converge > 0.0 else 0.0)
    order = np.argsort(-rank_key, kind="stable")
    vals = np.linspace(1.0, 0.0, A + 2, dtype=np.float64)[1:-1].astype(np.float32)
    assert len(np.unique(vals)) == A
    fg = np.empty(A, dtype=np.float32)
    fg[order] = vals
    rpn_class = np.stack([1.0 - fg, fg], axis=1).astype(np.float32)
    return rpn_class, rpn_bbox


def head_outputs(n, num_classes, seed, logit_scale=3.0, delta_sigma=0.1):
    """(probs [n,NC] with unique entries, deltas [n,NC,4]) for the detection layer."""
    rng = np.random.default_rng(seed)
    logits = rng.standard_normal((n, num_classes)) * logit_scale
    e = np.exp(logits - logits.max(1, keepdims=True))
    probs = (e / e.sum(1, keepdims=True)).astype(np.float32)
    # make the per-RoI maxima distinct across RoIs so that every score sort is tie-free
    top = probs.max(1)
    if len(np.unique(top)) != n:
        for i in range(n):
            c = probs[i].argmax()
            probs[i, c] = np.nextafter(probs[i, c], np.float32(2.0)) if i % 2 else probs[i, c]
        top = probs.max(1)
        u, cnt = np.unique(top, return_counts=True)
        for v in u[cnt > 1]:
            idx = np.where(top == v)[0]
            for k, i in enumerate(idx[1:], 1):
                c = probs[i].argmax()
                x = probs[i, c]
                for _ in range(k):
                    x = np.nextafter(x, np.float32(2.0))
                probs[i, c] = x
    deltas = (rng.standard_normal((n, num_classes, 4)) * delta_sigma).astype(np.float32)
    return probs, deltas


def pyramid_shapes(batch, channels, image=1024, strides=(4, 8, 16, 32)):
    return [(batch, channels, image // s, image // s) for s in strides]


def feature_pyramid(batch, channels, seed, image=1024, strides=(4, 8, 16, 32)):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal(s, dtype=np.float32) for s in pyramid_shapes(batch, channels, image, strides)]


def target_inputs(n_rois, n_gt, seed, image=1024, n_crowd=0, n_pad=0, positive_fraction=0.3):
    """Inputs of the detection-target layer (model.py:396 mrn_samples) for one image: proposals (normalised, a
    fraction of them jittered copies of gt boxes so that IoU >= 0.5 occurs), gt class ids (int32; `n_crowd` rows
    negative = COCO crowds, `n_pad` trailing rows 0 = padding with zero boxes), gt boxes (normalised) and binary gt
    masks [n_gt, image, image] (an ellipse inside each box)."""
    rng = np.random.default_rng(seed)
    gt = random_rois(n_gt, seed + 1, image=float(image), min_size=image / 16.0, max_size=image * 0.6)
    cls = rng.integers(1, 81, n_gt).astype(np.int32)
    if n_crowd:
        cls[rng.choice(n_gt - n_pad, n_crowd, replace=False)] = -1
    if n_pad:
        cls[n_gt - n_pad:] = 0
        gt[n_gt - n_pad:] = 0.0
    rois = random_rois(n_rois, seed + 2, image=float(image), min_size=image / 64.0, max_size=image * 0.7)
    n_pos = int(n_rois * positive_fraction)
    real = max(n_gt - n_pad, 1)
    src = rng.integers(0, real, n_pos)
    hw = np.stack([gt[src, 2] - gt[src, 0], gt[src, 3] - gt[src, 1]] * 2, 1)
    jitter = (rng.normal(0.0, 0.07, (n_pos, 4)) * hw).astype(np.float32)
    where = rng.permutation(n_rois)[:n_pos]
    rois[where] = np.clip(gt[src] + jitter, 0.0, 1.0)
    masks = np.zeros((n_gt, image, image), np.float32)
    yy, xx = np.mgrid[0:image, 0:image].astype(np.float32)
    for g in range(n_gt - n_pad):
        y1, x1, y2, x2 = gt[g] * image
        cy, cx, ry, rx = (y1 + y2) / 2, (x1 + x2) / 2, max((y2 - y1) / 2, 1.0), max((x2 - x1) / 2, 1.0)
        masks[g] = (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0).astype(np.float32)
    return rois.astype(np.float32), cls, gt.astype(np.float32), masks


def rpn_target_inputs(n_gt, seed, image=1024, n_crowd=0):
    """gt boxes in pixels (int32, as the dataset yields them) + class ids for the RPN anchor matching (data.py:449)."""
    rng = np.random.default_rng(seed)
    gt = np.round(random_rois(n_gt, seed + 1, image=float(image), min_size=image / 32.0, max_size=image * 0.7) * image).astype(np.int32)
    gt[:, 2:] = np.maximum(gt[:, 2:], gt[:, :2] + 2)
    cls = rng.integers(1, 81, n_gt).astype(np.int32)
    if n_crowd:
        cls[rng.choice(n_gt, n_crowd, replace=False)] = -1
    return cls, gt


def mask_head_outputs(n_det, num_classes, seed, image=1024, mask=28, min_size=None, max_size=None, n_pad=0):
    """Inputs of the mask paste-back (data.full_masks, data.py:287): class ids int64 [D] in [1, NC), boxes [D,4] =
    rounded pixel boxes inside the image (what the detection layer emits, model.py:1432) and masks [D,NC,mask,mask] =
    a sigmoid of a smooth random field per (detection, class), values on both sides of 0.5.  `n_pad` trailing rows are
    zero padding (class 0, zero box)."""
    rng = np.random.default_rng(seed)
    lo = image / 64.0 if min_size is None else min_size
    hi = image * 0.8 if max_size is None else max_size
    boxes = np.round(random_rois(n_det, seed + 1, image=float(image), min_size=lo, max_size=hi) * image).astype(np.float32)
    boxes[:, 2:] = np.minimum(np.maximum(boxes[:, 2:], boxes[:, :2] + 1), image)
    cls = rng.integers(1, num_classes, n_det).astype(np.int64)
    yy, xx = np.mgrid[0:mask, 0:mask].astype(np.float32) / mask
    f = rng.uniform(0.5, 3.0, (n_det, num_classes, 2, 1, 1)).astype(np.float32)
    ph = rng.uniform(0, 6.28, (n_det, num_classes, 2, 1, 1)).astype(np.float32)
    field = 2.5 * np.sin(6.28 * f[:, :, 0] * yy + ph[:, :, 0]) * np.cos(6.28 * f[:, :, 1] * xx + ph[:, :, 1])
    field = field + rng.standard_normal(field.shape).astype(np.float32) * 0.5
    masks = (1.0 / (1.0 + np.exp(-field))).astype(np.float32)
    if n_pad:
        cls[n_det - n_pad:] = 0
        boxes[n_det - n_pad:] = 0.0
    return cls, boxes, masks
