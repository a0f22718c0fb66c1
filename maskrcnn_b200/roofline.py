"""Algorithmic-byte accounting for the RoI hot path.  Host-side numpy."""
import numpy as np

_T3 = _T4 = _T5 = None


def _level_of_q(q):
    v = np.float32(4.0) + np.float32(np.log2(np.float64(q)))
    if not np.isfinite(v):
        return 2
    return int(min(5, max(2, np.rint(v))))


def roi_levels(boxes, image_area):
    """model.py:323-338 in fp32 (correctly rounded log2), vectorised."""
    b = np.asarray(boxes, np.float32)
    h = b[:, 2] - b[:, 0]
    w = b[:, 3] - b[:, 1]
    denom = np.float32(224.0) / np.sqrt(np.float32(image_area))
    with np.errstate(all="ignore"):
        q = (np.sqrt(h * w) / denom).astype(np.float32)
        v = (np.float32(4.0) + np.log2(q.astype(np.float64)).astype(np.float32)).astype(np.float32)
    lv = np.where(np.isfinite(v), np.rint(v), 2).astype(np.int64)
    return np.clip(lv, 2, 5).astype(np.int32)


def _axis_taps(a1, a2, size, crop):
    """(lo, hi) int arrays [crop] of the taps touched along one axis; -1 where the sample is outside."""
    sm1 = np.float32(size - 1)
    i = np.arange(crop, dtype=np.float32)
    if crop > 1:
        scale = ((a2 - a1) * sm1) / np.float32(crop - 1)
        pos = (a1 * sm1 + i * scale).astype(np.float32)
    else:
        pos = np.array([0.5 * float(a1 + a2) * float(size - 1)], np.float32)
    ok = (pos >= 0) & (pos <= sm1)
    lo = np.where(ok, np.floor(pos), -1).astype(np.int64)
    hi = np.where(ok, np.ceil(pos), -1).astype(np.int64)
    return lo, hi


def unique_taps(boxes, box_ind, pool, image_hw, level_hw, batch):
    """U = number of distinct (image, level, y, x) feature-map positions read by PyramidROIAlign.  `pool` may be a tuple of
    pool sizes: the positions read by ANY of those heads, each counted once (a fused multi-head forward)."""
    pools = tuple(pool) if isinstance(pool, (tuple, list)) else (pool,)
    boxes = np.asarray(boxes, np.float32)
    lv = roi_levels(boxes, float(image_hw[0] * image_hw[1]))
    ind = np.zeros(len(boxes), np.int64) if box_ind is None else np.asarray(box_ind, np.int64)
    masks = [np.zeros((batch, h, w), dtype=bool) for (h, w) in level_hw]
    for n in range(len(boxes)):
        l = lv[n] - 2
        H, W = level_hw[l]
        for pl in pools:
            ylo, yhi = _axis_taps(boxes[n, 0], boxes[n, 2], H, pl)
            xlo, xhi = _axis_taps(boxes[n, 1], boxes[n, 3], W, pl)
            ys = np.unique(np.concatenate([ylo, yhi]))
            xs = np.unique(np.concatenate([xlo, xhi]))
            ys, xs = ys[ys >= 0], xs[xs >= 0]
            if len(ys) and len(xs):
                masks[l][ind[n]][np.ix_(ys, xs)] = True
    return int(sum(m.sum() for m in masks)), lv


def roialign_fwd_bytes(n, c, pool, unique):
    """output written once + unique taps read once + boxes/index."""
    return n * c * pool * pool * 4 + 4 * c * unique + n * 20


def roialign_bwd_bytes(n, c, pool, pyramid_elems):
    """grad_out read once + the whole gradient pyramid written once (zero fill included) + boxes/index."""
    return n * c * pool * pool * 4 + 4 * pyramid_elems + n * 20


def proposal_bytes(a, pre, post):
    return a * 8 + pre * 32 + post * 16


def detection_bytes(n, nc, max_inst):
    return n * nc * 4 + n * 16 + n * 16 + max_inst * 24
