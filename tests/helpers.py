import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")


def golden():
    return np.load(GOLDEN)


def ulp_diff(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


def rel_err(got, want):
    """max |got-want| / max|want| — the 'relative' of north_star's 1e-5 (scale = output magnitude)."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    scale = max(np.abs(want).max(), 1e-30) if want.size else 1.0
    return float(np.abs(got - want).max() / scale) if want.size else 0.0


GOLDEN_TARGETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_targets_v1.npz")


def golden_targets():
    return np.load(GOLDEN_TARGETS)


def replay_perms(perms):
    """A stand-in for torch.randperm that replays the recorded draws in order (and checks their lengths)."""
    it = iter(perms)

    def randperm(n):
        p = np.asarray(next(it))
        assert len(p) == n, "permutation length differs from the reference's draw"
        return p
    return randperm


def golden_masks(tag):
    """(class ids, boxes, masks [D,1,mh,mw], H, W, expected bool [D,H,W]) of tests/golden/golden_masks_v1.npz; only the
    selected class plane of every detection was stored, so the class ids are all zero here."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_masks_v1.npz"))
    boxes, sel = g[f"{tag}_in_boxes"], g[f"{tag}_in_masks_sel"]
    h, w = (int(v) for v in g[f"{tag}_in_hw"])
    d = len(boxes)
    want = np.unpackbits(g[f"{tag}_out_bits"])[:d * h * w].reshape(d, h, w).astype(bool)
    return np.zeros(d, np.int64), boxes, np.ascontiguousarray(sel[:, None]), h, w, want


def golden_rpnhead():
    """(class-logit conv outputs per level, bbox conv outputs per level, expected logits / class / bbox) of
    tests/golden/golden_rpnhead_v1.npz."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_rpnhead_v1.npz"))
    n = sum(1 for k in g.files if k.startswith("in_logits_"))
    return [g[f"in_logits_{l}"] for l in range(n)], [g[f"in_bbox_{l}"] for l in range(n)], g["out_logits"], g["out_class"], g["out_bbox"]


SOFTMAX_TOL = 1e-6  # absolute, on probabilities in [0, 1]: torch's CPU softmax uses an approximate vectorised exp


def golden_detect():
    """tests/golden/golden_detect_v1.npz: the reference's predict.py flow (BASELINE configs[0]) recorded at the boundary
    of every operator this repo replaces (tests/golden/make_golden_detect.py)."""
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_detect_v1.npz"))


def detect_flow_expected_masks(g):
    d = len(g["mask_valid"])
    h, w = (int(v) for v in g["mask_in_hw"])
    return np.unpackbits(g["mask_out_bits"])[:d * h * w].reshape(d, h, w).astype(bool)


def golden_decode(tag):
    """(bool masks [D,H,W], scale, (window height, window width), expected uint8 [D,nh,nw]) of
    tests/golden/golden_decode_v1.npz (the reference's data.decode_masks, tests/golden/make_golden_decode.py)."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_decode_v1.npz"))
    d, h, w, y1, x1, y2, x2 = (int(v) for v in g[f"{tag}_in_geom"])
    m = np.unpackbits(g[f"{tag}_in_bits"])[:d * h * w].reshape(d, h, w).astype(bool)
    return m, float(g[f"{tag}_in_scale"]), (y2 - y1, x2 - x1), g[f"{tag}_out"]
