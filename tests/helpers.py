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


def check_detect_flow_with_oracle(g):
    """Replays a recorded predict.py flow (tests/golden/make_golden_detect.py) stage by stage with the oracle on the
    reference's own tensors; returns the number of detections whose masks the reference produced."""
    import hashlib
    import pytest
    import oracle
    from maskrcnn_b200 import synth
    size = int(g["image_dim"])
    keys = g.files if hasattr(g, "files") else list(g)
    nl = sum(1 for k in keys if k.startswith("rpn_in_logits_"))
    # rpn_detect (model.py:1294-1304)
    logits, cls, bbox = oracle.rpn_pack([g[f"rpn_in_logits_{l}"] for l in range(nl)], [g[f"rpn_in_bbox_{l}"] for l in range(nl)])
    assert hashlib.sha256(logits.tobytes() + bbox.tobytes()).digest() == g["rpn_out_logits_bbox_sha256"].tobytes()
    assert np.abs(cls - g["rpn_out_class"]).max() <= SOFTMAX_TOL
    # rpn_refine (model.py:1307-1382).  The flow's own rpn_class has thousands of anchors tied at fg = 1.0 (saturated
    # random-init softmax), where the reference's result is decided by torch's unstable sort; the recording holds a second
    # reference run with the ties broken in the order that sort chose (make_golden_detect.py)
    np.testing.assert_array_equal(synth.pyramid_anchors((size, size)), g["prop_in_anchors"])
    pre, post = (int(v) for v in g["prop_in_limits"])
    fg = g["prop_in_fg_detied"]
    rois = oracle.proposal_layer(np.stack([1.0 - fg, fg], 1).astype(np.float32), bbox[0], g["prop_in_anchors"], pre, post,
                                 float(g["prop_in_thr"]), height=float(size), width=float(size))
    assert rois.shape == g["prop_out_rois_detied"][0].shape
    assert ulp_diff(rois, g["prop_out_rois_detied"][0]).max() <= 4     # torch.exp's rounding
    # roi_align 7x7 and 14x14 (model.py:276-393)
    fms = [g[f"fm_{l}"] for l in range(4)]
    for pool in (7, 14):
        out, _ = oracle.pyramid_roi_align_fwd(fms, g[f"pool{pool}_in_rois"][0], None, pool, float(size * size))
        np.testing.assert_array_equal(out, g[f"pool{pool}_out"])
    # mrn_refine (model.py:1389-1487)
    det = oracle.detection_layer(g["prop_out_rois"][0], g["det_in_probs"], g["det_in_deltas"], g["det_in_window"],
                                 float(g["det_in_min_conf"]), float(g["det_in_thr"]), int(g["det_in_limits"][0]),
                                 height=float(size), width=float(size))
    np.testing.assert_array_equal(det[:, :4], g["det_out_boxes"][0])
    np.testing.assert_array_equal(det[:, 4], g["det_out_scores"][0])
    np.testing.assert_array_equal(det[:, 5].astype(np.int64), g["det_out_class_ids"][0])
    np.testing.assert_array_equal(g["pool14_in_rois"][0], g["det_out_boxes"][0] / np.float32(size))   # model.py:1188
    # full_masks (data.py:287-314).  Random-init heads give some detections an empty (rounded, window-clipped) box, on
    # which the reference raises from PIL - and so does the oracle; the recording holds the reference's masks of the others
    d = len(det)
    h, w = (int(v) for v in g["mask_in_hw"])
    want, ok = detect_flow_expected_masks(g), g["mask_valid"]
    sel = np.ascontiguousarray(g["mask_in_sel"][:, None])
    if not ok.all():
        with pytest.raises(ValueError):
            oracle.full_masks(np.zeros(d, np.int64), g["det_out_boxes"][0], sel, h, w)
    got = oracle.full_masks(np.zeros(int(ok.sum()), np.int64), g["det_out_boxes"][0][ok], sel[ok], h, w)
    np.testing.assert_array_equal(got, want[ok])
    assert not want[~ok].any()
    # decode_masks (data.py:265-284) back to the original frame
    y1, x1, y2, x2 = (int(v) for v in g["det_in_window"])
    dec = oracle.decode_masks(want[ok], float(g["decode_in_scale"]), (y2 - y1, x2 - x1))
    assert dec.shape[1:] == tuple(int(v) for v in g["decode_out_hw"])
    assert hashlib.sha256(dec.tobytes()).digest() == g["decode_out_sha256"].tobytes()
    return int(ok.sum())


def golden_pyr_wide():
    """tests/golden/golden_pyr_wide_v1.npz (make_golden_pyr_wide.py): the reference's model.roi_align forward + autograd
    backward on a C = 40 pyramid -> (feature maps fp32 [1,40,s,s] x 4, boxes, image_shape, {pool: (out, grads_in, [gfm x 4])})."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_pyr_wide_v1.npz"))
    fms = [g[f"in_fm{l}"].astype(np.float32) for l in range(4)]
    pools = {p: (g[f"pool{p}_out"], g[f"pool{p}_in_grads"].astype(np.float32), [g[f"pool{p}_out_gfm{l}"] for l in range(4)])
             for p in (7, 14)}
    return fms, g["in_boxes"], [int(v) for v in g["in_image_shape"]], pools


def nms_edge_inputs(n=1500):
    """Box sets that stress the NMS decision at a size the grid-wide fixed-point route takes: {name: dets [n,5]} - identical boxes,
    disjoint boxes, degenerate (zero-height, inverted, duplicated) boxes, coordinates whose areas overflow, NaN coordinates, and a
    plain clustered set.  Scores are unique."""
    import numpy as np
    from maskrcnn_b200 import synth
    rng = np.random.default_rng(8)
    base = synth.random_rois(n, 8, image=1024.0, min_size=16, max_size=400) * 1024.0
    base[n // 2:] = base[:n - n // 2] + rng.uniform(-6, 6, (n - n // 2, 4)).astype(np.float32)
    scores = synth.unique_scores(n, 9)[:, None]
    cases = {"clustered": base}
    cases["identical"] = np.tile(np.array([[10, 10, 50, 50]], np.float32), (n, 1))
    cases["disjoint"] = np.stack([np.arange(n) * 20.0, np.zeros(n), np.arange(n) * 20.0 + 10, np.full(n, 10.0)], 1).astype(np.float32)
    deg = base.copy()
    deg[::7, 2] = deg[::7, 0] - 1.0                      # height 0 with the +1 convention
    deg[3::11, [0, 2]] = deg[3::11, [2, 0]]              # inverted in y: negative area
    deg[5::13] = deg[4::13][: len(deg[5::13])]           # exact duplicates
    cases["degenerate"] = deg
    big = base.copy()
    big[::5] *= 1e18                                     # areas overflow to inf
    cases["huge"] = big
    nan = base.copy()
    nan[::9, 1] = np.nan
    cases["nan"] = nan
    return {k: np.concatenate([v, scores], 1).astype(np.float32) for k, v in cases.items()}


NMS_EDGE_THRESHOLDS = (0.5, 0.0, 1.0, 0.7, float("nan"), -0.5)
