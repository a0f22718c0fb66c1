"""Pins the plain-C oracle (oracle/mrcnn_oracle.c) against the reference itself, executed here:
the reference's compiled CPU extension (oracle/_ref, built unmodified from /root/reference) and the
reference's unmodified model.py.  Skipped where the reference is absent (the GPU box); the same
comparisons are frozen as golden vectors in tests/golden/ (tests/test_oracle_golden.py)."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import reference
from maskrcnn_b200 import synth

pytestmark = pytest.mark.skipif(not reference.ref_C_available(), reason="oracle/_ref not built")


def _ref_crop_fwd(img, boxes, ind, ch, cw, ev):
    C = reference.ref_C()
    out = torch.zeros(1)
    with reference.quiet_stdout():
        C.crop_forward(torch.from_numpy(img), torch.from_numpy(boxes), torch.from_numpy(ind), ev, ch, cw, out)
    return out.numpy()


def _ref_crop_bwd(g, boxes, ind, shape):
    C = reference.ref_C()
    gi = torch.empty(shape)
    C.crop_backward(torch.from_numpy(g), torch.from_numpy(boxes), torch.from_numpy(ind), gi)
    return gi.numpy()


def _boxes_px(n, seed, size=1024.0):
    b = synth.random_rois(n, seed, image=size, min_size=size / 64, max_size=size * 0.7) * size
    return b


@pytest.mark.parametrize("n,thr", [(1, 0.5), (63, 0.3), (64, 0.7), (65, 0.5), (1000, 0.3), (3000, 0.7)])
def test_nms_bit_exact(n, thr):
    rng = np.random.default_rng(n)
    b = _boxes_px(n, n)
    # cluster boxes so that a good share is suppressed
    b[n // 2:] = b[: n - n // 2] + rng.uniform(-6, 6, (n - n // 2, 4)).astype(np.float32)
    dets = np.concatenate([b, synth.unique_scores(n, n)[:, None]], 1).astype(np.float32)
    want = reference.ref_C().nms(torch.from_numpy(dets), thr).numpy()
    got = oracle.nms(dets, thr)
    assert got.dtype == np.int64
    np.testing.assert_array_equal(got, want)
    assert 0 < len(want) <= n


def test_nms_edge_inputs_at_fixed_point_size():
    """1500-box sets with degenerate, huge, NaN and duplicated boxes and thresholds 0 / 1 / negative / NaN: the oracle follows the
    reference's compiled nms on every one (the GPU test test_nms_edge_cases_fixed_point_route checks the CUDA path against it)."""
    import helpers
    C = reference.ref_C()
    for name, d in helpers.nms_edge_inputs().items():
        for thr in helpers.NMS_EDGE_THRESHOLDS:
            np.testing.assert_array_equal(oracle.nms(d, thr), C.nms(torch.from_numpy(d), thr).numpy(), err_msg="%s thr=%s" % (name, thr))


def test_nms_edge_cases():
    C = reference.ref_C()
    assert len(oracle.nms(np.zeros((0, 5), np.float32), 0.5)) == 0
    assert C.nms(torch.zeros(0, 5), 0.5).numel() == 0
    # degenerate / inverted boxes, exact-threshold IoU, NaN IoU (0/0)
    dets = np.array([[0, 0, 9, 9, 0.9], [0, 0, 9, 9, 0.8], [0, 0, 9, 4, 0.7],   # IoU = 0.5 exactly
                     [5, 5, 2, 2, 0.6], [5, 5, 2, 2, 0.5],                       # y2<y1
                     [0, 0, -1, -1, 0.4], [0, 0, -1, -1, 0.3],                   # area 0 -> 0/0
                     [100, 100, 120, 130, 0.2]], np.float32)
    for thr in (0.5, 0.3, 0.0, 1.0):
        np.testing.assert_array_equal(oracle.nms(dets, thr), C.nms(torch.from_numpy(dets), thr).numpy())


@pytest.mark.parametrize("B,C,H,W,N,ch,cw,ev", [
    (1, 3, 16, 16, 7, 7, 7, 0.0), (2, 5, 32, 24, 33, 14, 14, 0.0), (3, 1, 64, 64, 9, 28, 28, 0.0),
    (1, 4, 8, 8, 5, 1, 1, 0.0), (2, 2, 9, 13, 6, 1, 5, -1.5), (1, 2, 33, 17, 11, 3, 1, 2.0)])
def test_crop_fwd_bwd_bit_exact(B, C, H, W, N, ch, cw, ev):
    rng = np.random.default_rng(B * 1000 + N)
    img = rng.standard_normal((B, C, H, W), dtype=np.float32)
    boxes = synth.random_rois(N, N, image=64.0, min_size=4, max_size=60)
    # some boxes leave the image (extrapolation), one is inverted, one is a point
    boxes[0] += 0.3
    boxes[1] -= 0.25
    if N > 4:
        boxes[2] = boxes[2][[2, 3, 0, 1]]
        boxes[3] = [0.5, 0.5, 0.5, 0.5]
        boxes[4] = [0.0, 0.0, 1.0, 1.0]
    ind = rng.integers(0, B, N).astype(np.int32)
    want = _ref_crop_fwd(img, boxes, ind, ch, cw, ev)
    got = oracle.crop_forward(img, boxes, ind, ch, cw, ev)
    assert got.shape == want.shape
    np.testing.assert_array_equal(got, want)
    g = rng.standard_normal(want.shape, dtype=np.float32)
    np.testing.assert_array_equal(oracle.crop_backward(g, boxes, ind, img.shape),
                                  _ref_crop_bwd(g, boxes, ind, img.shape))


def test_nms_and_crop_hypothesis_against_the_compiled_reference():
    """Random sizes / thresholds / boxes drawn by hypothesis: the C restatement against the reference's own compiled
    nms / crop_forward / crop_backward (oracle/_ref), bit for bit."""
    from hypothesis import given, settings, strategies as st
    C = reference.ref_C()

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(n=st.integers(1, 400), seed=st.integers(0, 10 ** 6), thr=st.sampled_from([0.0, 0.05, 0.3, 0.5, 0.7, 0.99, 1.0]),
           jitter=st.sampled_from([0.0, 0.5, 4.0, 40.0]), integer=st.booleans())
    def check_nms(n, seed, thr, jitter, integer):
        rng = np.random.default_rng(seed)
        b = _boxes_px(n, seed, size=256.0)
        if n > 1:
            h = n // 2
            b[h:] = b[: n - h] + rng.uniform(-jitter, jitter, (n - h, 4)).astype(np.float32)
        if integer:
            b = np.round(b)       # rounded pixel boxes as the detection layer emits: exact-threshold IoUs become likely
        dets = np.concatenate([b, synth.unique_scores(n, seed + 1)[:, None]], 1).astype(np.float32)
        np.testing.assert_array_equal(oracle.nms(dets, thr), C.nms(torch.from_numpy(dets), thr).numpy())
    check_nms()

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(B=st.integers(1, 3), Cc=st.integers(1, 5), H=st.integers(1, 40), W=st.integers(1, 40), N=st.integers(1, 12),
           ch=st.integers(1, 15), cw=st.integers(1, 15), ev=st.sampled_from([0.0, -1.5, 3.0]), seed=st.integers(0, 10 ** 6))
    def check_crop(B, Cc, H, W, N, ch, cw, ev, seed):
        rng = np.random.default_rng(seed)
        img = rng.standard_normal((B, Cc, H, W), dtype=np.float32)
        boxes = rng.uniform(-0.3, 1.3, (N, 4)).astype(np.float32)          # inverted, outside, degenerate: everything goes
        if N > 2:
            boxes[0] = [0.0, 0.0, 1.0, 1.0]
            boxes[1, 2:] = boxes[1, :2]
        ind = rng.integers(0, B, N).astype(np.int32)
        want = _ref_crop_fwd(img, boxes, ind, ch, cw, ev)
        np.testing.assert_array_equal(oracle.crop_forward(img, boxes, ind, ch, cw, ev), want)
        g = rng.standard_normal(want.shape, dtype=np.float32)
        np.testing.assert_array_equal(oracle.crop_backward(g, boxes, ind, img.shape), _ref_crop_bwd(g, boxes, ind, img.shape))
    check_crop()


def test_crop_bad_box_index():
    img = np.zeros((1, 1, 4, 4), np.float32)
    with pytest.raises(oracle.OracleError):
        oracle.crop_forward(img, np.array([[0, 0, 1, 1]], np.float32), np.array([1], np.int32), 2, 2)


needs_model = pytest.mark.skipif(not reference.available(), reason="/root/reference absent")


def _small_pyramid(C, size, seed):
    return synth.feature_pyramid(1, C, seed, image=size)


@needs_model
@pytest.mark.parametrize("pool", [7, 14])
def test_pyramid_roi_align_matches_model_py(pool):
    ref = reference.load()
    size, C, N = 512, 6, 300
    fms = _small_pyramid(C, size, 3)
    boxes = synth.random_rois(N, 5, image=float(size), min_size=8, max_size=size * 0.9)
    t_fms = [torch.from_numpy(f).clone().requires_grad_(True) for f in fms]
    inputs = [torch.from_numpy(boxes).unsqueeze(0)] + list(t_fms)
    want = ref.model.roi_align(inputs, pool, [size, size, 3])
    got, lv = oracle.pyramid_roi_align_fwd(fms, boxes, None, pool, float(size * size))
    assert set(np.unique(lv)) == {2, 3, 4, 5}
    np.testing.assert_array_equal(got, want.detach().numpy())
    g = np.random.default_rng(9).standard_normal(got.shape, dtype=np.float32)
    want.backward(torch.from_numpy(g))
    gf = oracle.pyramid_roi_align_bwd(g, [f.shape for f in fms], boxes, None, float(size * size))
    for a, b in zip(gf, t_fms):
        np.testing.assert_array_equal(a, b.grad.numpy())


@needs_model
def test_roi_level_matches_model_py_at_full_scale():
    ref = reference.load()
    boxes = synth.random_rois(20000, 77)
    b = torch.from_numpy(boxes)
    h = b[:, 2] - b[:, 0]
    w = b[:, 3] - b[:, 1]
    area = torch.FloatTensor([1024.0 * 1024.0])
    want = (4 + torch.log2(torch.sqrt(h * w) / (224.0 / torch.sqrt(area)))).round().int().clamp(2, 5)  # model.py:331-338
    got = oracle.roi_levels(boxes, 1024.0 * 1024.0)
    np.testing.assert_array_equal(got, want.numpy())


def _ulp_diff(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


@needs_model
@pytest.mark.parametrize("size,seed,max_rois,thr,clusters", [(256, 11, 200, 0.7, 6), (512, 12, 500, 0.7, 12), (128, 13, 100, 0.5, 3),
                                                            (256, 14, 1000, 0.9, 4)])
def test_proposal_layer_matches_model_py(size, seed, max_rois, thr, clusters):
    ref = reference.load()
    import types
    anchors = synth.pyramid_anchors((size, size))
    rc, rb = synth.rpn_outputs(anchors, seed, image=float(size), n_clusters=clusters)
    cfg = types.SimpleNamespace(RPN_NMS_MAX_ROIS_NUM=max_rois, RPN_NMS_THRESHOLD=thr,
                                RPN_BBOX_STD_DEV=[0.1, 0.1, 0.2, 0.2], IMAGE_SHAPE=np.array([size, size, 3]),
                                GPU_COUNT=0)
    stub = types.SimpleNamespace(config=cfg, anchors=torch.from_numpy(anchors))
    want = ref.model.MaskRCNN.rpn_refine(stub, torch.from_numpy(rc).unsqueeze(0),
                                         torch.from_numpy(rb).unsqueeze(0))[0].numpy()
    got = oracle.proposal_layer(rc, rb, anchors, 500, max_rois, thr, height=float(size), width=float(size))  # model.py:1345
    assert got.shape == want.shape and 20 < len(got) <= max_rois
    assert _ulp_diff(got, want).max() <= 4   # torch.exp on CPU is not correctly rounded


@needs_model
@pytest.mark.parametrize("N,NC,window,seed", [(400, 81, (0, 0, 1024, 1024), 21), (300, 5, (64, 32, 900, 1000), 23),
                                              (1000, 81, (0, 128, 1024, 896), 25), (150, 2, (0, 0, 1024, 1024), 27)])
def test_detection_layer_matches_model_py(N, NC, window, seed):
    """81 classes and few-class cases (nearly every pair shares a class: heavy per-class suppression), windows that clip."""
    ref = reference.load()
    import types
    rois = synth.random_rois(N, seed)
    probs, deltas = synth.head_outputs(N, NC, seed + 1)
    window = np.array(window, np.float32)
    for min_conf, max_inst in ((0, 100), (0.7, 50)):
        cfg = types.SimpleNamespace(RPN_BBOX_STD_DEV=np.array([0.1, 0.1, 0.2, 0.2]), IMAGE_SHAPE=np.array([1024, 1024, 3]),
                                    GPU_COUNT=0, DETECTION_MIN_CONFIDENCE=min_conf, DETECTION_NMS_THRESHOLD=0.3,
                                    DETECTION_MAX_INSTANCES=max_inst)
        stub = types.SimpleNamespace(config=cfg)
        ci, sc, bx = ref.model.MaskRCNN.mrn_refine(stub, torch.from_numpy(rois).unsqueeze(0), torch.from_numpy(probs),
                                                   torch.from_numpy(deltas), window)
        got = oracle.detection_layer(rois, probs, deltas, window, min_conf, 0.3, max_inst)
        if ci is None:
            assert len(got) == 0
            continue
        assert len(got) == ci.shape[1] and len(got) > 0
        np.testing.assert_array_equal(got[:, 5].astype(np.int64), ci[0].numpy())
        np.testing.assert_array_equal(got[:, 4], sc[0].numpy())
        np.testing.assert_array_equal(got[:, :4], bx[0].numpy())


@needs_model
@pytest.mark.parametrize("n_rois,n_gt,n_crowd,n_pad,train_rois,seed", [
    (600, 12, 0, 0, 512, 1), (600, 12, 2, 3, 512, 2), (300, 5, 0, 2, 100, 3), (64, 3, 1, 0, 512, 4), (200, 6, 0, 0, 32, 5)])
def test_target_layer_matches_model_py(n_rois, n_gt, n_crowd, n_pad, train_rois, seed):
    """model.mrn_samples (unmodified, CPU) vs the oracle under the same torch seed: the oracle draws its two
    permutations with torch.randperm in the reference's order, so selections and their order must be identical."""
    ref = reference.load()
    import types
    S = 256
    rois, cls, gt, masks = synth.target_inputs(n_rois, n_gt, seed, image=S, n_crowd=n_crowd, n_pad=n_pad)
    cfg = types.SimpleNamespace(GPU_COUNT=0, TRAIN_ROIS_PER_IMAGE=train_rois, ROI_POSITIVE_RATIO=0.33,
                                BBOX_STD_DEV=np.array([0.1, 0.1, 0.2, 0.2]), MASK_SHAPE=[28, 28])
    torch.manual_seed(100 + seed)
    w_rois, w_cls, w_d, w_m = ref.model.mrn_samples(torch.from_numpy(rois)[None], torch.from_numpy(cls)[None],
                                                    torch.from_numpy(gt)[None], torch.from_numpy(masks)[None], cfg)
    torch.manual_seed(100 + seed)
    g_rois, g_cls, g_d, g_m = oracle.mrn_samples(rois, cls, gt, masks, train_rois, 0.33, [0.1, 0.1, 0.2, 0.2], (28, 28),
                                                 lambda n: torch.randperm(n).numpy())
    assert len(g_rois) == len(w_rois) and len(g_rois) > 10
    np.testing.assert_array_equal(g_rois, w_rois.numpy())
    np.testing.assert_array_equal(g_cls, w_cls.numpy())
    np.testing.assert_array_equal(g_m, w_m.numpy())
    assert (g_cls > 0).sum() >= 5 and g_m.sum() > 0
    np.testing.assert_array_equal(g_d[:, :2], w_d.numpy()[:, :2])          # dy, dx: + - * / only
    assert _ulp_diff(g_d[:, 2:], w_d.numpy()[:, 2:]).max() <= 2            # dh, dw: torch.log on CPU is not correctly rounded


@needs_model
def test_target_layer_no_positive():
    ref = reference.load()
    import types
    rois, cls, gt, masks = synth.target_inputs(50, 3, 9, image=128, positive_fraction=0.0)
    rois[:] = [0.0, 0.0, 0.01, 0.01]
    cfg = types.SimpleNamespace(GPU_COUNT=0, TRAIN_ROIS_PER_IMAGE=64, ROI_POSITIVE_RATIO=0.33,
                                BBOX_STD_DEV=np.array([0.1, 0.1, 0.2, 0.2]), MASK_SHAPE=[28, 28])
    w = ref.model.mrn_samples(torch.from_numpy(rois)[None], torch.from_numpy(cls)[None], torch.from_numpy(gt)[None],
                              torch.from_numpy(masks)[None], cfg)
    g = oracle.mrn_samples(rois, cls, gt, masks, 64, 0.33, [0.1, 0.1, 0.2, 0.2], (28, 28), lambda n: torch.randperm(n).numpy())
    assert all(t.numel() == 0 for t in w) and all(len(a) == 0 for a in g)   # model.py:563-574


@needs_model
@pytest.mark.parametrize("image,n_gt,n_crowd,train_anchors,seed", [(256, 6, 0, 256, 1), (256, 9, 2, 64, 2), (512, 15, 1, 256, 3), (128, 1, 0, 32, 4)])
def test_rpn_samples_matches_data_py(image, n_gt, n_crowd, train_anchors, seed):
    """data.rpn_samples (unmodified numpy code) vs the oracle under the same numpy seed: the oracle's permutation callback
    is np.random.permutation, which is what np.random.choice(ids, extra, replace=False) draws internally."""
    ref = reference.load()
    import types
    anchors = synth.pyramid_anchors((image, image)).astype(np.float64)
    cls, gt = synth.rpn_target_inputs(n_gt, seed, image=image, n_crowd=n_crowd)
    cfg = types.SimpleNamespace(RPN_TRAIN_ANCHORS_PER_IMAGE=train_anchors, RPN_BBOX_STD_DEV=np.array([0.1, 0.1, 0.2, 0.2]))
    np.random.seed(200 + seed)
    w_match, w_bbox = ref.data.rpn_samples(anchors, cls, gt, cfg)
    np.random.seed(200 + seed)
    g_match, g_bbox = oracle.rpn_samples(anchors, cls, gt, train_anchors, [0.1, 0.1, 0.2, 0.2], np.random.permutation)
    np.testing.assert_array_equal(g_match, w_match)
    assert (g_match == 1).sum() >= 1 and (g_match == -1).sum() >= 1
    np.testing.assert_array_equal(g_bbox[:, :2], w_bbox[:, :2])
    a, b = g_bbox[:, 2:].view(np.int64), w_bbox[:, 2:].view(np.int64)
    assert np.abs(a - b).max() <= 2      # float64 log: numpy's vs libm's


@needs_model
@pytest.mark.parametrize("H,W,lo,hi,seed", [(256, 256, 4, 200, 1), (200, 333, 1, 60, 2), (64, 80, 1, 28, 3), (512, 512, 28, 500, 4)])
def test_full_masks_matches_data_py(H, W, lo, hi, seed):
    """oracle.full_masks (a restatement of Pillow's 8-bit bilinear resample) against data.full_masks executed here, i.e.
    against the installed Pillow: upscales, downscales, unchanged sides, boxes leaving the image, fractional coordinates,
    saturated mask values."""
    d = reference.load().data
    rng = np.random.default_rng(seed)
    cls, boxes, masks = synth.mask_head_outputs(24, 7, 50 + seed, image=min(H, W), min_size=lo, max_size=hi)
    masks[::5] = masks[::5] * 1.5 - 0.25
    boxes[1, 2] = boxes[1, 0] + 28.0
    boxes[2, 3] = boxes[2, 1] + 28.0
    boxes[3] += np.float32([-9.0, -11.0, -9.0, -11.0])
    boxes[4, 2:] = np.maximum(boxes[4, 2:], boxes[4, :2] + 3.0)
    boxes[4] += rng.choice(np.float32([0.25, 0.5, 0.75]), 4)
    boxes[5, 2:] = [H + 13.0, W + 5.0]
    want = d.full_masks(torch.from_numpy(cls), torch.from_numpy(boxes), torch.from_numpy(masks), H, W).numpy()
    got = oracle.full_masks(cls, boxes, masks, H, W)
    assert want.dtype == np.bool_ and got.shape == want.shape
    np.testing.assert_array_equal(got, want)


@needs_model
def test_full_masks_empty_box_raises_like_pil():
    d = reference.load().data
    cls, boxes, masks = synth.mask_head_outputs(2, 3, 9, image=64)
    boxes[1] = [10.0, 10.0, 10.0, 30.0]
    with pytest.raises(ValueError):
        d.full_masks(torch.from_numpy(cls), torch.from_numpy(boxes), torch.from_numpy(masks), 64, 64)
    with pytest.raises(ValueError):
        oracle.full_masks(cls, boxes, masks, 64, 64)


def _decode_case(D, H, W, seed):
    """bool masks [D,H,W]: filled rectangles, discs, thin lines and isolated pixels (every edge gets interpolated)."""
    rng = np.random.default_rng(seed)
    m = np.zeros((D, H, W), bool)
    yy, xx = np.mgrid[0:H, 0:W]
    for i in range(D):
        y, x = rng.integers(0, H - 2), rng.integers(0, W - 2)
        h, w = rng.integers(1, max(2, H - y)), rng.integers(1, max(2, W - x))
        m[i, y:y + h, x:x + w] = True
        r = rng.integers(1, max(2, min(H, W) // 3))
        m[i] ^= (yy - rng.integers(0, H)) ** 2 + (xx - rng.integers(0, W)) ** 2 < r * r
        m[i, rng.integers(0, H), :] = True
        m[i, rng.integers(0, H, 20), rng.integers(0, W, 20)] = True
    return m


@needs_model
@pytest.mark.parametrize("H,W,window,scale,seed", [
    (256, 256, (48, 0, 208, 256), 256 / 1920, 1),          # the predict.py geometry at 256: 1920x1200 frame, upscale 7.5
    (256, 256, (0, 33, 256, 222), 0.75, 2),               # odd margins: CenterCrop's round-half-even origin
    (200, 200, (13, 0, 186, 200), 0.4161, 3),             # 173 rows: (200 - 173) / 2 = 13.5 -> 14, not the window's 13
    (128, 160, (0, 0, 128, 160), 2.0, 4),                 # the frame was enlarged (IMAGE_MIN_DIM): downscale by 2
    (120, 90, (10, 5, 111, 86), 3.3, 5),                  # downscale by 3.3: up to 9 taps
    (64, 64, (0, 0, 64, 64), 0.5, 6), (64, 64, (2, 2, 62, 62), 1.0001, 7)])
def test_decode_masks_matches_data_py(H, W, window, scale, seed):
    """oracle.decode_masks against data.decode_masks executed here (PIL '1' -> 'L', torchvision CenterCrop + Resize)."""
    d = reference.load().data
    m = _decode_case(5, H, W, seed)
    box = d.Box.fromlist(list(window))
    want = d.decode_masks(torch.from_numpy(m), scale, box).numpy()
    got = oracle.decode_masks(m, scale, (box.height(), box.width()))
    assert want.dtype == np.uint8 and got.shape == want.shape
    np.testing.assert_array_equal(got, want)
    same_size = want.shape[1:] == (box.height(), box.width())      # Resample skips both passes: a plain crop
    assert ((want > 0) & (want < 255)).any() or scale > 1.9 or same_size     # interpolated edge values are compared too
    # uint8 0/1 input ("tensor NxHxW with 1/0"): stays 0/1 through convert('L')
    want1 = d.decode_masks(torch.from_numpy(m.astype(np.uint8)), scale, box).numpy()
    np.testing.assert_array_equal(oracle.decode_masks(m.astype(np.uint8), scale, (box.height(), box.width())), want1)


@needs_model
def test_decode_masks_scale_one_is_identity():
    d = reference.load().data
    m = _decode_case(2, 32, 32, 1)
    t = torch.from_numpy(m)
    assert d.decode_masks(t, 1, d.Box.fromlist([0, 0, 32, 32])) is t
    assert oracle.decode_masks(m, 1, (32, 32)) is m


@needs_model
def test_masks_hypothesis_against_data_py():
    """Random geometry drawn by hypothesis through the reference's PIL / torchvision route (data.full_masks,
    data.decode_masks executed here) against the oracle's restatement of Pillow's resample: bit for bit."""
    from hypothesis import given, settings, strategies as st
    d = reference.load().data

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(H=st.integers(8, 96), W=st.integers(8, 96), mh=st.sampled_from([7, 14, 28]), seed=st.integers(0, 10 ** 6),
           n=st.integers(1, 6))
    def check_paste(H, W, mh, seed, n):
        rng = np.random.default_rng(seed)
        y1 = rng.integers(-4, H - 1, n)
        x1 = rng.integers(-4, W - 1, n)
        boxes = np.stack([y1, x1, y1 + rng.integers(1, H + 6, n), x1 + rng.integers(1, W + 6, n)], 1).astype(np.float32)
        boxes += rng.choice(np.float32([0.0, 0.0, 0.25, 0.5]), (n, 4))           # rounded boxes mostly, some fractional
        boxes[:, 2:] = np.maximum(boxes[:, 2:], boxes[:, :2] + 1.0)
        cls = rng.integers(0, 3, n).astype(np.int64)
        masks = rng.uniform(-0.2, 1.2, (n, 3, mh, mh)).astype(np.float32)
        want = d.full_masks(torch.from_numpy(cls), torch.from_numpy(boxes), torch.from_numpy(masks), H, W).numpy()
        np.testing.assert_array_equal(oracle.full_masks(cls, boxes, masks, H, W), want)
    check_paste()

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(H=st.integers(4, 80), W=st.integers(4, 80), seed=st.integers(0, 10 ** 6),
           scale=st.sampled_from([0.2, 0.3333, 0.5, 0.53333336, 0.77, 0.999, 1.25, 1.7, 2.0, 3.1]))
    def check_decode(H, W, seed, scale):
        rng = np.random.default_rng(seed)
        m = rng.random((2, H, W)) < rng.uniform(0.1, 0.9)
        m[0, H // 4:H // 2 + 1, W // 4:W // 2 + 1] = True
        ch, cw = int(rng.integers(1, H + 1)), int(rng.integers(1, W + 1))
        y0, x0 = (H - ch) // 2, (W - cw) // 2
        box = d.Box.fromlist([y0, x0, y0 + ch, x0 + cw])
        if round(ch * 1.0 / scale) < 1 or round(cw * 1.0 / scale) < 1:
            return                                                              # PIL raises; covered by the edge-case tests
        want = d.decode_masks(torch.from_numpy(m), scale, box).numpy()
        np.testing.assert_array_equal(oracle.decode_masks(m, scale, (ch, cw)), want)
    check_decode()


@needs_model
def test_rpn_pack_matches_model_py():
    """oracle.rpn_pack against the reference's RPN module (model.py:573-653) and the torch.cat of rpn_detect (:1294-1304)."""
    m = reference.load().model
    torch.manual_seed(5)
    rpn = m.RPN(3, 1, 16)
    feats = [torch.randn(2, 16, s, s + 2) * 3.0 for s in (16, 8, 4, 2, 1)]
    with torch.no_grad():
        want = [torch.cat(list(o), dim=1).numpy() for o in zip(*[rpn(f) for f in feats])]
        ls, bs = [], []
        for f in feats:
            x = rpn.relu(rpn.conv_shared(rpn.padding(f)))
            ls.append(rpn.conv_class(x).numpy())
            bs.append(rpn.conv_bbox(x).numpy())
    got = oracle.rpn_pack(ls, bs)
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_array_equal(got[2], want[2])
    assert np.abs(got[1] - want[1]).max() <= 1e-6   # torch's CPU softmax: approximate vectorised exp


@needs_model
@pytest.mark.parametrize("image_name,dim,seed", [("girl.jpg", 192, 7), ("messi.jpg", 320, 11)])
def test_detect_flow_live(image_name, dim, seed):
    """The reference's whole predict.py flow executed HERE on other images / frame sizes / seeds than the committed
    fixture (tests/golden/make_golden_detect.py records it at every replaced operator's boundary), replayed stage by
    stage with the oracle."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_detect
    from helpers import check_detect_flow_with_oracle
    g = make_golden_detect.run(seed, image_name=image_name, image_dim=dim, save=False)
    assert check_detect_flow_with_oracle(g) >= 1
