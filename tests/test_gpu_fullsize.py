"""GPU parity at the EXACT sizes of BASELINE.json's configs, against the oracle and - where oracle/_ref travelled to the
box - against the reference's own compiled CPU extension (vision.cpp:11-15 crop_forward / crop_backward / nms):

  configs[1]  proposal layer, 261,888 anchors, top-6000 -> NMS 0.7 -> 1000, batch 8 (eight distinct images), both NMS
              algorithms                                                    model.py:1307-1382, nms_cpu.cpp:11-70
  configs[2]  PyramidROIAlign forward, 1000 RoIs x 256 ch on the 1024^2 pyramid, 7x7 and 14x14, NCHW and channels-last
              pyramids, NCHW and channels-last crops                        model.py:276-393, crop_cpu.cpp:119-164
  configs[3]  training step geometry: batch 16 x 512 RoIs x 256 ch launched as ONE call; images 3 and 15 checked in full
              (forward bit-exact, gather and scatter backward <= 1e-5)      crop_cpu.cpp:167-265
  configs[4]  detection layer 64 images x 1000 RoIs x 81 classes + 14x14 mask RoIAlign of the detections

Bit-exact for selections and the forward interpolation; <= 1e-5 relative (scale = max |reference|) for the backward, the
tolerance north_star states."""
import numpy as np
import pytest
import torch

import oracle
from helpers import golden_pyr_wide, rel_err
from maskrcnn_b200 import synth

pytestmark = pytest.mark.gpu

BWD_TOL = 1e-5
IMAGE = 1024
LEVEL_HW = [(256, 256), (128, 128), (64, 64), (32, 32)]


@pytest.fixture(scope="module")
def ops():
    import maskrcnn_b200
    return maskrcnn_b200


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def ref_C_or_none():
    from oracle import reference
    return reference.ref_C() if reference.ref_C_available() else None


# ------------------------------------------------------------------ configs[1]
@pytest.mark.parametrize("algo", ["lazy", "mask", "hybrid"])
def test_proposal_layer_config1_vs_oracle(ops, algo):
    """261,888 anchors, 6000 -> 1000 at IoU 0.7, batch 8: every image's proposals bit-exact against the oracle; the NMS
    input of image 0 (the 6000 decoded, clipped boxes in score order) also goes through the reference's own nms."""
    anchors = synth.pyramid_anchors((IMAGE, IMAGE))
    assert len(anchors) == 261888
    rcs, rbs = zip(*[synth.rpn_outputs(anchors, 4100 + i) for i in range(8)])
    ops.set_proposal_nms(algo)
    try:
        rois, counts = ops.proposal_layer(dev(np.stack(rcs)), dev(np.stack(rbs)), dev(anchors), 6000, 1000, 0.7)
    finally:
        ops.set_proposal_nms("auto")
    rois, counts = rois.cpu().numpy(), counts.cpu().numpy()
    for i in range(8):
        want = oracle.proposal_layer(rcs[i], rbs[i], anchors, 6000, 1000, 0.7, height=float(IMAGE), width=float(IMAGE))
        assert counts[i] == len(want), "image %d: %d proposals, oracle %d" % (i, counts[i], len(want))
        np.testing.assert_array_equal(rois[i, :counts[i]], want)
        assert not rois[i, counts[i]:].any()
    C = ref_C_or_none()
    if C is not None:
        # the reference's own NMS on the oracle's pre-NMS boxes of image 0 (model.py:1364-1366), then :1371-1374
        _, d5, _ = oracle.proposal_layer(rcs[0], rbs[0], anchors, 6000, 1000, 0.7, height=float(IMAGE), width=float(IMAGE),
                                         return_intermediate=True)                    # [6000,5] boxes px + score, score order
        keep = C.nms(torch.from_numpy(d5), 0.7).numpy()[:1000]
        want = d5[keep, :4] / np.array([IMAGE, IMAGE, IMAGE, IMAGE], np.float32)
        np.testing.assert_array_equal(rois[0, :counts[0]], want)


# ------------------------------------------------------------------ configs[2]
@pytest.mark.parametrize("pool", [7, 14])
def test_roialign_forward_config2_vs_oracle_and_reference(ops, pool):
    """1000 RoIs x 256 channels on P2..P5 of a 1024^2 image: all four (pyramid layout, crop layout) combinations bit-exact
    against the oracle, and the oracle's rows of every level bit-exact against the reference's crop_forward."""
    fms = synth.feature_pyramid(1, 256, 77)
    boxes = synth.random_rois(1000, 1234)
    want, lv = oracle.pyramid_roi_align_fwd(fms, boxes, None, pool, float(IMAGE * IMAGE))
    assert want.shape == (1000, 256, pool, pool) and len(np.unique(lv)) == 4
    C = ref_C_or_none()
    if C is not None:
        from oracle import reference
        for l in range(4):
            sel = np.nonzero(lv == l + 2)[0]
            crops = torch.zeros(1)
            with reference.quiet_stdout():
                C.crop_forward(torch.from_numpy(fms[l]), torch.from_numpy(boxes[sel]), torch.zeros(len(sel), dtype=torch.int32),
                               0.0, pool, pool, crops)
            np.testing.assert_array_equal(want[sel], crops.numpy())
    b = dev(boxes)
    for pyr_cl in (True, False):
        ts = [cl(dev(f)) if pyr_cl else dev(f) for f in fms]
        for out_cl in (True, False):
            out = ops.pyramid_roi_align(ts, b, None, pool, (IMAGE, IMAGE, 3), out_channels_last=out_cl)
            assert out.is_contiguous(memory_format=torch.channels_last if out_cl else torch.contiguous_format)
            np.testing.assert_array_equal(out.cpu().numpy(), want, err_msg="pyramid cl=%s crops cl=%s" % (pyr_cl, out_cl))
        # the reference-shaped call (model.py:276: inputs = [boxes [1,N,4], P2..P5])
        out = ops.roi_align([b.unsqueeze(0)] + ts, pool, [IMAGE, IMAGE, 3])
        np.testing.assert_array_equal(out.cpu().numpy(), want)
    ops.check_device_errors()


# ------------------------------------------------------------------ configs[3]
def _train_inputs(B=16, R=512):
    g = torch.Generator(device="cuda")
    g.manual_seed(4242)
    fm = [torch.randn((B, 256, h, w), device="cuda", generator=g) for h, w in LEVEL_HW]          # NCHW master copy
    boxes = np.concatenate([synth.random_rois(R, 7000 + 13 * i) for i in range(B)], 0)
    ind = np.repeat(np.arange(B, dtype=np.int32), R)
    return fm, boxes, ind


@pytest.mark.parametrize("pool", [7, 14])
@pytest.mark.parametrize("pyr_cl", [True, False])
def test_train_step_config3_vs_oracle(ops, pool, pyr_cl):
    """batch 16 x 512 RoIs x 256 ch as ONE launch (8192 RoIs, 1.43 GB pyramid): images 3 and 15 compared in full with the
    oracle - forward bit-exact, backward (every algorithm the layout offers) <= 1e-5 of max |reference|; image 15 also
    against the reference's own crop_backward."""
    B, R = 16, 512
    fm, boxes, ind = _train_inputs(B, R)
    check = (3, 15)
    g = torch.Generator(device="cuda")
    g.manual_seed(99 + pool)
    ts = [(cl(f) if pyr_cl else f.clone()).requires_grad_(True) for f in fm]
    grad = torch.randn((B * R, 256, pool, pool), device="cuda", generator=g)
    if pyr_cl:
        grad = cl(grad)
    algos = ("gather", "scatter") if pyr_cl else ("auto",)
    got = {}
    for algo in algos:
        ops.set_backward_algorithm(algo)
        try:
            for t in ts:
                t.grad = None
            out = ops.pyramid_roi_align(ts, dev(boxes), dev(ind), pool, (IMAGE, IMAGE, 3))
            out.backward(grad)
            got[algo] = (out.detach(), [t.grad for t in ts])
        finally:
            ops.set_backward_algorithm("auto")
    C = ref_C_or_none()
    for i in check:
        rs = slice(i * R, (i + 1) * R)
        f_i = [f[i:i + 1].cpu().numpy() for f in fm]
        want, lv = oracle.pyramid_roi_align_fwd(f_i, boxes[rs], None, pool, float(IMAGE * IMAGE))
        g_i = grad[rs].cpu().numpy()
        want_g = oracle.pyramid_roi_align_bwd(g_i, [f.shape for f in f_i], boxes[rs], None, float(IMAGE * IMAGE))
        if C is not None and i == 15:
            for l in range(4):
                sel = np.nonzero(lv == l + 2)[0]
                gi = torch.zeros((1, 256) + LEVEL_HW[l])
                C.crop_backward(torch.from_numpy(np.ascontiguousarray(g_i[sel])), torch.from_numpy(boxes[rs][sel]),
                                torch.zeros(len(sel), dtype=torch.int32), gi)
                np.testing.assert_array_equal(want_g[l], gi.numpy())       # the oracle IS the reference here
        for algo, (out, grads) in got.items():
            np.testing.assert_array_equal(out[rs].cpu().numpy(), want, err_msg="image %d forward (%s)" % (i, algo))
            for l in range(4):
                e = rel_err(grads[l][i:i + 1].cpu().numpy(), want_g[l])
                assert e <= BWD_TOL, "image %d level %d backward (%s): %.3g" % (i, l, algo, e)
    ops.check_device_errors()


def test_train_step_config3_pair_backward_vs_oracle(ops):
    """The fused two-head backward (7x7 + 14x14 into one gradient pyramid) at configs[3] size: image 9 against the sum of
    the oracle's two backward passes."""
    B, R = 16, 512
    fm, boxes, ind = _train_inputs(B, R)
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    g7 = cl(torch.randn((B * R, 256, 7, 7), device="cuda", generator=g))
    g14 = cl(torch.randn((B * R, 256, 14, 14), device="cuda", generator=g))
    both = ops.pyramid_roi_align_backward_pair(g7, g14, [tuple(f.shape) for f in fm], dev(boxes), dev(ind), (IMAGE, IMAGE, 3))
    i = 9
    rs = slice(i * R, (i + 1) * R)
    shapes = [(1, 256) + hw for hw in LEVEL_HW]
    a = oracle.pyramid_roi_align_bwd(g7[rs].cpu().numpy(), shapes, boxes[rs], None, float(IMAGE * IMAGE))
    b = oracle.pyramid_roi_align_bwd(g14[rs].cpu().numpy(), shapes, boxes[rs], None, float(IMAGE * IMAGE))
    for l in range(4):
        assert rel_err(both[l][i:i + 1].cpu().numpy(), a[l] + b[l]) <= BWD_TOL


def test_mask_targets_config3_vs_oracle(ops):
    """28x28 mask targets of configs[3]: 168 positives per image cropped out of [G,1,1024,1024] gt masks by (image,
    instance) index (model.py:492-507), two images' worth, bit-exact (round half to even included)."""
    G, P = 8, 168
    rng = np.random.default_rng(31)
    gt = np.zeros((2 * G, 1, IMAGE, IMAGE), np.float32)
    for k in range(2 * G):
        y, x = rng.integers(0, IMAGE - 64, 2)
        h, w = rng.integers(32, 512, 2)
        gt[k, 0, y:y + h, x:x + w] = 1.0
    boxes = np.concatenate([synth.random_rois(P, 800 + i) for i in range(2)], 0)
    ind = (np.repeat(np.arange(2), P) * G + rng.integers(0, G, 2 * P)).astype(np.int32)
    want = oracle.crop_forward(gt, boxes, ind, 28, 28, 0.0)
    got = ops.CropFunction(28, 28, 0)(dev(gt), dev(boxes), dev(ind))
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    np.testing.assert_array_equal(torch.round(got).cpu().numpy(), np.round(want))           # model.py:507


# ------------------------------------------------------------------ configs[4]
def test_detection_path_config4_vs_oracle(ops):
    """64 images x 1000 RoIs x 81 classes, NMS 0.3, top-100, then the 14x14 mask RoIAlign of the detections: every image's
    detections bit-exact against the oracle; the mask-head crops of images 0 and 63 bit-exact."""
    B, N, NC, D = 64, 1000, 81, 100
    rois = np.stack([synth.random_rois(N, 300 + i) for i in range(B)])
    heads = [synth.head_outputs(N, NC, 900 + i) for i in range(B)]
    probs, deltas = np.stack([h[0] for h in heads]), np.stack([h[1] for h in heads])
    win = np.tile(np.array([[0, 0, IMAGE, IMAGE]], np.float32), (B, 1))
    dets, counts = ops.detection_layer(dev(rois), dev(probs), dev(deltas), dev(win), 0.0, 0.3, D)
    dn, cn = dets.cpu().numpy(), counts.cpu().numpy()
    for i in range(B):
        want = oracle.detection_layer(rois[i], probs[i], deltas[i], win[i], 0.0, 0.3, D, height=float(IMAGE), width=float(IMAGE))
        assert cn[i] == len(want)
        np.testing.assert_array_equal(dn[i, :cn[i]], want)
        assert not dn[i, cn[i]:].any()
    fms = synth.feature_pyramid(2, 256, 5)
    ts = [cl(dev(f)) for f in fms]
    pick = (0, 63)
    b = (dets[list(pick), :, :4] / float(IMAGE)).reshape(-1, 4)
    ind = torch.arange(2, dtype=torch.int32, device="cuda").repeat_interleave(D)
    out = ops.pyramid_roi_align(ts, b, ind, 14, (IMAGE, IMAGE, 3)).cpu().numpy()
    for k, i in enumerate(pick):
        bb = dn[i, :, :4] / np.float32(IMAGE)                                              # model.py:1188
        want, _ = oracle.pyramid_roi_align_fwd([f[k:k + 1] for f in fms], bb, None, 14, float(IMAGE * IMAGE))
        np.testing.assert_array_equal(out[k * D:(k + 1) * D], want)


# ------------------------------------------------------------------ reference-held vector on the vectorised kernels
@pytest.mark.parametrize("pool", [7, 14])
@pytest.mark.parametrize("pyr_cl", [True, False])
@pytest.mark.parametrize("algo", ["gather", "scatter"])
def test_roi_align_dropin_wide_golden(ops, pool, pyr_cl, algo):
    """golden_pyr_wide_v1.npz: the reference's own roi_align forward + backward with C = 40 (C % 4 == 0: the channel-vectorised
    kernels, one partially filled 64-channel chunk), through the drop-in, both pyramid layouts, both backward algorithms."""
    fms, boxes, shape, pools = golden_pyr_wide()
    want, grads, want_g = pools[pool]
    ts = [(cl(dev(f)) if pyr_cl else dev(f)).requires_grad_(True) for f in fms]
    ops.set_backward_algorithm(algo)
    try:
        out = ops.roi_align([dev(boxes).unsqueeze(0)] + ts, pool, shape)
        np.testing.assert_array_equal(out.detach().cpu().numpy(), want)
        out.backward(dev(grads) if not pyr_cl else cl(dev(grads)))
    finally:
        ops.set_backward_algorithm("auto")
    for l in range(4):
        assert rel_err(ts[l].grad.cpu().numpy(), want_g[l]) <= BWD_TOL


def test_detection_exchange_single_rank_matches_detection_layer(ops):
    """The fused detection layer + exchange (mrcnn_detection_layer_exchange / mrcnn_detection_collect) with one rank: the same
    detections as mrcnn_detection_layer, the mask-head RoIs of model.py:1188, and the collected copy - three exchanges in a row
    (both buffer parities, flags carried over)."""
    from maskrcnn_b200 import dist as mdist
    B, N, NC, D = 12, 1000, 81, 100
    ex = mdist.DetectionExchange(B, D)
    assert (ex.world, ex.rank, ex.begin, ex.end) == (1, 0, 0, B)
    win = dev(np.tile(np.array([[0, 0, IMAGE, IMAGE]], np.float32), (B, 1)))
    for it in range(3):
        rois = np.stack([synth.random_rois(N, 40 + 10 * it + i) for i in range(B)])
        heads = [synth.head_outputs(N, NC, 60 + 10 * it + i) for i in range(B)]
        probs, deltas = dev(np.stack([h[0] for h in heads])), dev(np.stack([h[1] for h in heads]))
        want, want_c = ops.detection_layer(dev(rois), probs, deltas, win, 0.0, 0.3, D)
        dets, counts = ex.run(dev(rois), probs, deltas, win, 0.0, 0.3, ind_offset=5, ind_mod=7)
        assert torch.equal(dets, want) and torch.equal(counts, want_c)
        assert torch.equal(ex.mask_boxes.view(B, D, 4), want[:, :, :4] / float(IMAGE))                      # model.py:1188
        assert torch.equal(ex.mask_box_ind.view(B, D), ((torch.arange(B, device="cuda") + 5) % 7).int()[:, None].expand(B, D))
        all_d, all_c = ex.collect()
        assert torch.equal(all_d, want) and torch.equal(all_c, want_c)
    assert int(ex.state[0]) == 3 and int(ex.state[1]) == 0


@pytest.mark.parametrize("variant", ["row", "tma"])
def test_fwd14_kernel_variants_are_bit_identical(variant):
    """The two measured alternatives to the default 14x14 channels-last forward kernel (MRCNN_FWD14=row: row-walking with the
    separable blend cached per feature row; =tma: producer thread + mbarrier ring of bulk row loads, bulk row stores) produce
    the default kernel's bytes at the configs[3] geometry.  The choice is fixed per process, hence the subprocesses."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, hashlib, numpy as np, torch
sys.path.insert(0, %r)
import maskrcnn_b200 as m
from maskrcnn_b200 import synth
g = torch.Generator(device="cuda"); g.manual_seed(7)
fms = [torch.randn((4, 256, s, s), device="cuda", generator=g).contiguous(memory_format=torch.channels_last) for s in (256, 128, 64, 32)]
boxes = np.concatenate([synth.random_rois(512, 50 + i) for i in range(4)])
boxes[0] = [0.0, 0.0, 1.0, 1.0]; boxes[1] = [0.2, -0.3, 0.9, 1.4]; boxes[2] = [0.5, 0.5, 0.5, 0.5]; boxes[3] = [0.9, 0.1, 0.1, 0.9]
ind = torch.arange(4, dtype=torch.int32, device="cuda").repeat_interleave(512)
out = m.pyramid_roi_align(fms, torch.from_numpy(boxes).cuda(), ind, 14, (1024, 1024, 3))
m.check_device_errors()
print(hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest())
''' % root
    digests = {}
    for v in ("col", variant):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, MRCNN_FWD14=v), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        digests[v] = r.stdout.strip().splitlines()[-1]
    assert digests["col"] == digests[variant]


def test_fused_two_head_forward_config3_and_wide_golden(ops):
    """mrcnn_pyramid_roi_align_forward_pair (both heads in one launch, what the bench step issues): at the configs[3] geometry its
    two outputs are the bytes of the two single-head launches; on the reference-recorded C = 40 pyramid (one partially filled
    channel chunk) they are the reference's own roi_align outputs."""
    fm, boxes, ind = _train_inputs(16, 512)
    ts = [cl(f) for f in fm]
    b, i = dev(boxes), dev(ind)
    o7, o14 = ops.pyramid_roi_align_pair(ts, b, i, (7, 14), (IMAGE, IMAGE, 3))
    assert torch.equal(o7, ops.pyramid_roi_align(ts, b, i, 7, (IMAGE, IMAGE, 3)))
    assert torch.equal(o14, ops.pyramid_roi_align(ts, b, i, 14, (IMAGE, IMAGE, 3)))
    o14b, o7b = ops.pyramid_roi_align_pair(ts, b, i, (14, 7), (IMAGE, IMAGE, 3))           # either order of the pool sizes
    assert torch.equal(o7b, o7) and torch.equal(o14b, o14)
    fms, gboxes, shape, pools = golden_pyr_wide()
    g7, g14 = ops.pyramid_roi_align_pair([cl(dev(f)) for f in fms], dev(gboxes), None, (7, 14), shape)
    np.testing.assert_array_equal(g7.cpu().numpy(), pools[7][0])
    np.testing.assert_array_equal(g14.cpu().numpy(), pools[14][0])
    ops.check_device_errors()


@pytest.mark.parametrize("pool", [7, 14])
@pytest.mark.parametrize("counts,C", [([40, 0, 25], 72), ([3, 130], 64), ([200], 136), ([0, 0, 9, 1], 8)])
def test_nchw_backward_image_by_image_vs_oracle(ops, pool, counts, C):
    """NCHW pyramid + NCHW gradients with the boxes grouped by image (rois_per_image: clear + scatter image by image) against the
    oracle; images without boxes, partial channel chunks, stale values in the gradient buffers."""
    size, B = 256, len(counts)
    N = sum(counts)
    fms = synth.feature_pyramid(B, C, 11 + C, image=size)
    boxes = synth.random_rois(N, 12 + N, image=float(size), min_size=6, max_size=size * 0.9)
    ind = np.repeat(np.arange(B, dtype=np.int32), counts)
    ts = [dev(f).requires_grad_(True) for f in fms]
    for t in ts:   # stale values: the kernel must clear every slice itself, also those of images without boxes
        t.grad = torch.full_like(t, 3.0)
    ops.set_backward_algorithm("scatter")
    try:
        out = ops.pyramid_roi_align(ts, dev(boxes), None, pool, (size, size, 3), rois_per_image=counts)
        want, _ = oracle.pyramid_roi_align_fwd(fms, boxes, ind, pool, float(size * size))
        np.testing.assert_array_equal(out.detach().cpu().numpy(), want)
        g = np.random.default_rng(pool + C).standard_normal(want.shape, dtype=np.float32)
        grads = torch.autograd.grad(out, ts, dev(g))
    finally:
        ops.set_backward_algorithm("auto")
    want_g = oracle.pyramid_roi_align_bwd(g, [f.shape for f in fms], boxes, ind, float(size * size))
    for got, w in zip(grads, want_g):
        assert rel_err(got.cpu().numpy(), w) <= BWD_TOL
    ops.check_device_errors()


@pytest.mark.parametrize("fused", [False, True])
def test_host_train_step_matches_device_ops(ops, fused):
    """maskrcnn_b200.hoststep.HostTrainStep (pinned host tensors in and out, per-image three-stream pipeline: what bench.py's e2e leg
    times) against the oracle: crops bit-exact, gradient pyramids <= 1e-5 (their sum when the two heads' backward is fused), mask
    targets bit-exact."""
    from maskrcnn_b200.hoststep import HostTrainStep
    B, R, C, size, P, G = 3, 40, 72, 256, 6, 4
    level_hw = [(size // s, size // s) for s in (4, 8, 16, 32)]
    hs = HostTrainStep(B, R, C, level_hw, (size, size), fused_backward=fused)
    rng = np.random.default_rng(3)
    fms = synth.feature_pyramid(B, C, 21, image=size)
    boxes = np.concatenate([synth.random_rois(R, 30 + i, image=float(size), min_size=6, max_size=size * 0.9) for i in range(B)])
    ind = np.repeat(np.arange(B, dtype=np.int32), R)
    g7 = rng.standard_normal((B * R, C, 7, 7), dtype=np.float32)
    g14 = rng.standard_normal((B * R, C, 14, 14), dtype=np.float32)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    nhwc = lambda a: pin(a.transpose(0, 2, 3, 1))  # noqa: E731
    h_out7, h_out14 = hs.pinned_like(hs.out7), hs.pinned_like(hs.out14)
    h_ga = [hs.pinned_like(f) for f in hs.fm]
    h_gb = None if fused else [hs.pinned_like(f) for f in hs.fm]
    gt = np.zeros((B * G, 1, size, size), np.float32)
    for k in range(B * G):
        y, x = rng.integers(0, size - 40, 2)
        gt[k, 0, y:y + 40, x:x + 60] = 1.0
    mboxes = np.concatenate([synth.random_rois(P, 70 + i, image=float(size), min_size=6, max_size=size * 0.9) for i in range(B)])
    mind = (np.repeat(np.arange(B), P) * G + rng.integers(0, G, B * P)).astype(np.int32)
    mask = {"d_images": dev(gt), "h_boxes": pin(mboxes), "h_index": pin(mind), "d_boxes": torch.empty((B * P, 4), device="cuda"),
            "d_index": torch.empty(B * P, dtype=torch.int32, device="cuda"), "d_targets": torch.empty((B * P, 1, 28, 28), device="cuda"),
            "h_targets": torch.empty((B * P, 1, 28, 28)).pin_memory()}
    for _ in range(2):      # twice: buffers and workspaces are reused
        hs.run([nhwc(f) for f in fms], pin(boxes), nhwc(g7), nhwc(g14), h_out7, h_out14, h_ga, h_gb, mask)
    area = float(size * size)
    w7, _ = oracle.pyramid_roi_align_fwd(fms, boxes, ind, 7, area)
    w14, _ = oracle.pyramid_roi_align_fwd(fms, boxes, ind, 14, area)
    np.testing.assert_array_equal(h_out7.numpy().transpose(0, 3, 1, 2), w7)
    np.testing.assert_array_equal(h_out14.numpy().transpose(0, 3, 1, 2), w14)
    shapes = [f.shape for f in fms]
    b7 = oracle.pyramid_roi_align_bwd(g7, shapes, boxes, ind, area)
    b14 = oracle.pyramid_roi_align_bwd(g14, shapes, boxes, ind, area)
    for l in range(4):
        if fused:
            assert rel_err(h_ga[l].numpy().transpose(0, 3, 1, 2), b7[l] + b14[l]) <= BWD_TOL
        else:
            assert rel_err(h_ga[l].numpy().transpose(0, 3, 1, 2), b14[l]) <= BWD_TOL
            assert rel_err(h_gb[l].numpy().transpose(0, 3, 1, 2), b7[l]) <= BWD_TOL
    np.testing.assert_array_equal(mask["h_targets"].numpy(), oracle.crop_forward(gt, mboxes, mind, 28, 28, 0.0))
    assert hs.h2d_bytes(mask) > 0 and hs.d2h_bytes(mask) > 0
    ops.check_device_errors()
