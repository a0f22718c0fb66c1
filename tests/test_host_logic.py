"""Host-side logic of maskrcnn_b200.ops without a GPU: the library is replaced by a recorder (and, where the host code
reads a result back, by a stub that writes it), the CUDA-only guards by dtype checks, and the tensors live on the CPU.
What is checked is what the host layer decides on its own: memory-layout selection, the arguments marshalled into the
C ABI, the reference's Python arithmetic restated on the host (decode_masks' crop origin / target size, the negatives
count and the two permutation draws of mrn_samples), un-padding of the padded results and the error paths.
The kernels themselves are covered by tests/test_gpu_parity.py."""
import contextlib
import os
import ctypes
import types

import numpy as np
import pytest
import torch

from maskrcnn_b200 import _lib, ops


class FakeLib:
    """Records (name, args) of every C-ABI call; `on[name]` may hold a stub that produces the call's side effects."""

    def __init__(self):
        self.calls = []
        self.on = {}

    def __getattr__(self, name):
        if not name.startswith("mrcnn_"):
            raise AttributeError(name)
        assert name in _lib.SIGNATURES, "ops.py calls %s, which include/mrcnn_b200.h does not declare" % name

        def call(*args):
            assert len(args) == len(_lib.SIGNATURES[name][1]), "%s: argument count differs from its signature" % name
            self.calls.append((name, args))
            if name in self.on:
                return self.on[name](*args)
            return 64 if name.endswith("workspace_bytes") else 0
        return call

    def named(self, name):
        return [a for n, a in self.calls if n == name]


def _write_i32(ptr, values):
    arr = (ctypes.c_int32 * len(values))(*[int(v) for v in values])
    ctypes.memmove(ptr, arr, 4 * len(values))


@pytest.fixture
def fake(monkeypatch):
    lib = FakeLib()

    def require(t, name, dtype=None):
        if not isinstance(t, torch.Tensor):
            raise TypeError(name)
        if dtype is not None and t.dtype != dtype:
            raise TypeError("%s must have dtype %s" % (name, dtype))
        return t
    monkeypatch.setattr(ops, "lib", lib)
    monkeypatch.setattr(ops, "check", lambda rc: None)
    monkeypatch.setattr(ops, "_require_cuda", require)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    return lib


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


# ------------------------------------------------------------------ layouts
def test_layout_selection():
    x = torch.zeros(2, 8, 5, 6)
    assert ops._layout4(x)[1] == _lib.NCHW and ops._layout4(x)[0] is x
    assert ops._layout4(cl(x))[1] == _lib.NHWC
    # [N,C,1,1] is dense in both orders: channels-last when the vector kernels can take it (C % 4 == 0)
    assert ops._layout4(torch.zeros(3, 8, 1, 1))[1] == _lib.NHWC
    assert ops._layout4(torch.zeros(3, 6, 1, 1))[1] == _lib.NCHW
    # neither order (a slice along W): made dense NCHW, like the reference's .contiguous() meant to (crop_cpu.cpp:129-131)
    y, lay = ops._layout4(torch.zeros(2, 8, 5, 12)[..., ::2])
    assert lay == _lib.NCHW and y.is_contiguous()
    with pytest.raises(ValueError):
        ops._layout4(torch.zeros(2, 3, 4))
    # a pyramid with mixed layouts settles on channels-last
    fms = [torch.zeros(1, 8, s, s) for s in (16, 8, 4, 2)]
    outs, lay = ops._pyramid_layout([cl(fms[0])] + fms[1:])
    assert lay == _lib.NHWC and all(o.is_contiguous(memory_format=torch.channels_last) for o in outs)
    assert ops._pyramid_layout(fms)[1] == _lib.NCHW
    e = ops._empty4((3, 8, 7, 7), _lib.NHWC, fms[0])
    assert e.is_contiguous(memory_format=torch.channels_last) and e.shape == (3, 8, 7, 7)


# ------------------------------------------------------------------ nms / crop
def test_nms_unpads_with_the_count_the_library_wrote(fake):
    fake.on["mrcnn_nms"] = lambda dets, n, thr, keep, count, ws, wsb, st: (_write_i32(count, [3]), 0)[1]
    dets = torch.rand(10, 5)
    keep = ops.nms(dets, 0.7)
    assert keep.dtype == torch.int64 and keep.shape == (3,)
    (args,) = fake.named("mrcnn_nms")
    assert args[0] == dets.data_ptr() and args[1] == 10 and args[2] == pytest.approx(0.7) and args[6] == 64
    assert fake.named("mrcnn_nms_workspace_bytes") == [(10,)]
    # nms.h:20-21: empty in -> empty out, the library is not called
    fake.calls.clear()
    assert ops.nms(torch.zeros(0, 5), 0.5).numel() == 0 and not fake.calls
    with pytest.raises(ValueError):
        ops.nms(torch.zeros(4, 4), 0.5)
    with pytest.raises(TypeError):
        ops.nms(torch.zeros(4, 5, dtype=torch.float64), 0.5)


def test_crop_function_mirrors_the_reference_object(fake):
    f = ops.CropFunction(7, 9, 0.5)                                      # __init__.py:27-30
    assert (f.crop_height, f.crop_width, f.extrapolation_value) == (7, 9, 0.5)
    assert ops.CropFunction(3, 3).extrapolation_value == 0
    image = torch.zeros(2, 4, 16, 16, requires_grad=True)
    boxes = torch.tensor([[0., 0., 1., 1.], [0.2, 0.2, 0.8, 0.8], [0., 0., .5, .5]])
    ind = torch.tensor([0, 1, 1], dtype=torch.int32)
    out = f(image, boxes, ind)
    assert out.shape == (3, 4, 7, 9) and out.is_contiguous()
    (a,) = fake.named("mrcnn_crop_forward")
    assert a[1:6] == (2, 4, 16, 16, _lib.NCHW) and a[8] == 3 and a[9] == 0.5 and a[10:12] == (7, 9) and a[13] == _lib.NCHW
    out.backward(torch.ones_like(out))
    (b,) = fake.named("mrcnn_crop_backward")
    assert b[4:7] == (3, 7, 9) and b[8:12] == (2, 4, 16, 16) and b[13] == 1       # zero_fill: the op owns the clear
    assert image.grad.shape == image.shape
    # channels-last image -> channels-last crops and gradient
    fake.calls.clear()
    img2 = cl(torch.zeros(2, 4, 16, 16)).requires_grad_()
    out2 = f.forward(img2, boxes, ind)
    assert out2.is_contiguous(memory_format=torch.channels_last)
    assert fake.named("mrcnn_crop_forward")[0][5] == _lib.NHWC
    # no boxes: an empty result without a launch
    fake.calls.clear()
    assert f(image, boxes[:0], ind[:0]).shape == (0, 4, 7, 9) and not fake.named("mrcnn_crop_forward")
    # nothing to differentiate: the same call without an autograd node
    fake.calls.clear()
    plain = f(image.detach(), boxes, ind)
    with torch.no_grad():
        quiet = f(image, boxes, ind)
    assert plain.grad_fn is None and quiet.grad_fn is None and f(image, boxes, ind).grad_fn is not None
    x, y, z = fake.named("mrcnn_crop_forward")
    assert (x[:12], x[13:]) == (y[:12], y[13:]) == (z[:12], z[13:])      # everything but the output pointer
    with pytest.raises(TypeError):
        f(image, boxes, ind.long())                                       # __init__.py:34-35: box_ind is int32
    with pytest.raises(ValueError):
        f(image, boxes, ind[:2])


# ------------------------------------------------------------------ PyramidROIAlign
def _pyramid(b=2, c=8):
    return [torch.zeros(b, c, s, s) for s in (16, 8, 4, 2)]


def test_pyramid_roi_align_marshalling(fake):
    fms = _pyramid()
    boxes = torch.rand(5, 4)
    ind = torch.tensor([0, 0, 1, 1, 1], dtype=torch.int32)
    out = ops.pyramid_roi_align(fms, boxes, ind, 7, (64, 32, 3))
    assert out.shape == (5, 8, 7, 7)
    (a,) = fake.named("mrcnn_pyramid_roi_align_forward")
    assert list(a[0]) == [f.data_ptr() for f in fms] and list(a[1]) == [16, 8, 4, 2] and list(a[2]) == [16, 8, 4, 2]
    assert a[3:6] == (2, 8, _lib.NCHW) and a[8:10] == (5, 7) and a[10] == 64.0 * 32.0 and a[12] == _lib.NCHW   # model.py:331
    # channels-last pyramid -> channels-last crops, unless told otherwise
    fake.calls.clear()
    assert ops.pyramid_roi_align([cl(f) for f in fms], boxes, ind, 14, (64, 64, 3)).is_contiguous(memory_format=torch.channels_last)
    assert ops.pyramid_roi_align([cl(f) for f in fms], boxes, ind, 14, (64, 64, 3), out_channels_last=False).is_contiguous()
    assert [a[12] for a in fake.named("mrcnn_pyramid_roi_align_forward")] == [_lib.NHWC, _lib.NCHW]
    # rois_per_image: counts -> box_ind (grouped by image) and validated against the boxes
    fake.calls.clear()
    ops.pyramid_roi_align(fms, boxes, None, 7, (64, 64, 3), rois_per_image=[2, 3])
    with pytest.raises(ValueError):
        ops.pyramid_roi_align(fms, boxes, None, 7, (64, 64, 3), rois_per_image=[2, 2])
    with pytest.raises(ValueError):
        ops.pyramid_roi_align(fms, boxes, None, 7, (64, 64, 3), rois_per_image=[5])
    with pytest.raises(ValueError):      # one source of truth for the image of a box: the counts OR box_ind, never both
        ops.pyramid_roi_align(fms, boxes, ind, 7, (64, 64, 3), rois_per_image=[2, 3])
    with pytest.raises(ValueError):
        ops.pyramid_roi_align(fms, boxes, ind[:4], 7, (64, 64, 3))
    with pytest.raises(ValueError):
        ops.pyramid_roi_align(fms[:3] + [torch.zeros(2, 4, 2, 2)], boxes, ind, 7, (64, 64, 3))
    with pytest.raises(ValueError):
        ops.pyramid_roi_align(fms[:3], boxes, ind, 7, (64, 64, 3))


def test_roi_align_dropin_is_batch_one_like_the_reference(fake):
    fms = _pyramid(b=1)
    boxes = torch.rand(1, 6, 4)
    assert ops.roi_align([boxes] + fms, 7, [64, 64, 3]).shape == (6, 8, 7, 7)            # model.py:312-313 squeeze(0)
    a = fake.named("mrcnn_pyramid_roi_align_forward")[0]
    assert a[3] == 1 and a[7] is None and a[8] == 6                                      # box_ind = zeros (model.py:369)
    with pytest.raises(ValueError):
        ops.roi_align([torch.rand(2, 6, 4)] + fms, 7, [64, 64, 3])


def test_backward_algorithm_knob():
    for name in ("gather", "scatter", "auto"):
        ops.set_backward_algorithm(name)
        assert ops.BACKWARD_ALGORITHM == name
    with pytest.raises(ValueError):
        ops.set_backward_algorithm("atomic")


# ------------------------------------------------------------------ proposal / detection layers
def _model_self(**cfg):
    base = dict(IMAGE_SHAPE=[1024, 1024, 3], RPN_NMS_MAX_ROIS_NUM=1000, RPN_NMS_THRESHOLD=0.7, RPN_BBOX_STD_DEV=[0.1, 0.1, 0.2, 0.2],
                DETECTION_MIN_CONFIDENCE=0, DETECTION_NMS_THRESHOLD=0.3, DETECTION_MAX_INSTANCES=50)
    base.update(cfg)
    return types.SimpleNamespace(config=types.SimpleNamespace(**base), anchors=torch.zeros(2000, 4))


def test_rpn_refine_reads_the_config_and_unpads(fake):
    def prop(cls, bbox, anc, B, A, pre, post, thr, std, h, w, rois, counts, ws, wsb, st):
        _write_i32(counts, [37])
        return 0
    fake.on["mrcnn_proposal_layer"] = prop
    me = _model_self()
    rois = ops.rpn_refine(me, torch.zeros(1, 2000, 2), torch.zeros(1, 2000, 4))
    assert rois.shape == (1, 37, 4)                                                      # not padded, model.py:1366-1381
    (a,) = fake.named("mrcnn_proposal_layer")
    assert a[3:7] == (1, 2000, 500, 1000) and a[7] == pytest.approx(0.7)                 # pre-NMS 500: model.py:1345
    assert [round(v, 6) for v in a[8]] == [0.1, 0.1, 0.2, 0.2] and a[9:11] == (1024.0, 1024.0)
    assert fake.named("mrcnn_proposal_workspace_bytes") == [(1, 2000, 500)]
    # fewer anchors than the limit; explicit pre-NMS limit; fg-only scores take the _fg entry point
    fake.calls.clear()
    me.anchors = torch.zeros(300, 4)
    ops.rpn_refine(me, torch.zeros(1, 300, 2), torch.zeros(1, 300, 4))
    assert fake.named("mrcnn_proposal_layer")[0][5] == 300
    fake.calls.clear()
    ops.proposal_layer(torch.zeros(2, 300), torch.zeros(2, 300, 4), torch.zeros(300, 4), 100, 50, 0.7)
    assert fake.named("mrcnn_proposal_layer_fg")[0][3:7] == (2, 300, 100, 50) and not fake.named("mrcnn_proposal_layer")
    with pytest.raises(ValueError):
        ops.proposal_layer(torch.zeros(1, 300, 2), torch.zeros(1, 300, 4), torch.zeros(299, 4), 100, 50, 0.7)
    with pytest.raises(ValueError):
        ops.proposal_layer(torch.zeros(1, 300, 3), torch.zeros(1, 300, 4), torch.zeros(300, 4), 100, 50, 0.7)


def test_mrn_refine_none_when_nothing_survives_and_split_otherwise(fake):
    state = {"count": 0}

    def det(rois, probs, deltas, win, B, N, NC, mc, thr, D, std, h, w, dets, counts, index, ws, wsb, st):
        _write_i32(counts, [state["count"]])
        vals = np.zeros((B, D, 6), np.float32)
        vals[0, :2] = [[1, 2, 30, 40, 0.9, 17], [5, 6, 70, 80, 0.8, 3]]
        ctypes.memmove(dets, vals.ctypes.data, vals.nbytes)
        return 0
    fake.on["mrcnn_detection_layer"] = det
    me = _model_self()
    args = (torch.zeros(1, 100, 4), torch.zeros(100, 81), torch.zeros(100, 81, 4), np.array([0, 0, 1024, 1024]))
    assert ops.mrn_refine(me, *args) == (None, None, None)                               # model.py:1445-1447
    state["count"] = 2
    ci, sc, bx = ops.mrn_refine(me, *args)
    assert ci.dtype == torch.int64 and ci.tolist() == [[17, 3]]
    assert sc.shape == (1, 2) and bx.tolist() == [[[1, 2, 30, 40], [5, 6, 70, 80]]]
    a = fake.named("mrcnn_detection_layer")[-1]
    assert a[4:7] == (1, 100, 81) and a[7] == 0.0 and a[8] == pytest.approx(0.3) and a[9] == 50 and a[15] is None
    with pytest.raises(ValueError):
        ops.detection_layer(torch.zeros(1, 100, 4), torch.zeros(1, 100, 81), torch.zeros(1, 99, 81, 4), torch.zeros(1, 4), 0, 0.3, 50)


# ------------------------------------------------------------------ detection targets / anchor matching
def test_negatives_count_is_the_reference_python_arithmetic():
    for ratio in (0.33, 0.25, 0.5, 1.0 / 3.0):
        for p in range(0, 200):
            r = 1.0 / ratio                                                              # model.py:518-519
            assert ops._negatives_for(p, ratio) == int(r * p - p)
    assert ops._negatives_for(168, 0.33) == 341


def test_mrn_samples_draws_in_the_reference_order(fake, monkeypatch):
    """torch.randperm(P)[:cap] first, torch.randperm(Q)[:negatives] second (model.py:468, :520); padded into [1,N] rows."""
    N, G, P, Q = 40, 3, 9, 20

    def classify(rois, gtb, cls, B, n, g, pos, neg, assign, iou, counts, st):
        _write_i32(counts, [P, Q])
        return 0
    fake.on["mrcnn_target_classify"] = classify
    draws = []
    real = torch.randperm

    def randperm(n):
        draws.append(n)
        return real(n)
    monkeypatch.setattr(torch, "randperm", randperm)
    seen = {}

    def emit(*a):
        # a[11], a[12] = perm_pos / perm_neg, a[13] = take, a[18] = T
        seen["pp"] = np.ctypeslib.as_array(ctypes.cast(a[11], ctypes.POINTER(ctypes.c_int32)), (N,)).copy()
        seen["pn"] = np.ctypeslib.as_array(ctypes.cast(a[12], ctypes.POINTER(ctypes.c_int32)), (N,)).copy()
        seen["take"] = np.ctypeslib.as_array(ctypes.cast(a[13], ctypes.POINTER(ctypes.c_int32)), (2,)).copy()
        seen["T"] = a[18]
        return 0
    fake.on["mrcnn_target_emit"] = emit
    cfg = types.SimpleNamespace(ROI_POSITIVE_RATIO=0.33, TRAIN_ROIS_PER_IMAGE=12, BBOX_STD_DEV=[0.1, 0.1, 0.2, 0.2], MASK_SHAPE=[28, 28])
    inputs = (torch.rand(1, N, 4), torch.ones(1, G, dtype=torch.int32), torch.rand(1, G, 4), torch.zeros(1, G, 32, 32))
    torch.manual_seed(5)
    rois, cls, deltas, masks = ops.mrn_samples(*inputs, cfg)
    cap = int(12 * 0.33)                                                                 # 3 positives, model.py:466-467
    neg = ops._negatives_for(cap, 0.33)
    assert draws == [P, Q] and seen["take"].tolist() == [cap, neg] and seen["T"] == cap + neg
    torch.manual_seed(5)
    want_p, want_n = real(P)[:cap], real(Q)[:neg]
    assert seen["pp"][:cap].tolist() == want_p.tolist() and seen["pn"][:neg].tolist() == want_n.tolist()
    assert rois.shape == (cap + neg, 4) and cls.shape == (cap + neg,) and masks.shape == (cap + neg, 28, 28)
    # no positives -> the reference's empty results (model.py:563-574), nothing drawn
    draws.clear()
    P = 0
    out = ops.mrn_samples(torch.rand(1, N, 4), torch.ones(1, G, dtype=torch.int32), torch.rand(1, G, 4), torch.zeros(1, G, 32, 32), cfg)
    assert [t.numel() for t in out] == [0, 0, 0, 0] and not draws
    with pytest.raises(ValueError):
        ops.mrn_samples(torch.rand(2, N, 4), torch.ones(2, G, dtype=torch.int32), torch.rand(2, G, 4), torch.zeros(2, G, 32, 32), cfg)


# ------------------------------------------------------------------ masks
@pytest.mark.parametrize("H,W,ch,cw,scale", [(1024, 1024, 640, 1024, 0.5333333), (1024, 1024, 1024, 683, 1.7066), (256, 256, 160, 255, 0.2),
                                             (65, 63, 64, 33, 3.1), (128, 128, 128, 128, 0.75), (101, 99, 100, 98, 0.999)])
def test_decode_masks_geometry_is_torchvisions_and_the_references(fake, H, W, ch, cw, scale):
    """CenterCrop's origin (torchvision: int(round((H - h) / 2.))) and data.py:276-277's target size, against torchvision
    itself where it is installed."""
    masks = torch.zeros(3, H, W, dtype=torch.bool)
    out = ops.decode_masks(masks, scale, types.SimpleNamespace(height=lambda: ch, width=lambda: cw))
    (a,) = fake.named("mrcnn_decode_masks")
    is_bool, D, h, w, top, left, c_h, c_w, nh, nw = a[1:11]
    assert (is_bool, D, h, w, c_h, c_w) == (1, 3, H, W, ch, cw)
    assert (nh, nw) == (int(round(ch * 1.0 / scale)), int(round(cw * 1.0 / scale))) and out.shape == (3, nh, nw) and out.dtype == torch.uint8
    tv = pytest.importorskip("torchvision.transforms.functional")
    probe = torch.arange(H * W, dtype=torch.int32).reshape(1, H, W)
    crop = tv.center_crop(probe, [ch, cw])
    assert int(crop[0, 0, 0]) == top * W + left
    # the (height, width) pair form and uint8 masks
    fake.calls.clear()
    ops.decode_masks(masks.to(torch.uint8), scale, (ch, cw))
    assert fake.named("mrcnn_decode_masks")[0][1] == 0 and fake.named("mrcnn_decode_masks")[0][5:9] == (top, left, ch, cw)


def test_decode_and_full_masks_argument_errors(fake):
    m = torch.zeros(2, 32, 32, dtype=torch.bool)
    assert ops.decode_masks(m, 1, (32, 32)) is m                                         # data.py:267-268
    with pytest.raises(ValueError):
        ops.decode_masks(m, 0.5, (33, 32))                                               # CenterCrop would pad
    with pytest.raises(ValueError):
        ops.decode_masks(m.float(), 0.5, (32, 32))
    with pytest.raises(ValueError):
        ops.decode_masks(m, 1000.0, (32, 32))                                            # rounds to a 0-pixel frame: PIL's error
    out = ops.full_masks(torch.zeros(2, 3, dtype=torch.int64), torch.zeros(2, 3, 4), torch.zeros(2, 3, 81, 28, 28), 40, 48)
    assert out.shape == (2, 3, 40, 48) and out.dtype == torch.bool                       # leading batch dimensions kept
    a = fake.named("mrcnn_full_masks")[0]
    assert a[3:9] == (6, 81, 28, 28, 40, 48)
    with pytest.raises(ValueError):
        ops.full_masks(torch.zeros(3, dtype=torch.int64), torch.zeros(2, 4), torch.zeros(3, 81, 28, 28), 40, 48)


# ------------------------------------------------------------------ RPN head plumbing
def test_rpn_pack_geometry_and_argument_errors(fake):
    K = 3
    sizes = (8, 4, 2)
    logits = [torch.zeros(2, 2 * K, s, s) for s in sizes]
    bboxes = [torch.zeros(2, 4 * K, s, s) for s in sizes]
    lg, cls, bb, fg = ops.rpn_pack(logits, bboxes)
    A = K * sum(s * s for s in sizes)                                                    # model.py:1294-1304: cat over levels
    assert lg.shape == (2, A, 2) and cls.shape == (2, A, 2) and bb.shape == (2, A, 4) and fg.shape == (2, A)
    a = fake.named("mrcnn_rpn_pack")[0]
    assert a[4:8] == (3, 2, K, _lib.NCHW)
    assert ops.rpn_pack(logits, bboxes, want_class=False)[1] is None
    assert fake.named("mrcnn_rpn_pack")[1][9] is None
    assert ops.rpn_pack([cl(t) for t in logits], [cl(t) for t in bboxes])[0].shape == (2, A, 2)
    assert fake.named("mrcnn_rpn_pack")[2][7] == _lib.NHWC
    with pytest.raises(ValueError):
        ops.rpn_pack(logits, bboxes[:2])
    with pytest.raises(ValueError):
        ops.rpn_pack([torch.zeros(2, 5, 8, 8)], [torch.zeros(2, 10, 8, 8)])             # odd class channels
    with pytest.raises(ValueError):
        ops.rpn_pack(logits, [torch.zeros(2, 4 * K, s, s + 1) for s in sizes])


def test_pyramid_roi_align_skips_the_autograd_node_when_nothing_needs_a_gradient(fake):
    """Inference issues a kernel shorter than an autograd.Function.apply: without a feature map that requires grad (or under
    torch.no_grad) the same launch is made directly.  Both routes marshal the same arguments."""
    boxes = torch.rand(5, 4)
    ind = torch.tensor([0, 0, 1, 1, 1], dtype=torch.int32)
    plain = _pyramid()
    out = ops.pyramid_roi_align(plain, boxes, ind, 7, (64, 48, 3))
    assert out.grad_fn is None and out.shape == (5, 8, 7, 7)
    needs = [f.clone().requires_grad_() for f in plain]
    out_g = ops.pyramid_roi_align(needs, boxes, ind, 7, (64, 48, 3))
    assert out_g.grad_fn is not None
    with torch.no_grad():
        out_n = ops.pyramid_roi_align(needs, boxes, ind, 7, (64, 48, 3))
    assert out_n.grad_fn is None
    a, b, c = fake.named("mrcnn_pyramid_roi_align_forward")

    def scalars(args):   # everything but the pointers
        return ([list(args[1]), list(args[2])], args[3:6], args[8:11], args[12:14])
    assert scalars(a) == scalars(b) == scalars(c)
    assert list(a[0]) == [f.data_ptr() for f in plain] and list(b[0]) == list(c[0]) == [f.data_ptr() for f in needs]
    assert a[6] == b[6] == c[6] and a[7] == b[7] == c[7] == ind.data_ptr()
    # the gradient route still reaches the backward entry point
    out_g.backward(torch.ones_like(out_g))
    assert len(fake.named("mrcnn_pyramid_roi_align_backward")) == 1 and all(f.grad is not None for f in needs)
    # channels-last / explicit output layout / no boxes behave the same on the direct route
    fake.calls.clear()
    cl_fms = [cl(f) for f in plain]
    assert ops.pyramid_roi_align(cl_fms, boxes, ind, 14, (64, 64, 3)).is_contiguous(memory_format=torch.channels_last)
    assert ops.pyramid_roi_align(cl_fms, boxes, ind, 14, (64, 64, 3), out_channels_last=False).is_contiguous()
    assert [x[12] for x in fake.named("mrcnn_pyramid_roi_align_forward")] == [_lib.NHWC, _lib.NCHW]
    fake.calls.clear()
    assert ops.pyramid_roi_align(plain, boxes[:0], None, 7, (64, 64, 3)).shape == (0, 8, 7, 7) and not fake.calls


def test_roofline_unique_taps_single_and_fused_heads():
    """roofline.unique_taps: the positions one head reads, and - for a fused two-head forward - the positions EITHER head reads,
    each counted once: at least what the larger head reads alone, at most the sum; a single RoI that covers one level entirely
    touches every position of its footprint rows / columns."""
    import numpy as np
    from maskrcnn_b200 import roofline, synth
    level_hw = [(64, 64), (32, 32), (16, 16), (8, 8)]
    boxes = np.concatenate([synth.random_rois(40, 70 + i) for i in range(2)], 0)
    ind = np.repeat(np.arange(2), 40)
    u7, lv = roofline.unique_taps(boxes, ind, 7, (256, 256), level_hw, 2)
    u14, _ = roofline.unique_taps(boxes, ind, 14, (256, 256), level_hw, 2)
    both, _ = roofline.unique_taps(boxes, ind, (7, 14), (256, 256), level_hw, 2)
    assert 0 < u7 <= u14 <= both <= u7 + u14
    assert both < u7 + u14                                   # the heads share most of their footprint
    assert set(np.unique(lv)) <= {2, 3, 4, 5}
    whole = np.array([[0.0, 0.0, 1.0, 1.0]], np.float32)     # the whole 256 x 256 image: level 4 (16 x 16); 7 taps per axis at i * 15 / 6
    # -> 0, 2.5, 5, 7.5, 10, 12.5, 15: ten distinct floor / ceil positions per axis
    u, lvw = roofline.unique_taps(whole, None, 7, (256, 256), level_hw, 1)
    assert lvw[0] == 4 and u == 100


def test_algorithm_knobs_reach_the_library(fake):
    """set_proposal_nms / set_detection_nms / set_deterministic: names map onto the header's constants, unknown names raise before
    anything reaches the library."""
    for name, code in (("auto", 0), ("mask", 1), ("lazy", 2), ("hybrid", 3)):
        ops.set_proposal_nms(name)
        assert fake.named("mrcnn_set_proposal_nms")[-1] == (code,)
    for name, code in (("auto", 0), ("mask", 1), ("lazy", 2)):
        ops.set_detection_nms(name)
        assert fake.named("mrcnn_set_detection_nms")[-1] == (code,)
    n = len(fake.calls)
    with pytest.raises(ValueError):
        ops.set_proposal_nms("fixpoint")
    with pytest.raises(ValueError):
        ops.set_detection_nms("hybrid")
    assert len(fake.calls) == n
    ops.set_deterministic(True)
    ops.set_deterministic(0)
    assert fake.named("mrcnn_set_deterministic") == [(1,), (0,)]
    # the header's constants are the ones ops.py sends
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "mrcnn_b200.h")).read()
    for macro, code in (("MRCNN_PROPOSAL_NMS_AUTO", 0), ("MRCNN_PROPOSAL_NMS_MASK", 1), ("MRCNN_PROPOSAL_NMS_LAZY", 2), ("MRCNN_PROPOSAL_NMS_HYBRID", 3)):
        assert "#define %s %d" % (macro, code) in hdr
