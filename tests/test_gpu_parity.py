"""GPU parity tests: the CUDA path (through the C ABI via maskrcnn_b200.ops) against the plain-C oracle
and the committed golden vectors.  Bit-exact for indices/selections and for the forward interpolation;
<= 1e-5 relative (scale = max |reference|) for the atomically accumulated backward."""
import numpy as np
import pytest
import torch

import oracle
from helpers import golden, rel_err
from maskrcnn_b200 import synth

pytestmark = pytest.mark.gpu

BWD_TOL = 1e-5  # north_star: RoIAlign forward/backward within 1e-5 relative in fp32


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


# ------------------------------------------------------------------ nms
def _dets(n, seed, thr_cluster=True):
    rng = np.random.default_rng(seed)
    b = synth.random_rois(n, seed, image=1024.0, min_size=16, max_size=500) * 1024.0
    if thr_cluster and n > 3:
        h = n // 2
        b[h:] = b[: n - h] + rng.uniform(-8, 8, (n - h, 4)).astype(np.float32)
    return np.concatenate([b, synth.unique_scores(n, seed)[:, None]], 1).astype(np.float32)


@pytest.mark.parametrize("n,thr", [(1, 0.5), (2, 0.5), (63, 0.3), (64, 0.7), (65, 0.5), (500, 0.7), (1000, 0.3),
                                   (6000, 0.7), (8192, 0.5), (8193, 0.7), (20000, 0.6)])
def test_nms_matches_oracle(ops, n, thr):
    dets = _dets(n, 1000 + n)
    want = oracle.nms(dets, thr)
    got = ops.nms(dev(dets), thr)
    assert got.dtype == torch.int64 and got.is_cuda
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_nms_edge_cases(ops):
    assert ops.nms(torch.zeros(0, 5, device="cuda"), 0.5).numel() == 0
    dets = np.array([[0, 0, 9, 9, 0.9], [0, 0, 9, 9, 0.8], [0, 0, 9, 4, 0.7], [5, 5, 2, 2, 0.6], [5, 5, 2, 2, 0.5],
                     [0, 0, -1, -1, 0.4], [0, 0, -1, -1, 0.3], [100, 100, 120, 130, 0.2]], np.float32)
    for thr in (0.5, 0.3, 0.0, 1.0):
        np.testing.assert_array_equal(ops.nms(dev(dets), thr).cpu().numpy(), oracle.nms(dets, thr))
    # all suppressed but the first / none suppressed
    same = np.tile(np.array([[10, 10, 50, 50]], np.float32), (200, 1))
    d = np.concatenate([same, synth.unique_scores(200, 3)[:, None]], 1)
    np.testing.assert_array_equal(ops.nms(dev(d), 0.5).cpu().numpy(), oracle.nms(d, 0.5))
    far = np.stack([np.arange(300) * 100.0, np.zeros(300), np.arange(300) * 100.0 + 10, np.full(300, 10.0)], 1).astype(np.float32)
    d = np.concatenate([far, synth.unique_scores(300, 4)[:, None]], 1)
    np.testing.assert_array_equal(ops.nms(dev(d), 0.5).cpu().numpy(), np.arange(300))


def _chain(n, step=20.0, side=100.0):
    """Boxes in a row, each overlapping only its neighbours above the threshold (IoU 0.67 with the next, 0.43 with the one after),
    scores descending along the row: the survivors alternate, and box i's fate hangs on box i - 1's - the longest dependency chain
    there is (the worst case for the grid-wide fixed-point route: one pass per chunk)."""
    x = np.arange(n, dtype=np.float32) * step
    b = np.stack([np.zeros(n, np.float32), x, np.full(n, side, np.float32), x + side], 1)
    return np.concatenate([b, np.linspace(0.99, 0.01, n, dtype=np.float32)[:, None]], 1).astype(np.float32)


@pytest.mark.parametrize("n", [449, 1000, 6000, 10000])
def test_nms_longest_dependency_chain(ops, n):
    dets = _chain(n)
    want = oracle.nms(dets, 0.5)
    assert len(want) > n // 3
    np.testing.assert_array_equal(ops.nms(dev(dets), 0.5).cpu().numpy(), want)
    perm = np.random.default_rng(n).permutation(n)          # the same chain with the boxes stored in a random order
    np.testing.assert_array_equal(ops.nms(dev(dets[perm]), 0.5).cpu().numpy(), oracle.nms(dets[perm], 0.5))


def test_nms_edge_cases_fixed_point_route(ops):
    """The edge cases of test_nms_edge_cases at a size that takes the grid-wide fixed-point route (1500 boxes): identical boxes,
    disjoint boxes, thresholds 0, 1, negative and NaN, degenerate (inverted, zero-area, duplicated) boxes, huge coordinates, NaN
    coordinates - the division-free "certainly not" bound must hand every such pair to the reference's expression.  (The oracle is
    pinned against the reference's compiled nms on the same inputs: tests/test_oracle_vs_ref.py.)"""
    import helpers
    for name, d in helpers.nms_edge_inputs().items():
        for thr in helpers.NMS_EDGE_THRESHOLDS:
            np.testing.assert_array_equal(ops.nms(dev(d), thr).cpu().numpy(), oracle.nms(d, thr), err_msg="%s thr=%s" % (name, thr))


def test_nms_routes_agree():
    """The single-CTA sweep and both forms of the grid-wide fixed-point iteration (MRCNN_NMS_SWEEP, MRCNN_NMS_PUB; read once per
    process) give the same list."""
    import subprocess, sys, os
    code = ("import numpy as np, torch, sys; sys.path.insert(0, %r); import maskrcnn_b200 as m; from maskrcnn_b200 import synth\n"
            "out = []\n"
            "for n, thr in ((449, 0.5), (2000, 0.3), (6000, 0.7), (12000, 0.6)):\n"
            "    rng = np.random.default_rng(n); b = synth.random_rois(n, n, image=1024.0, min_size=16, max_size=500) * 1024.0\n"
            "    b[n // 2:] = b[: n - n // 2] + rng.uniform(-8, 8, (n - n // 2, 4)).astype(np.float32)\n"
            "    d = np.concatenate([b, synth.unique_scores(n, n)[:, None]], 1).astype(np.float32)\n"
            "    out.append(m.nms(torch.from_numpy(d).cuda(), thr).cpu().numpy())\n"
            "np.save(sys.argv[1], np.concatenate(out))\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import tempfile
    got = {}
    for route, extra in (("serial", {}), ("fixpoint", {}), ("fixpoint-barrier", {"MRCNN_NMS_PUB": "0"})):
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "k.npy")
            env = dict(os.environ, MRCNN_NMS_SWEEP=route.split("-")[0], **extra)
            subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=600)
            got[route] = np.load(path)
    assert len(got["serial"]) > 5000
    np.testing.assert_array_equal(got["serial"], got["fixpoint"])             # published words, no central barrier (<= 148 chunks)
    np.testing.assert_array_equal(got["serial"], got["fixpoint-barrier"])     # the grid-barrier form (what larger inputs run)


def test_nms_properties_full_size(ops):
    dets = _dets(6000, 77)
    keep = ops.nms(dev(dets), 0.7)
    k = keep.cpu().numpy()
    assert np.all(np.diff(k) > 0)                                    # ascending
    again = ops.nms(dev(dets[k]), 0.7).cpu().numpy()                 # idempotent
    np.testing.assert_array_equal(again, np.arange(len(k)))


def test_nms_hypothesis_properties(ops):
    """Property tests on random box sets drawn by hypothesis, unique scores: the CUDA keep list equals the
    oracle's, is ascending, is a fixed point (NMS of the survivors keeps them all), and does not change when boxes that a
    survivor suppresses are appended with lower scores."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(n=st.integers(1, 300), seed=st.integers(0, 10 ** 6), thr=st.sampled_from([0.0, 0.1, 0.3, 0.5, 0.7, 0.95, 1.0]),
           cluster=st.booleans(), scale=st.sampled_from([64.0, 1024.0]))
    def check(n, seed, thr, cluster, scale):
        rng = np.random.default_rng(seed)
        b = synth.random_rois(n, seed, image=scale, min_size=scale / 64, max_size=scale / 2) * scale
        if cluster and n > 3:
            h = n // 2
            b[h:] = b[: n - h] + rng.uniform(-scale / 128, scale / 128, (n - h, 4)).astype(np.float32)
        if n > 5:
            b[rng.integers(0, n)] = b[rng.integers(0, n)]                      # an exact duplicate (IoU = 1)
        dets = np.concatenate([b, synth.unique_scores(n, seed + 1)[:, None]], 1).astype(np.float32)
        want = oracle.nms(dets, thr)
        keep = ops.nms(dev(dets), thr).cpu().numpy()
        np.testing.assert_array_equal(keep, want)
        assert np.all(np.diff(keep) > 0)
        np.testing.assert_array_equal(ops.nms(dev(dets[keep]), thr).cpu().numpy(), np.arange(len(keep)))
        if 0.0 < thr < 1.0:
            # copies of survivors' boxes with scores below every existing one: IoU = 1 >= thr with a higher-scoring
            # survivor, so each is suppressed and nothing else changes
            extra = dets[keep[: min(5, len(keep))]].copy()
            extra[:, 4] = dets[:, 4].min() * np.linspace(0.5, 0.1, len(extra)).astype(np.float32)
            more = np.concatenate([dets, extra]).astype(np.float32)
            np.testing.assert_array_equal(ops.nms(dev(more), thr).cpu().numpy(), keep)
    check()


def test_roialign_hypothesis_adjoint_and_linearity(ops):
    """Property tests on random shapes: <RoIAlign(x), g> == <x, RoIAlign^T(g)> and RoIAlign(a x + y) ==
    a RoIAlign(x) + RoIAlign(y) up to fp32 rounding, for pools / channel counts / level sizes drawn by hypothesis."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=12, deadline=None, derandomize=True)
    @given(pool=st.sampled_from([1, 2, 7, 14]), C=st.sampled_from([4, 12, 64, 132]), size=st.sampled_from([64, 128, 192]),
           N=st.integers(1, 60), B=st.integers(1, 3), seed=st.integers(0, 10 ** 6), chl=st.booleans())
    def check(pool, C, size, N, B, seed, chl):
        fms = synth.feature_pyramid(B, C, seed, image=size)
        boxes = synth.random_rois(N, seed + 1, image=float(size), min_size=4, max_size=size * 0.9)
        ind = np.random.default_rng(seed).integers(0, B, N).astype(np.int32)
        conv = (lambda a: cl(dev(a))) if chl else dev
        xs = [conv(f).requires_grad_(True) for f in fms]
        out = ops.pyramid_roi_align(xs, dev(boxes), dev(ind), pool, (size, size, 3))
        g = torch.randn_like(out)
        out.backward(g)
        lhs = float((out.detach().double() * g.double()).sum())
        rhs = float(sum((x.detach().double() * x.grad.double()).sum() for x in xs))
        assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))
        ys = [torch.randn_like(x) for x in xs]
        with torch.no_grad():
            oy = ops.pyramid_roi_align(ys, dev(boxes), dev(ind), pool, (size, size, 3))
            oz = ops.pyramid_roi_align([2.5 * x + y for x, y in zip(xs, ys)], dev(boxes), dev(ind), pool, (size, size, 3))
        assert float((oz - (2.5 * out.detach() + oy)).abs().max()) <= 1e-4 * float(oz.abs().max() + 1.0)
    check()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_nms_golden(ops, tag):
    g = golden()
    got = ops.nms(dev(g[f"nms_{tag}_in_dets"]), float(g[f"nms_{tag}_in_thr"]))
    np.testing.assert_array_equal(got.cpu().numpy(), g[f"nms_{tag}_out_keep"])


def test_cpu_tensor_is_rejected(ops):
    with pytest.raises(TypeError):
        ops.nms(torch.zeros(4, 5), 0.5)
    with pytest.raises(TypeError):
        ops.CropFunction(7, 7)(torch.zeros(1, 4, 8, 8), torch.zeros(1, 4), torch.zeros(1, dtype=torch.int32))


# ------------------------------------------------------------------ crop_and_resize
def _crop_case(B, C, H, W, N, seed):
    rng = np.random.default_rng(seed)
    img = rng.standard_normal((B, C, H, W), dtype=np.float32)
    boxes = synth.random_rois(N, seed, image=64.0, min_size=4, max_size=60)
    boxes[0] += 0.3
    boxes[1] -= 0.25
    if N > 4:
        boxes[2] = boxes[2][[2, 3, 0, 1]]       # inverted
        boxes[3] = [0.5, 0.5, 0.5, 0.5]         # a point
        boxes[4] = [0.0, 0.0, 1.0, 1.0]         # integer sample positions
    ind = rng.integers(0, B, N).astype(np.int32)
    return img, boxes, ind


@pytest.mark.parametrize("B,C,H,W,N,ch,cw,ev", [
    (1, 3, 16, 16, 7, 7, 7, 0.0), (2, 5, 32, 24, 33, 14, 14, 0.0), (3, 1, 64, 64, 9, 28, 28, 0.0),
    (1, 4, 8, 8, 5, 1, 1, 0.0), (2, 2, 9, 13, 6, 1, 5, -1.5), (2, 8, 33, 17, 11, 3, 1, 2.0),
    (2, 64, 20, 20, 40, 7, 7, 0.0), (1, 256, 16, 16, 10, 14, 14, 0.5), (2, 72, 12, 10, 13, 5, 9, 0.0)])
@pytest.mark.parametrize("channels_last", [False, True])
def test_crop_fwd_bwd(ops, B, C, H, W, N, ch, cw, ev, channels_last):
    img, boxes, ind = _crop_case(B, C, H, W, N, B * 100 + N + C)
    want = oracle.crop_forward(img, boxes, ind, ch, cw, ev)
    t = dev(img)
    if channels_last:
        t = cl(t)
    t.requires_grad_(True)
    out = ops.CropFunction(ch, cw, ev)(t, dev(boxes), dev(ind))
    assert tuple(out.shape) == want.shape
    np.testing.assert_array_equal(out.detach().cpu().numpy(), want)          # bit-exact forward
    g = np.random.default_rng(5).standard_normal(want.shape, dtype=np.float32)
    gt = dev(g)
    if channels_last:
        gt = cl(gt)
    out.backward(gt)
    want_g = oracle.crop_backward(g, boxes, ind, img.shape)
    assert rel_err(t.grad.cpu().numpy(), want_g) <= BWD_TOL
    ops.check_device_errors()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_crop_golden(ops, tag):
    g = golden()
    want = g[f"crop_{tag}_out_crops"]
    t = dev(g[f"crop_{tag}_in_image"]).requires_grad_(True)
    out = ops.CropFunction(want.shape[2], want.shape[3], float(g[f"crop_{tag}_in_ev"]))(
        t, dev(g[f"crop_{tag}_in_boxes"]), dev(g[f"crop_{tag}_in_ind"]))
    np.testing.assert_array_equal(out.detach().cpu().numpy(), want)
    out.backward(dev(g[f"crop_{tag}_in_grads"]))
    assert rel_err(t.grad.cpu().numpy(), g[f"crop_{tag}_out_gimage"]) <= BWD_TOL


def test_crop_bad_box_index_is_reported(ops):
    img = torch.zeros(1, 4, 8, 8, device="cuda")
    out = ops.CropFunction(2, 2, 3.0)(img, torch.tensor([[0., 0., 1., 1.]], device="cuda"),
                                      torch.tensor([1], dtype=torch.int32, device="cuda"))
    assert torch.all(out == 3.0)
    with pytest.raises(ops.MrcnnError):
        ops.check_device_errors()
    ops.check_device_errors()   # flag is cleared


def test_crop_adjoint_property(ops):
    """<crop(x), g> == <x, crop_bwd(g)> at mask-target size (C=1, 28x28, model.py:501-502)."""
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal((6, 1, 256, 256), dtype=np.float32)).cuda().requires_grad_(True)
    boxes = dev(synth.random_rois(6, 9))
    ind = torch.arange(6, dtype=torch.int32, device="cuda")
    y = ops.CropFunction(28, 28, 0)(x, boxes, ind)
    g = torch.randn_like(y)
    y.backward(g)
    lhs = (y.detach().double() * g.double()).sum().item()
    rhs = (x.detach().double() * x.grad.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)


# ------------------------------------------------------------------ PyramidROIAlign
@pytest.mark.parametrize("pool", [7, 14])
@pytest.mark.parametrize("channels_last", [True, False])
@pytest.mark.parametrize("B,C,size,N", [(1, 64, 512, 200), (3, 256, 256, 120), (2, 40, 512, 64)])
def test_pyramid_roi_align(ops, pool, channels_last, B, C, size, N):
    fms = synth.feature_pyramid(B, C, 3 + B, image=size)
    boxes = synth.random_rois(N, 5 + N, image=float(size), min_size=6, max_size=size * 0.9)
    ind = np.random.default_rng(N).integers(0, B, N).astype(np.int32)
    want, lv = oracle.pyramid_roi_align_fwd(fms, boxes, ind, pool, float(size * size))
    ts = []
    for f in fms:
        t = dev(f)
        if channels_last:
            t = cl(t)
        ts.append(t.requires_grad_(True))
    out = ops.pyramid_roi_align(ts, dev(boxes), dev(ind), pool, (size, size, 3))
    np.testing.assert_array_equal(out.detach().cpu().numpy(), want)          # bit-exact, input box order
    g = np.random.default_rng(8).standard_normal(want.shape, dtype=np.float32)
    out.backward(dev(g))                                                     # NCHW-contiguous incoming grad
    want_g = oracle.pyramid_roi_align_bwd(g, [f.shape for f in fms], boxes, ind, float(size * size))
    for t, w in zip(ts, want_g):
        assert rel_err(t.grad.cpu().numpy(), w) <= BWD_TOL
    ops.check_device_errors()


def test_roi_align_dropin_golden(ops):
    g = golden()
    size = int(g["pyr_in_image_size"])
    for pool in (7, 14):
        for chl in (False, True):
            ts = []
            for l in range(4):
                t = dev(g[f"pyr_in_fm{l}"])
                ts.append((cl(t) if chl else t).requires_grad_(True))
            out = ops.roi_align([dev(g["pyr_in_boxes"]).unsqueeze(0)] + [t for t in ts], pool, [size, size, 3])
            np.testing.assert_array_equal(out.detach().cpu().numpy(), g[f"pyr{pool}_out"])
            out.backward(dev(g[f"pyr{pool}_in_grads"]))
            for l in range(4):
                assert rel_err(ts[l].grad.cpu().numpy(), g[f"pyr{pool}_out_gfm{l}"]) <= BWD_TOL


def test_pyramid_levels_full_scale(ops):
    """Level assignment on 1024^2 geometry: 20000 boxes incl. ones hugging the k+0.5 boundaries."""
    from maskrcnn_b200 import _lib
    boxes = synth.random_rois(20000, 77)
    # boxes whose sqrt(h*w) sits within a few ulp of the level boundaries 224*2^(k-4-0.5)
    extra = []
    for k in (2, 3, 4):
        s = 224.0 * 2.0 ** (k - 4 + 0.5) / 1024.0
        for d in range(-3, 4):
            v = np.nextafter(np.float32(s), np.float32(2.0 if d > 0 else 0.0)) if d else np.float32(s)
            for _ in range(abs(d) - 1 if d else 0):
                v = np.nextafter(v, np.float32(2.0 if d > 0 else 0.0))
            extra.append([0.1, 0.1, 0.1 + v, 0.1 + v])
    boxes = np.concatenate([boxes, np.array(extra, np.float32)], 0)
    want = oracle.roi_levels(boxes, 1024.0 * 1024.0)
    N = len(boxes)
    fms = [torch.zeros((1, 4, s, s), device="cuda").contiguous(memory_format=torch.channels_last) for s in (256, 128, 64, 32)]
    out = torch.empty((N, 4, 1, 1), device="cuda")
    lv = torch.empty(N, dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib.mrcnn_pyramid_roi_align_forward(
        _lib.vp4([f.data_ptr() for f in fms]), _lib.i4([256, 128, 64, 32]), _lib.i4([256, 128, 64, 32]), 1, 4, _lib.NHWC,
        dev(boxes).data_ptr(), None, N, 1, 1024.0 * 1024.0, out.data_ptr(), _lib.NHWC, lv.data_ptr(),
        torch.cuda.current_stream().cuda_stream))
    np.testing.assert_array_equal(lv.cpu().numpy(), want)


def test_roialign_linearity_full_size(ops):
    """Size-independent property at BASELINE size (1000 RoIs x 256 ch, 1024^2 pyramid): linear in the feature map."""
    torch.manual_seed(0)
    a = [cl(torch.randn(1, 256, s, s, device="cuda")) for s in (256, 128, 64, 32)]
    b = [cl(torch.randn(1, 256, s, s, device="cuda")) for s in (256, 128, 64, 32)]
    boxes = dev(synth.random_rois(1000, 1234))
    for pool in (7, 14):
        ya = ops.pyramid_roi_align(a, boxes, None, pool, (1024, 1024, 3))
        yb = ops.pyramid_roi_align(b, boxes, None, pool, (1024, 1024, 3))
        yab = ops.pyramid_roi_align([x + 2.0 * y for x, y in zip(a, b)], boxes, None, pool, (1024, 1024, 3))
        assert tuple(ya.shape) == (1000, 256, pool, pool)
        err = (yab - (ya + 2.0 * yb)).abs().max().item()
        assert err <= 1e-5 * yab.abs().max().item()


@pytest.mark.parametrize("algo", ["gather", "scatter"])
def test_roialign_adjoint_full_size(ops, algo):
    """Size-independent property at BASELINE configs[3] size (batch 16 x 512 RoIs x 256 ch, 1024^2 pyramid, 7x7 and 14x14):
    the backward is the adjoint of the forward, <RoIAlign(x), g> == <x, RoIAlign^T(g)>, for both backward algorithms; and
    the fused two-head backward equals the sum of the two single-head ones."""
    torch.manual_seed(1)
    B, C = 16, 256
    boxes = dev(np.concatenate([synth.random_rois(512, 2000 + i) for i in range(B)]))
    ind = torch.arange(B, dtype=torch.int32, device="cuda").repeat_interleave(512)
    ops.set_backward_algorithm(algo)
    grads = {}
    try:
        for pool in (7, 14):
            x = [cl(torch.randn(B, C, s, s, device="cuda")).requires_grad_(True) for s in (256, 128, 64, 32)]
            y = ops.pyramid_roi_align(x, boxes, ind, pool, (1024, 1024, 3))
            g = cl(torch.randn_like(y))
            lhs = (y.double() * g.double()).sum().item()
            y.backward(g)
            rhs = sum((t.detach().double() * t.grad.double()).sum().item() for t in x)
            scale = (y.double() * g.double()).pow(2).sum().sqrt().item()    # size of a random-sign sum of these terms
            assert abs(lhs - rhs) <= 1e-5 * scale
            grads[pool] = (g, [t.grad for t in x])
            del x, y
    finally:
        ops.set_backward_algorithm("auto")
    if algo == "gather":
        shapes = [tuple(t.shape) for t in grads[7][1]]
        both = ops.pyramid_roi_align_backward_pair(grads[7][0], grads[14][0], shapes, boxes, ind, (1024, 1024, 3))
        for t, a, b in zip(both, grads[7][1], grads[14][1]):
            ref = a + b
            assert (t - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


# ------------------------------------------------------------------ proposal layer
@pytest.fixture(params=["lazy", "mask", "hybrid"])
def nms_algo(request, ops):
    """Every NMS implementation of the proposal layer must give the oracle's result."""
    ops.set_proposal_nms(request.param)
    yield request.param
    ops.set_proposal_nms("auto")


@pytest.mark.parametrize("size,pre,post,B,thr", [(256, 500, 200, 1, 0.7), (256, 1000, 300, 3, 0.7), (512, 6000, 1000, 2, 0.7),
                                                 (128, 6000, 1000, 2, 0.7), (256, 6000, 2000, 2, 0.3), (256, 777, 1000, 2, 0.9),
                                                 (256, 3000, 37, 1, 0.5)])
def test_proposal_layer_matches_oracle(ops, nms_algo, size, pre, post, B, thr):
    anchors = synth.pyramid_anchors((size, size))
    rcs, rbs = zip(*[synth.rpn_outputs(anchors, 50 + i, image=float(size), n_clusters=8) for i in range(B)])
    rois, counts = ops.proposal_layer(dev(np.stack(rcs)), dev(np.stack(rbs)), dev(anchors), pre, post, thr,
                                      image_hw=(size, size))
    rois, counts = rois.cpu().numpy(), counts.cpu().numpy()
    for i in range(B):
        want = oracle.proposal_layer(rcs[i], rbs[i], anchors, pre, post, thr, height=float(size), width=float(size))
        assert counts[i] == len(want)
        np.testing.assert_array_equal(rois[i, :counts[i]], want)       # bit-exact boxes and selection
        assert not rois[i, counts[i]:].any()


def test_proposal_layer_many_images(ops, nms_algo):
    """40 images in one call: 320 CTAs in 8-CTA clusters, several waves of clusters on the 148 SMs."""
    size, B = 128, 40
    anchors = synth.pyramid_anchors((size, size))
    rcs, rbs = zip(*[synth.rpn_outputs(anchors, 90 + i, image=float(size), n_clusters=5) for i in range(4)])
    rc = np.stack([rcs[i % 4] for i in range(B)])
    rb = np.stack([rbs[(i * 3) % 4] for i in range(B)])
    rois, counts = ops.proposal_layer(dev(rc), dev(rb), dev(anchors), 2000, 300, 0.6, image_hw=(size, size))
    rois, counts = rois.cpu().numpy(), counts.cpu().numpy()
    for i in range(B):
        want = oracle.proposal_layer(rc[i], rb[i], anchors, 2000, 300, 0.6, height=float(size), width=float(size))
        assert counts[i] == len(want)
        np.testing.assert_array_equal(rois[i, :counts[i]], want)


def test_proposal_layer_with_score_ties(ops, nms_algo):
    """All-equal and heavily tied scores: the selection must be the stable one (lowest anchor index first)."""
    size = 128
    anchors = synth.pyramid_anchors((size, size))
    A = len(anchors)
    rng = np.random.default_rng(1)
    rb = (rng.standard_normal((A, 4)) * 0.3).astype(np.float32)
    for kind in ("all_equal", "few_values"):
        fg = np.full(A, 0.5, np.float32) if kind == "all_equal" else rng.integers(0, 7, A).astype(np.float32) / 8
        rc = np.stack([1 - fg, fg], 1).astype(np.float32)
        rois, counts = ops.proposal_layer(dev(rc[None]), dev(rb[None]), dev(anchors), 600, 100, 0.7, image_hw=(size, size))
        want = oracle.proposal_layer(rc, rb, anchors, 600, 100, 0.7, height=float(size), width=float(size))
        assert int(counts[0]) == len(want)
        np.testing.assert_array_equal(rois[0, :len(want)].cpu().numpy(), want)


def test_proposal_hybrid_tail_and_image_groups(ops):
    """The hybrid NMS on inputs whose top boxes pile up on a few objects (synth.rpn_outputs(converge=0.9): ~600 survivors in 6000
    boxes, so the prefix never holds post_nms survivors and every image goes on into the lazy tail), mixed in one batch with
    ordinary images that finish inside the prefix; and 70 images in one call (more than one cooperative launch can hold)."""
    size = 256
    anchors = synth.pyramid_anchors((size, size))
    ops.set_proposal_nms("hybrid")
    try:
        rcs, rbs = zip(*[synth.rpn_outputs(anchors, 300 + i, image=float(size), n_clusters=4, converge=0.9 if i % 2 else 0.0) for i in range(6)])
        for pre, post, thr in ((6000, 1000, 0.7), (3000, 200, 0.5), (6000, 2048, 0.7)):
            rois, counts = ops.proposal_layer(dev(np.stack(rcs)), dev(np.stack(rbs)), dev(anchors), pre, post, thr, image_hw=(size, size))
            rois, counts = rois.cpu().numpy(), counts.cpu().numpy()
            for i in range(6):
                want = oracle.proposal_layer(rcs[i], rbs[i], anchors, pre, post, thr, height=float(size), width=float(size))
                assert counts[i] == len(want), (pre, post, thr, i, counts[i], len(want))
                np.testing.assert_array_equal(rois[i, :counts[i]], want)
                assert not rois[i, counts[i]:].any()
        B = 70
        rc = np.stack([rcs[i % 6] for i in range(B)])
        rb = np.stack([rbs[(i * 5) % 6] for i in range(B)])
        rois, counts = ops.proposal_layer(dev(rc), dev(rb), dev(anchors), 6000, 1000, 0.7, image_hw=(size, size))
        rois, counts = rois.cpu().numpy(), counts.cpu().numpy()
        for i in range(0, B, 3):
            want = oracle.proposal_layer(rc[i], rb[i], anchors, 6000, 1000, 0.7, height=float(size), width=float(size))
            assert counts[i] == len(want)
            np.testing.assert_array_equal(rois[i, :counts[i]], want)
    finally:
        ops.set_proposal_nms("auto")


def test_proposal_nms_algorithms_agree_at_full_size(ops):
    """configs[1] size (261,888 anchors, 6000 -> 1000, batch 8): lazy, hybrid and mask + sweep NMS return the same bytes; a low
    threshold (heavy suppression: the lazy kernel walks all 94 chunks, the hybrid's prefix leaves most of the work to its tail) as
    well."""
    anchors = synth.pyramid_anchors((1024, 1024))
    rc, rb = zip(*[synth.rpn_outputs(anchors, 60 + i) for i in range(2)])
    rc_d, rb_d, an_d = dev(np.stack([rc[i % 2] for i in range(8)])), dev(np.stack([rb[i % 2] for i in range(8)])), dev(anchors)
    try:
        for thr, post in ((0.7, 1000), (0.1, 1000), (0.5, 2048)):
            out = {}
            for algo in ("lazy", "mask", "hybrid"):
                ops.set_proposal_nms(algo)
                out[algo] = ops.proposal_layer(rc_d, rb_d, an_d, 6000, post, thr)
            assert torch.equal(out["lazy"][0], out["mask"][0]) and torch.equal(out["lazy"][1], out["mask"][1])
            assert torch.equal(out["hybrid"][0], out["mask"][0]) and torch.equal(out["hybrid"][1], out["mask"][1])
            assert int(out["lazy"][1].min()) > 0
    finally:
        ops.set_proposal_nms("auto")


def test_proposal_golden_and_dropin(ops):
    import types
    g = golden()
    size = int(g["prop_in_image_size"])
    anchors = synth.pyramid_anchors((size, size))
    cfg = types.SimpleNamespace(RPN_NMS_MAX_ROIS_NUM=200, RPN_NMS_THRESHOLD=0.7, RPN_BBOX_STD_DEV=[0.1, 0.1, 0.2, 0.2],
                                IMAGE_SHAPE=np.array([size, size, 3]), GPU_COUNT=1)
    stub = types.SimpleNamespace(config=cfg, anchors=dev(anchors))
    got = ops.rpn_refine(stub, dev(g["prop_in_rpn_class"]).unsqueeze(0), dev(g["prop_in_rpn_bbox"]).unsqueeze(0))
    want = g["prop_out_rois"]
    assert tuple(got.shape) == (1,) + want.shape
    from helpers import ulp_diff
    assert ulp_diff(got[0].cpu().numpy(), want).max() <= 4      # reference torch.exp is not correctly rounded


# ------------------------------------------------------------------ detection layer
@pytest.fixture(params=["lazy", "mask"])
def det_algo(request, ops):
    """Both NMS implementations of the detection layer must give the oracle's result."""
    ops.set_detection_nms(request.param)
    yield request.param
    ops.set_detection_nms("auto")


@pytest.mark.parametrize("N,B,min_conf,D,NC", [(1000, 3, 0.0, 100, 81), (400, 2, 0.7, 50, 81), (70, 1, 0.0, 100, 81), (1500, 2, 0.0, 100, 81),
                                               (1000, 2, 0.0, 300, 3), (1000, 2, 0.0, 1000, 2), (640, 1, 0.0, 7, 5), (129, 2, 0.0, 1024, 4)])
def test_detection_layer_matches_oracle(ops, det_algo, N, B, min_conf, D, NC):
    """81 classes (COCO) and few-class cases, where nearly every pair of boxes shares a class: heavy suppression, the lazy
    NMS walks many chunks; max_inst from 7 to more than the number of RoIs."""
    rois = np.stack([synth.random_rois(N, 300 + i) for i in range(B)])
    pd = [synth.head_outputs(N, NC, 400 + i) for i in range(B)]
    probs = np.stack([p for p, _ in pd])
    deltas = np.stack([d for _, d in pd])
    windows = np.tile(np.array([[0, 0, 1024, 1024]], np.float32), (B, 1))
    windows[-1] = [64, 32, 900, 1000]
    dets, counts, index = ops.detection_layer(dev(rois), dev(probs), dev(deltas), dev(windows), min_conf, 0.3, D,
                                              return_index=True)
    dets, counts, index = dets.cpu().numpy(), counts.cpu().numpy(), index.cpu().numpy()
    for i in range(B):
        want, widx = oracle.detection_layer(rois[i], probs[i], deltas[i], windows[i], min_conf, 0.3, D, return_index=True)
        assert counts[i] == len(want) and len(want) > 0
        np.testing.assert_array_equal(dets[i, :counts[i]], want)
        np.testing.assert_array_equal(index[i, :counts[i]], widx)
        assert not dets[i, counts[i]:].any()


def test_detection_nms_algorithms_agree_at_full_size(ops):
    """configs[4] size (64 images x 1000 RoIs x 81 classes, top-100) and a 2-class variant at several thresholds: lazy and
    mask + sweep return the same bytes."""
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    B, N = 64, 1000
    rois = dev(np.stack([synth.random_rois(N, 500 + i) for i in range(B)]))
    win = torch.tensor([[0, 0, 1024, 1024]], dtype=torch.float32, device="cuda").repeat(B, 1)
    try:
        for NC, thr, D in ((81, 0.3, 100), (2, 0.3, 100), (2, 0.05, 50), (3, 0.7, 600)):
            probs = torch.softmax(3 * torch.randn(B, N, NC, device="cuda", generator=g), -1)
            deltas = 0.1 * torch.randn(B, N, NC, 4, device="cuda", generator=g)
            out = {}
            for algo in ("lazy", "mask"):
                ops.set_detection_nms(algo)
                out[algo] = ops.detection_layer(rois, probs, deltas, win, 0.0, thr, D, return_index=True)
            for a, b in zip(out["lazy"], out["mask"]):
                assert torch.equal(a, b)
            assert int(out["lazy"][1].min()) > 0
    finally:
        ops.set_detection_nms("auto")


def test_detection_golden_and_dropin(ops, det_algo):
    import types
    g = golden()
    cfg = types.SimpleNamespace(RPN_BBOX_STD_DEV=np.array([0.1, 0.1, 0.2, 0.2]), IMAGE_SHAPE=np.array([1024, 1024, 3]),
                                GPU_COUNT=1, DETECTION_MIN_CONFIDENCE=0, DETECTION_NMS_THRESHOLD=0.3,
                                DETECTION_MAX_INSTANCES=100)
    ci, sc, bx = ops.mrn_refine(types.SimpleNamespace(config=cfg), dev(g["det_in_rois"]).unsqueeze(0),
                                dev(g["det_in_probs"]), dev(g["det_in_deltas"]), g["det_in_window"])
    want = g["det_out"]
    assert ci.dtype == torch.int64 and tuple(ci.shape) == (1, len(want))
    np.testing.assert_array_equal(bx[0].cpu().numpy(), want[:, :4])
    np.testing.assert_array_equal(sc[0].cpu().numpy(), want[:, 4])
    np.testing.assert_array_equal(ci[0].cpu().numpy(), want[:, 5].astype(np.int64))
    # nothing survives -> (None, None, None), model.py:1445-1447
    probs = torch.zeros(10, 81, device="cuda")
    probs[:, 0] = 1.0
    assert ops.mrn_refine(types.SimpleNamespace(config=cfg), torch.rand(1, 10, 4, device="cuda"), probs,
                          torch.zeros(10, 81, 4, device="cuda"), g["det_in_window"]) == (None, None, None)


@pytest.mark.parametrize("pool", [7, 14, 5])
@pytest.mark.parametrize("out_cl", [True, False])
def test_pyramid_grouped_by_image_backward(ops, pool, out_cl):
    """rois_per_image: per-image clear + scatter in the backward; channels-last crops and upstream grads."""
    B, C, size = 3, 128, 256
    counts = [40, 0, 75]
    fms = synth.feature_pyramid(B, C, 11, image=size)
    boxes = synth.random_rois(sum(counts), 12, image=float(size), min_size=6, max_size=size * 0.9)
    ind = np.repeat(np.arange(B), counts).astype(np.int32)
    want, _ = oracle.pyramid_roi_align_fwd(fms, boxes, ind, pool, float(size * size))
    ts = [cl(dev(f)).requires_grad_(True) for f in fms]
    out = ops.pyramid_roi_align(ts, dev(boxes), None, pool, (size, size, 3), out_channels_last=out_cl, rois_per_image=counts)
    assert out.is_contiguous(memory_format=torch.channels_last if out_cl else torch.contiguous_format)
    np.testing.assert_array_equal(out.detach().cpu().numpy(), want)
    g = np.random.default_rng(8).standard_normal(want.shape, dtype=np.float32)
    gt = dev(g)
    out.backward(cl(gt) if out_cl else gt)
    want_g = oracle.pyramid_roi_align_bwd(g, [f.shape for f in fms], boxes, ind, float(size * size))
    for t, w in zip(ts, want_g):
        assert rel_err(t.grad.cpu().numpy(), w) <= BWD_TOL


@pytest.mark.parametrize("mode", ["gather", "auto", "scatter"])
@pytest.mark.parametrize("pool", [7, 14, 5, 1])
@pytest.mark.parametrize("B,C,size,N", [(3, 128, 256, 150), (2, 320, 200, 90), (1, 40, 512, 300), (4, 256, 160, 64)])
def test_pyramid_backward_gather(ops, pool, B, C, size, N, mode):
    """Channels-last upstream gradient + channels-last pyramid -> the row-owner gather backward
    (no atomics on gradient data), MRCNN_BWD_AUTO or the scatter.
    Level maps whose sides are not multiples of the 8-pixel unit, C not a multiple of 128 / 256 (partial and
    several channel passes), inverted / out-of-image boxes."""
    fms = synth.feature_pyramid(B, C, 21 + B, image=size)
    boxes = synth.random_rois(N, 22 + N, image=float(size), min_size=6, max_size=size * 0.9)
    boxes[0] += 0.4                       # partly outside
    boxes[1] = boxes[1][[2, 3, 0, 1]]     # inverted
    boxes[2] = [0.0, 0.0, 1.0, 1.0]       # integer sample positions
    boxes[3] = [0.3, 0.3, 0.3, 0.3]       # a point
    ind = np.random.default_rng(N).integers(0, B, N).astype(np.int32)
    want, _ = oracle.pyramid_roi_align_fwd(fms, boxes, ind, pool, float(size * size))
    g = np.random.default_rng(8).standard_normal(want.shape, dtype=np.float32)
    want_g = oracle.pyramid_roi_align_bwd(g, [f.shape for f in fms], boxes, ind, float(size * size))
    grads = []
    ops.set_backward_algorithm(mode)
    for rep in range(2):
        ts = [cl(dev(f)).requires_grad_(True) for f in fms]
        out = ops.pyramid_roi_align(ts, dev(boxes), dev(ind), pool, (size, size, 3), out_channels_last=True)
        np.testing.assert_array_equal(out.detach().cpu().numpy(), want)
        out.backward(cl(dev(g)))
        grads.append([t.grad.cpu().numpy() for t in ts])
    ops.set_backward_algorithm("auto")
    for a, w in zip(grads[0], want_g):
        assert rel_err(a, w) <= BWD_TOL
    for a, b in zip(grads[0], grads[1]):             # run to run: same sums up to fp32 association order
        assert rel_err(a, b) <= BWD_TOL
    ops.check_device_errors()


@pytest.mark.parametrize("planning", [True, False])
@pytest.mark.parametrize("pool,B,C,size,N", [(7, 3, 256, 256, 200), (14, 2, 128, 200, 90), (14, 4, 64, 512, 700)])
def test_backward_planning(ops, planning, pool, B, C, size, N):
    """The gather backward's item queues built by the FORWARD on a side stream (mrcnn_pyramid_roi_align_backward_plan) and
    the one-launch backward over them (_planned) against the oracle, next to the unplanned path; a plan serves several
    backward calls (retain_graph)."""
    from maskrcnn_b200 import _lib
    fms = synth.feature_pyramid(B, C, 61 + B, image=size)
    boxes = synth.random_rois(N, 62 + N, image=float(size), min_size=6, max_size=size * 0.9)
    boxes[0] += 0.4
    boxes[1] = boxes[1][[2, 3, 0, 1]]
    ind = np.random.default_rng(N).integers(0, B, N).astype(np.int32)
    g = np.random.default_rng(9).standard_normal((N, C, pool, pool), dtype=np.float32)
    want_g = oracle.pyramid_roi_align_bwd(g, [f.shape for f in fms], boxes, ind, float(size * size))
    calls = {"plan": 0, "planned": 0}
    real_plan, real_planned = _lib.lib.mrcnn_pyramid_roi_align_backward_plan, _lib.lib.mrcnn_pyramid_roi_align_backward_planned

    def count(name, fn):
        def wrapped(*a):
            calls[name] += 1
            return fn(*a)
        return wrapped
    _lib.lib.mrcnn_pyramid_roi_align_backward_plan = count("plan", real_plan)
    _lib.lib.mrcnn_pyramid_roi_align_backward_planned = count("planned", real_planned)
    ops.set_backward_planning(planning)
    try:
        ts = [cl(dev(f)).requires_grad_(True) for f in fms]
        out = ops.pyramid_roi_align(ts, dev(boxes), dev(ind), pool, (size, size, 3), out_channels_last=True)
        out.backward(cl(dev(g)), retain_graph=True)
        first = [t.grad.clone() for t in ts]
        out.backward(cl(dev(g)))                       # the same plan again: gradients accumulate to twice the value
        # no feature map requires grad -> no plan is built
        ops.pyramid_roi_align([cl(dev(f)) for f in fms], dev(boxes), dev(ind), pool, (size, size, 3), out_channels_last=True)
        # an NCHW upstream gradient cannot use the plan: the regular path answers
        ts2 = [cl(dev(f)).requires_grad_(True) for f in fms]
        out2 = ops.pyramid_roi_align(ts2, dev(boxes), dev(ind), pool, (size, size, 3), out_channels_last=True)
        out2.backward(dev(g).contiguous())
    finally:
        ops.set_backward_planning(True)
        _lib.lib.mrcnn_pyramid_roi_align_backward_plan, _lib.lib.mrcnn_pyramid_roi_align_backward_planned = real_plan, real_planned
    assert calls == ({"plan": 2, "planned": 2} if planning else {"plan": 0, "planned": 0})
    for a, t, t2, w in zip(first, ts, ts2, want_g):
        assert rel_err(a.cpu().numpy(), w) <= BWD_TOL
        assert rel_err(t.grad.cpu().numpy(), 2.0 * w) <= BWD_TOL
        assert rel_err(t2.grad.cpu().numpy(), w) <= BWD_TOL
    ops.check_device_errors()


# ------------------------------------------------------------------ detection-target layer (mrn_samples)
def _target_cfg(train_rois):
    import types
    return types.SimpleNamespace(GPU_COUNT=1, TRAIN_ROIS_PER_IMAGE=train_rois, ROI_POSITIVE_RATIO=0.33,
                                 BBOX_STD_DEV=np.array([0.1, 0.1, 0.2, 0.2]), MASK_SHAPE=[28, 28])


def _check_targets(got, want):
    g_rois, g_cls, g_d, g_m = (t.cpu().numpy() for t in got)
    w_rois, w_cls, w_d, w_m = want
    assert g_rois.shape == w_rois.shape and g_cls.dtype == np.int32
    np.testing.assert_array_equal(g_rois, w_rois)
    np.testing.assert_array_equal(g_cls, w_cls)
    np.testing.assert_array_equal(g_d, w_d)        # dy, dx exact; dh, dw: both sides use the correctly rounded log
    np.testing.assert_array_equal(g_m, w_m)


@pytest.mark.parametrize("n_rois,n_gt,n_crowd,n_pad,train_rois,image,seed", [
    (1000, 20, 0, 0, 512, 256, 1), (1000, 20, 3, 4, 512, 256, 2), (333, 7, 0, 2, 100, 200, 3), (65, 3, 1, 0, 512, 96, 4),
    (2000, 100, 5, 10, 512, 128, 5), (500, 1, 0, 0, 64, 64, 6)])
def test_mrn_samples_dropin_matches_oracle(ops, n_rois, n_gt, n_crowd, n_pad, train_rois, image, seed):
    """ops.mrn_samples (reference signature; draws torch.randperm like the reference) vs the oracle under the same seed."""
    rois, cls, gt, masks = synth.target_inputs(n_rois, n_gt, seed, image=image, n_crowd=n_crowd, n_pad=n_pad)
    torch.manual_seed(40 + seed)
    want = oracle.mrn_samples(rois, cls, gt, masks, train_rois, 0.33, [0.1, 0.1, 0.2, 0.2], (28, 28),
                              lambda n: torch.randperm(n).numpy())
    torch.manual_seed(40 + seed)
    got = ops.mrn_samples(dev(rois)[None], dev(cls)[None], dev(gt)[None], dev(masks)[None], _target_cfg(train_rois))
    assert len(want[0]) > 3
    _check_targets(got, want)
    ops.check_device_errors()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_mrn_samples_golden(ops, tag):
    """Against what the reference's model.mrn_samples returned (tests/golden/make_golden_targets.py) under the recorded seed."""
    from helpers import golden_targets, ulp_diff
    g = golden_targets()
    torch.manual_seed(int(g[f"{tag}_in_seed"]))
    got = ops.mrn_samples(dev(g[f"{tag}_in_rois"])[None], dev(g[f"{tag}_in_cls"])[None], dev(g[f"{tag}_in_gt"])[None],
                          dev(g[f"{tag}_in_masks"].astype(np.float32))[None], _target_cfg(int(g[f"{tag}_in_train_rois"])))
    g_rois, g_cls, g_d, g_m = (t.cpu().numpy() for t in got)
    np.testing.assert_array_equal(g_rois, g[f"{tag}_out_rois"])
    np.testing.assert_array_equal(g_cls, g[f"{tag}_out_cls"])
    np.testing.assert_array_equal(g_d[:, :2], g[f"{tag}_out_deltas"][:, :2])
    assert ulp_diff(g_d[:, 2:], g[f"{tag}_out_deltas"][:, 2:]).max() <= 2      # torch.log on CPU is not correctly rounded
    np.testing.assert_array_equal(g_m, g[f"{tag}_out_masks"].astype(np.float32))


def test_mrn_samples_edge_cases(ops):
    cfg = _target_cfg(64)
    rois, cls, gt, masks = synth.target_inputs(50, 3, 9, image=64, positive_fraction=0.0)
    rois[:] = [0.0, 0.0, 0.01, 0.01]                                 # no positive -> four empty tensors (model.py:563-574)
    out = ops.mrn_samples(dev(rois)[None], dev(cls)[None], dev(gt)[None], dev(masks)[None], cfg)
    assert all(t.numel() == 0 for t in out)
    rois, cls, gt, masks = synth.target_inputs(40, 2, 10, image=64, positive_fraction=1.0)   # every proposal positive
    torch.manual_seed(1)
    want = oracle.mrn_samples(rois, cls, gt, masks, 64, 0.33, [0.1, 0.1, 0.2, 0.2], (28, 28), lambda n: torch.randperm(n).numpy())
    torch.manual_seed(1)
    got = ops.mrn_samples(dev(rois)[None], dev(cls)[None], dev(gt)[None], dev(masks)[None], cfg)
    _check_targets(got, want)
    with pytest.raises(TypeError):
        ops.mrn_samples(torch.from_numpy(rois)[None], torch.from_numpy(cls)[None], torch.from_numpy(gt)[None],
                        torch.from_numpy(masks)[None], cfg)          # no CPU path


@pytest.mark.parametrize("B,n_rois,n_gt,train_rois,image", [(4, 1000, 24, 512, 128), (3, 257, 9, 100, 96), (2, 8192, 16, 512, 64)])
def test_detection_targets_batched_matches_oracle(ops, B, n_rois, n_gt, train_rois, image):
    """Sync-free batched layer: random keys stand for the two torch.randperm draws (perm = stable argsort of the first
    P / Q keys); images with crowds, padding and different positive fractions in one call."""
    rng = np.random.default_rng(B * 1000 + n_rois)
    ins = [synth.target_inputs(n_rois, n_gt, 70 + b, image=image, n_crowd=b % 3, n_pad=(2 * b) % 5,
                               positive_fraction=(0.0 if b == 2 else 0.1 + 0.2 * b)) for b in range(B)]
    if B > 2:
        ins[2][0][:] = [0.0, 0.0, 0.01, 0.01]                        # an image without positives
    kp = rng.permutation(B * n_rois).reshape(B, n_rois).astype(np.float32) / (B * n_rois)
    kn = rng.permutation(B * n_rois).reshape(B, n_rois).astype(np.float32) / (B * n_rois)
    stack = lambda k: np.stack([x[k] for x in ins])  # noqa: E731
    o_rois, o_cls, o_d, o_m, take = ops.detection_targets(dev(stack(0)), dev(stack(1)), dev(stack(2)), dev(stack(3)), dev(kp), dev(kn),
                                                          train_rois_per_image=train_rois)
    take = take.cpu().numpy()
    for b in range(B):
        draws = [lambda n, b=b: np.argsort(kp[b, :n], kind="stable"), lambda n, b=b: np.argsort(kn[b, :n], kind="stable")]
        it = iter(draws)
        want = oracle.mrn_samples(*ins[b], train_rois, 0.33, [0.1, 0.1, 0.2, 0.2], (28, 28), lambda n: next(it)(n))
        t = len(want[0])
        assert take[b].sum() == t and take[b][0] == (want[1] > 0).sum()
        _check_targets((o_rois[b, :t], o_cls[b, :t], o_d[b, :t], o_m[b, :t]), want)
        assert float(o_rois[b, t:].abs().sum()) == 0.0 and int(o_cls[b, t:].abs().sum()) == 0 and float(o_m[b, t:].sum()) == 0.0
    ops.check_device_errors()


@pytest.mark.parametrize("B,C,size,N", [(3, 256, 256, 150), (2, 96, 200, 90), (1, 512, 128, 40)])
def test_pyramid_backward_pair_is_the_sum_of_both_heads(ops, B, C, size, N):
    """One fused row-owner gather for the 7x7 and the 14x14 head == the sum of the two separate backward passes
    (what autograd accumulates in the reference), and `accumulate` adds on top."""
    fms = synth.feature_pyramid(B, C, 31 + B, image=size)
    boxes = synth.random_rois(N, 32 + N, image=float(size), min_size=6, max_size=size * 0.9)
    boxes[0] += 0.4
    boxes[1] = boxes[1][[2, 3, 0, 1]]
    ind = np.random.default_rng(N).integers(0, B, N).astype(np.int32)
    rng = np.random.default_rng(9)
    g7 = rng.standard_normal((N, C, 7, 7), dtype=np.float32)
    g14 = rng.standard_normal((N, C, 14, 14), dtype=np.float32)
    shapes = [f.shape for f in fms]
    w7 = oracle.pyramid_roi_align_bwd(g7, shapes, boxes, ind, float(size * size))
    w14 = oracle.pyramid_roi_align_bwd(g14, shapes, boxes, ind, float(size * size))
    got = ops.pyramid_roi_align_backward_pair(cl(dev(g7)), cl(dev(g14)), shapes, dev(boxes), dev(ind), (size, size, 3))
    for t, a, b in zip(got, w7, w14):
        assert t.is_contiguous(memory_format=torch.channels_last)
        assert rel_err(t.cpu().numpy(), a + b) <= BWD_TOL
    again = ops.pyramid_roi_align_backward_pair(cl(dev(g7)), cl(dev(g14)), shapes, dev(boxes), dev(ind), (size, size, 3),
                                                out=got, accumulate=True)
    for t, a, b in zip(again, w7, w14):
        assert rel_err(t.cpu().numpy(), 2 * (a + b)) <= BWD_TOL
    ops.check_device_errors()


def test_pyramid_roi_align_pair_autograd(ops):
    """Both heads from one autograd node: forwards bit-exact, the fused backward == sum of the two oracle backwards."""
    B, C, size, N = 2, 128, 256, 120
    fms = synth.feature_pyramid(B, C, 77, image=size)
    boxes = synth.random_rois(N, 78, image=float(size), min_size=6, max_size=size * 0.9)
    ind = np.random.default_rng(5).integers(0, B, N).astype(np.int32)
    ts = [cl(dev(f)).requires_grad_(True) for f in fms]
    o7, o14 = ops.pyramid_roi_align_pair(ts, dev(boxes), dev(ind), (7, 14), (size, size, 3))
    w7, _ = oracle.pyramid_roi_align_fwd(fms, boxes, ind, 7, float(size * size))
    w14, _ = oracle.pyramid_roi_align_fwd(fms, boxes, ind, 14, float(size * size))
    np.testing.assert_array_equal(o7.detach().cpu().numpy(), w7)
    np.testing.assert_array_equal(o14.detach().cpu().numpy(), w14)
    rng = np.random.default_rng(6)
    g7, g14 = rng.standard_normal(w7.shape, dtype=np.float32), rng.standard_normal(w14.shape, dtype=np.float32)
    (o7 * dev(g7)).sum().add((o14 * dev(g14)).sum()).backward()
    shapes = [f.shape for f in fms]
    a = oracle.pyramid_roi_align_bwd(g7, shapes, boxes, ind, float(size * size))
    b = oracle.pyramid_roi_align_bwd(g14, shapes, boxes, ind, float(size * size))
    for t, x, y in zip(ts, a, b):
        assert rel_err(t.grad.cpu().numpy(), x + y) <= BWD_TOL


# ------------------------------------------------------------------ RPN anchor matching (data.rpn_samples)
def _f64_ulp(a, b):
    return np.abs(np.ascontiguousarray(a).view(np.int64) - np.ascontiguousarray(b).view(np.int64)).max() if a.size else 0


@pytest.mark.parametrize("image,n_gt,n_crowd,T,seed", [(1024, 20, 0, 256, 1), (1024, 40, 3, 256, 2), (512, 5, 1, 64, 3), (256, 1, 0, 32, 4),
                                                       (1024, 100, 5, 512, 5)])
def test_rpn_samples_matches_oracle(ops, image, n_gt, n_crowd, T, seed):
    """ops.rpn_samples (reference signature, numpy in / numpy out) vs the oracle under the same numpy seed; 261,888 anchors
    at 1024^2."""
    import types
    anchors = synth.pyramid_anchors((image, image)).astype(np.float64)
    cls, gt = synth.rpn_target_inputs(n_gt, 50 + seed, image=image, n_crowd=n_crowd)
    cfg = types.SimpleNamespace(RPN_TRAIN_ANCHORS_PER_IMAGE=T, RPN_BBOX_STD_DEV=np.array([0.1, 0.1, 0.2, 0.2]))
    np.random.seed(seed)
    w_match, w_bbox = oracle.rpn_samples(anchors, cls, gt, T, [0.1, 0.1, 0.2, 0.2], np.random.permutation)
    np.random.seed(seed)
    g_match, g_bbox = ops.rpn_samples(anchors, cls, gt, cfg)
    assert g_match.dtype == np.int32 and g_bbox.dtype == np.float64 and g_bbox.shape == (T, 4)
    np.testing.assert_array_equal(g_match, w_match)
    assert (g_match == 1).sum() >= 1 and (g_match == 1).sum() <= T // 2 and (g_match != 0).sum() <= T
    np.testing.assert_array_equal(g_bbox[:, :2], w_bbox[:, :2])
    assert _f64_ulp(g_bbox[:, 2:], w_bbox[:, 2:]) <= 2          # float64 log: CUDA's vs libm's
    ops.check_device_errors()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_rpn_samples_golden(ops, tag):
    import os
    import types
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_rpn_v1.npz"))
    image = int(g[f"{tag}_in_image"])
    anchors = torch.from_numpy(synth.pyramid_anchors((image, image)).astype(np.float64)).cuda()    # anchors resident on the device
    cfg = types.SimpleNamespace(RPN_TRAIN_ANCHORS_PER_IMAGE=int(g[f"{tag}_in_T"]), RPN_BBOX_STD_DEV=np.array([0.1, 0.1, 0.2, 0.2]))
    np.random.seed(int(g[f"{tag}_in_seed"]))
    match, bbox = ops.rpn_samples(anchors, g[f"{tag}_in_cls"], g[f"{tag}_in_gt"], cfg)
    np.testing.assert_array_equal(match, g[f"{tag}_out_match"].astype(np.int32))
    np.testing.assert_array_equal(bbox[:, :2], g[f"{tag}_out_bbox"][:, :2])
    assert _f64_ulp(bbox[:, 2:], g[f"{tag}_out_bbox"][:, 2:]) <= 2


# ------------------------------------------------------------------ mask paste-back (data.full_masks)
@pytest.mark.parametrize("D,NC,H,W,lo,hi,seed", [(100, 81, 1024, 1024, 16, 800, 1), (40, 81, 256, 256, 1, 200, 2), (24, 5, 200, 333, 1, 60, 3),
                                                 (16, 3, 64, 80, 1, 28, 4), (7, 2, 100, 48, 2, 40, 5), (1, 81, 1024, 1024, 1000, 1024, 6)])
def test_full_masks_matches_oracle(ops, D, NC, H, W, lo, hi, seed):
    rng = np.random.default_rng(seed)
    cls, boxes, masks = synth.mask_head_outputs(D, NC, 70 + seed, image=min(H, W), min_size=lo, max_size=min(hi, min(H, W)))
    masks[::5] = masks[::5] * 1.5 - 0.25                      # saturating values (< 0, > 1)
    if D >= 7:
        boxes[1, 2] = boxes[1, 0] + 28.0                      # unchanged height: no vertical pass
        boxes[2, 3] = boxes[2, 1] + 28.0                      # unchanged width: no horizontal pass
        boxes[3] += np.float32([-9.0, -11.0, -9.0, -11.0])    # leaves the image top / left
        boxes[4, 2:] = np.maximum(boxes[4, 2:], boxes[4, :2] + 3.0)
        boxes[4] += rng.choice(np.float32([0.25, 0.5, 0.75]), 4)
        boxes[5, 2:] = [H + 13.0, W + 5.0]                    # leaves the image bottom / right
    want = oracle.full_masks(cls, boxes, masks, H, W)
    got = ops.full_masks(dev(cls), dev(boxes), dev(masks), H, W)
    assert got.dtype == torch.bool and tuple(got.shape) == (D, H, W)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    ops.check_device_errors()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_full_masks_golden(ops, tag):
    from helpers import golden_masks
    cls, boxes, masks, h, w, want = golden_masks(tag)
    np.testing.assert_array_equal(ops.full_masks(dev(cls), dev(boxes), dev(masks), h, w).cpu().numpy(), want)


def test_full_masks_edge_cases(ops):
    # zero-padded detection rows (empty boxes) give empty masks; batched leading dimension; D = 0
    cls, boxes, masks = synth.mask_head_outputs(12, 4, 5, image=96, n_pad=4)
    got = ops.full_masks(dev(cls).view(2, 6), dev(boxes).view(2, 6, 4), dev(masks).view(2, 6, 4, 28, 28), 96, 96)
    assert tuple(got.shape) == (2, 6, 96, 96)
    want = oracle.full_masks(cls[:8], boxes[:8], masks[:8], 96, 96)
    np.testing.assert_array_equal(got.view(12, 96, 96)[:8].cpu().numpy(), want)
    assert not bool(got.view(12, 96, 96)[8:].any())
    assert ops.full_masks(dev(cls[:0]), dev(boxes[:0]), dev(masks[:0]), 32, 32).shape == (0, 32, 32)
    # inverted and fully outside boxes: nothing is pasted
    boxes[0] = [50.0, 50.0, 20.0, 20.0]
    boxes[1] = [200.0, 10.0, 260.0, 40.0]
    got = ops.full_masks(dev(cls), dev(boxes), dev(masks), 96, 96)
    assert not bool(got[:2].any())
    # a class id outside [0, NC) raises through the device error word
    bad = cls.copy()
    bad[2] = 9
    ops.full_masks(dev(bad), dev(boxes), dev(masks), 96, 96)
    with pytest.raises(RuntimeError):
        ops.check_device_errors()
    ops.check_device_errors()
    with pytest.raises(TypeError):
        ops.full_masks(torch.from_numpy(cls), torch.from_numpy(boxes), torch.from_numpy(masks), 96, 96)


def test_full_masks_full_size_properties(ops):
    """configs[4] size: 64 images x 100 detections would be 6.7 GB of masks; 8 images x 100 here.  Properties: nothing
    outside the box, translation of a box by whole pixels translates its mask, a constant mask > 127/255 fills the box."""
    D, H, W = 800, 1024, 1024
    cls, boxes, masks = synth.mask_head_outputs(D, 81, 99, image=1024)
    got = ops.full_masks(dev(cls), dev(boxes), dev(masks), H, W)
    yy = torch.arange(H, device="cuda").view(1, H, 1)
    xx = torch.arange(W, device="cuda").view(1, 1, W)
    b = dev(boxes).view(D, 4, 1, 1)
    inside = (yy >= b[:, 0]) & (yy < b[:, 2]) & (xx >= b[:, 1]) & (xx < b[:, 3])
    assert not bool((got & ~inside).any())
    assert int(got.sum()) > 0
    # constant masks: 0.6 * 255 = 153 > 127 fills the box exactly, 0.4 * 255 = 102 leaves it empty
    ones = torch.full((D, 81, 28, 28), 0.6, device="cuda")
    assert bool((ops.full_masks(dev(cls), dev(boxes), ones, H, W) == inside).all())
    assert not bool(ops.full_masks(dev(cls), dev(boxes), ones * (0.4 / 0.6), H, W).any())
    # whole-pixel translation
    sub = slice(0, 16)
    small = boxes[sub].copy()
    small[:, [0, 2]] -= small[:, 0:1]
    small[:, [1, 3]] -= small[:, 1:2]
    at0 = ops.full_masks(dev(cls[sub]), dev(small), dev(masks[sub]), H, W)
    for i in range(16):
        y1, x1, y2, x2 = (int(v) for v in boxes[i])
        assert torch.equal(at0[i, :y2 - y1, :x2 - x1], got[i, y1:y2, x1:x2])


# ------------------------------------------------------------------ RPN head output plumbing (rpn_detect)
def _rpn_conv_outputs(B, K, sides, seed):
    rng = np.random.default_rng(seed)
    ls = [(rng.standard_normal((B, 2 * K, h, w)) * 4.0).astype(np.float32) for h, w in sides]
    bs = [rng.standard_normal((B, 4 * K, h, w)).astype(np.float32) for h, w in sides]
    return ls, bs


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("B,K,sides", [(2, 3, [(64, 64), (32, 32), (16, 16), (8, 8), (4, 4)]), (1, 3, [(7, 9), (3, 5)]), (3, 1, [(5, 4)]),
                                       (2, 5, [(6, 6), (1, 1)])])
def test_rpn_pack_matches_oracle(ops, B, K, sides, channels_last):
    ls, bs = _rpn_conv_outputs(B, K, sides, 3 + B + K)
    want = oracle.rpn_pack(ls, bs)
    conv = (lambda a: cl(dev(a))) if channels_last else dev
    logits, cls, bbox, fg = ops.rpn_pack([conv(a) for a in ls], [conv(a) for a in bs])
    np.testing.assert_array_equal(logits.cpu().numpy(), want[0])
    np.testing.assert_array_equal(cls.cpu().numpy(), want[1])       # same operation order, correctly rounded exp: bit-exact
    np.testing.assert_array_equal(bbox.cpu().numpy(), want[2])
    np.testing.assert_array_equal(fg.cpu().numpy(), want[1][:, :, 1])


def test_rpn_pack_golden(ops):
    from helpers import SOFTMAX_TOL, golden_rpnhead
    ls, bs, w_logits, w_class, w_bbox = golden_rpnhead()
    logits, cls, bbox, fg = ops.rpn_pack([dev(a) for a in ls], [dev(a) for a in bs])
    np.testing.assert_array_equal(logits.cpu().numpy(), w_logits)
    np.testing.assert_array_equal(bbox.cpu().numpy(), w_bbox)
    assert np.abs(cls.cpu().numpy() - w_class).max() <= SOFTMAX_TOL
    assert np.abs(fg.cpu().numpy() - w_class[:, :, 1]).max() <= SOFTMAX_TOL


@pytest.mark.parametrize("channels_last", [False, True])
def test_rpn_pack_backward_is_the_inverse_layout(ops, channels_last):
    """The losses read rpn_class_logits and rpn_bbox: their gradients must come back exactly as autograd returns them for
    the reference's permute / view / cat chain."""
    ls, bs = _rpn_conv_outputs(2, 3, [(12, 10), (6, 5), (3, 3)], 21)
    conv = (lambda a: cl(dev(a))) if channels_last else dev
    a_l, a_b = [conv(a).requires_grad_(True) for a in ls], [conv(a).requires_grad_(True) for a in bs]
    r_l, r_b = [dev(a).requires_grad_(True) for a in ls], [dev(a).requires_grad_(True) for a in bs]
    logits, _, bbox, _ = ops.rpn_pack(a_l, a_b)
    ref_logits = torch.cat([t.permute(0, 2, 3, 1).contiguous().view(t.size(0), -1, 2) for t in r_l], 1)
    ref_bbox = torch.cat([t.permute(0, 2, 3, 1).contiguous().view(t.size(0), -1, 4) for t in r_b], 1)
    assert torch.equal(logits, ref_logits) and torch.equal(bbox, ref_bbox)
    g1, g2 = torch.randn_like(ref_logits), torch.randn_like(ref_bbox)
    (logits * g1).sum().backward()
    (ref_logits * g1).sum().backward()
    for a, r in zip(a_l, r_l):
        assert torch.equal(a.grad, r.grad)
    assert all(t.grad is None for t in a_b)          # bbox unused: no gradient, no launch for it
    logits, _, bbox, _ = ops.rpn_pack(a_l, a_b)
    ((bbox * g2).sum() + (logits * g1).sum()).backward()
    (ref_bbox * g2).sum().backward()
    for a, r in zip(a_b, r_b):
        assert torch.equal(a.grad, r.grad)


def test_rpn_detect_dropin_and_fg_proposals(ops):
    """MaskRCNN.rpn_detect drop-in on a stand-in model (the head's convolutions stay on stock PyTorch) against the
    reference's per-level permute / softmax / cat expressed in torch, and the proposal layer fed with the fg
    probabilities alone against the [B,A,2] form."""
    import types
    torch.manual_seed(3)

    class Pad(torch.nn.Module):
        def forward(self, x):
            return torch.nn.functional.pad(x, (1, 1, 1, 1))
    rpn = types.SimpleNamespace(padding=Pad(), conv_shared=torch.nn.Conv2d(16, 32, 3).cuda(), relu=torch.nn.ReLU(),
                                conv_class=torch.nn.Conv2d(32, 6, 1).cuda(), conv_bbox=torch.nn.Conv2d(32, 12, 1).cuda())
    me = types.SimpleNamespace(rpn=rpn)
    size = 128
    feats = [torch.randn(2, 16, size // s, size // s, device="cuda") for s in (4, 8, 16, 32, 64)]
    with torch.no_grad():
        logits, cls, bbox = ops.rpn_detect(me, feats)
        want_l, want_c, want_b = [], [], []
        for p in feats:
            x = rpn.relu(rpn.conv_shared(rpn.padding(p)))
            lg = rpn.conv_class(x).permute(0, 2, 3, 1).contiguous().view(2, -1, 2)
            want_l.append(lg)
            want_c.append(torch.softmax(lg, 2))
            want_b.append(rpn.conv_bbox(x).permute(0, 2, 3, 1).contiguous().view(2, -1, 4))
    assert torch.equal(logits, torch.cat(want_l, 1)) and torch.equal(bbox, torch.cat(want_b, 1))
    assert float((cls - torch.cat(want_c, 1)).abs().max()) <= 1e-6
    anchors = dev(synth.pyramid_anchors((size, size)))
    assert anchors.size(0) == logits.size(1)
    a = ops.proposal_layer(cls, bbox, anchors, 600, 100, 0.7, image_hw=(size, size))
    b = ops.proposal_layer(cls[:, :, 1].contiguous(), bbox, anchors, 600, 100, 0.7, image_hw=(size, size))
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


# ------------------------------------------------------------------ decode_masks (data.py:265-284)
def _decode_case(D, H, W, seed):
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


@pytest.mark.parametrize("D,H,W,crop,scale", [
    (3, 256, 256, (160, 256), 256 / 1920),      # predict.py geometry at 256 (upscale 7.5; 1920 columns: the 128-bit path)
    (4, 256, 256, (256, 189), 0.75),            # odd margins, 252 columns (byte path)
    (4, 200, 200, (173, 200), 0.4161),          # CenterCrop origin 13.5 -> 14 (half-even)
    (5, 128, 160, (128, 160), 2.0),             # downscale by 2 -> 80 columns (128-bit path, 5 taps)
    (3, 120, 90, (101, 81), 3.3),               # downscale by 3.3, 9 taps
    (2, 64, 64, (60, 60), 1.0001),              # same size: both passes are the identity
    (2, 64, 96, (64, 96), 0.5),                 # exact doubling
    (1, 1024, 1024, (640, 1024), 1024 / 1920),  # the real predict.py frame: 1024 -> 1200 x 1920
    (70, 40, 40, (40, 40), 0.3)])               # many small masks
def test_decode_masks_matches_oracle(ops, D, H, W, crop, scale):
    m = _decode_case(D, H, W, D + H)
    want = oracle.decode_masks(m, scale, crop)
    got = ops.decode_masks(dev(m), scale, crop)
    assert got.dtype == torch.uint8 and tuple(got.shape) == want.shape
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    # uint8 0/1 input keeps its pixel values ("tensor NxHxW with 1/0"); a Box-like window object
    import types
    box = types.SimpleNamespace(height=lambda: crop[0], width=lambda: crop[1])
    np.testing.assert_array_equal(ops.decode_masks(dev(m.astype(np.uint8)), scale, box).cpu().numpy(),
                                  oracle.decode_masks(m.astype(np.uint8), scale, crop))
    ops.check_device_errors()


@pytest.mark.parametrize("tag", ["a", "b", "c", "d", "e"])
def test_decode_masks_golden(ops, tag):
    from helpers import golden_decode
    m, scale, crop_hw, want = golden_decode(tag)
    np.testing.assert_array_equal(ops.decode_masks(dev(m), scale, crop_hw).cpu().numpy(), want)


def test_decode_masks_edge_cases(ops):
    m = dev(_decode_case(2, 32, 48, 3))
    assert ops.decode_masks(m, 1, (32, 48)) is m                                  # data.py:267-268
    assert tuple(ops.decode_masks(m[:0], 0.5, (32, 48)).shape) == (0, 64, 96)     # no detections
    with pytest.raises(ValueError):
        ops.decode_masks(m, 0.5, (33, 48))                                        # window larger than the mask
    with pytest.raises(ValueError):
        ops.decode_masks(m, 100.0, (32, 48))                                      # target rounds to 0 pixels: PIL raises too
    with pytest.raises(ValueError):
        ops.decode_masks(m.float(), 0.5, (32, 48))
    with pytest.raises(TypeError):
        ops.decode_masks(m.cpu(), 0.5, (32, 48))
    # a non-contiguous view (full_masks output sliced) and a full-masks -> decode chain against the oracle chain
    cls, boxes, masks = synth.mask_head_outputs(6, 5, 23, image=128)
    pasted = ops.full_masks(dev(cls), dev(boxes), dev(masks), 128, 128)
    got = ops.decode_masks(pasted[::2], 128 / 300, (80, 128)).cpu().numpy()
    want = oracle.decode_masks(oracle.full_masks(cls, boxes, masks, 128, 128)[::2], 128 / 300, (80, 128))
    np.testing.assert_array_equal(got, want)
    # all-False and all-True masks stay constant (the weights of every output pixel sum to 1 in fixed point +- rounding)
    z = torch.zeros(1, 64, 64, dtype=torch.bool, device="cuda")
    assert not ops.decode_masks(z, 0.37, (64, 64)).any()
    assert bool((ops.decode_masks(~z, 0.37, (64, 64)) == 255).all())


# ------------------------------------------------------------------ the whole predict.py flow (BASELINE configs[0])
def test_detect_flow_golden_dropins(ops):
    """The reference's predict.py flow on images/car58a54312d.jpg (random-init ResNet-101-FPN, CPU c++ext ops), recorded
    at the boundary of every operator this repo replaces (tests/golden/make_golden_detect.py), replayed stage by stage
    through the drop-ins: reference tensors in, reference tensors expected."""
    import hashlib
    import types
    from helpers import SOFTMAX_TOL, detect_flow_expected_masks, golden_detect, ulp_diff
    g = golden_detect()
    size = int(g["image_dim"])
    nl = sum(1 for k in g.files if k.startswith("rpn_in_logits_"))
    # rpn_detect (model.py:1294-1304): pack of the head's conv outputs
    logits, cls, bbox, fg = ops.rpn_pack([dev(g[f"rpn_in_logits_{l}"]) for l in range(nl)], [dev(g[f"rpn_in_bbox_{l}"]) for l in range(nl)])
    assert hashlib.sha256(logits.cpu().numpy().tobytes() + bbox.cpu().numpy().tobytes()).digest() == g["rpn_out_logits_bbox_sha256"].tobytes()
    assert np.abs(cls.cpu().numpy() - g["rpn_out_class"]).max() <= SOFTMAX_TOL
    # rpn_refine (model.py:1307-1382) on the de-tied scores (see the generator: the flow's own scores tie at 1.0 by the thousand)
    pre, post = (int(v) for v in g["prop_in_limits"])
    cfg = types.SimpleNamespace(RPN_NMS_MAX_ROIS_NUM=post, RPN_NMS_THRESHOLD=float(g["prop_in_thr"]), RPN_BBOX_STD_DEV=[0.1, 0.1, 0.2, 0.2],
                                IMAGE_SHAPE=np.array([size, size, 3]), GPU_COUNT=1, DETECTION_MIN_CONFIDENCE=float(g["det_in_min_conf"]),
                                DETECTION_NMS_THRESHOLD=float(g["det_in_thr"]), DETECTION_MAX_INSTANCES=int(g["det_in_limits"][0]))
    me = types.SimpleNamespace(config=cfg, anchors=dev(g["prop_in_anchors"]))
    fgd = g["prop_in_fg_detied"]
    for algo in ("lazy", "mask", "hybrid"):
        ops.set_proposal_nms(algo)
        try:
            rois = ops.rpn_refine(me, dev(np.stack([1.0 - fgd, fgd], 1).astype(np.float32)).unsqueeze(0), bbox)
        finally:
            ops.set_proposal_nms("auto")
        want = g["prop_out_rois_detied"]
        assert tuple(rois.shape) == want.shape
        assert ulp_diff(rois.cpu().numpy(), want).max() <= 4           # torch.exp's rounding
    # the flow's own scores: the tie-break is the documented stable one (lowest anchor index first) = the oracle's
    rois = ops.rpn_refine(me, dev(g["rpn_out_class"]), bbox)
    np.testing.assert_array_equal(rois[0].cpu().numpy(), oracle.proposal_layer(g["rpn_out_class"][0], bbox[0].cpu().numpy(), g["prop_in_anchors"], pre,
                                                                              post, float(g["prop_in_thr"]), height=float(size), width=float(size)))
    # roi_align (model.py:276-393), both heads, on the reference's own RoIs
    fms = [dev(g[f"fm_{l}"]) for l in range(4)]
    for pool in (7, 14):
        for chl in (False, True):
            out = ops.roi_align([dev(g[f"pool{pool}_in_rois"])] + [cl(f) if chl else f for f in fms], pool, [size, size, 3])
            np.testing.assert_array_equal(out.cpu().numpy(), g[f"pool{pool}_out"])
    # mrn_refine (model.py:1389-1487)
    ci, sc, bx = ops.mrn_refine(me, dev(g["prop_out_rois"]), dev(g["det_in_probs"]), dev(g["det_in_deltas"]), g["det_in_window"])
    assert ci.dtype == torch.int64
    np.testing.assert_array_equal(ci.cpu().numpy(), g["det_out_class_ids"])
    np.testing.assert_array_equal(sc.cpu().numpy(), g["det_out_scores"])
    np.testing.assert_array_equal(bx.cpu().numpy(), g["det_out_boxes"])
    np.testing.assert_array_equal((bx.float() * 1.0 / size).cpu().numpy(), g["pool14_in_rois"])          # model.py:1188
    # full_masks (data.py:287-314): the reference's masks where it has any (it raises on the empty boxes), empty elsewhere
    h, w = (int(v) for v in g["mask_in_hw"])
    want, ok = detect_flow_expected_masks(g), g["mask_valid"]
    got = ops.full_masks(torch.zeros(len(ok), dtype=torch.int64, device="cuda"), bx[0], dev(g["mask_in_sel"]).unsqueeze(1), h, w).cpu().numpy()
    np.testing.assert_array_equal(got[ok], want[ok])
    assert not got[~ok].any()
    # decode_masks (data.py:265-284, model.py:1130): back to the 1920 x 1200 frame
    y1, x1, y2, x2 = (int(v) for v in g["det_in_window"])
    dec = ops.decode_masks(dev(want[ok]), float(g["decode_in_scale"]), (y2 - y1, x2 - x1)).cpu().numpy()
    assert dec.shape[1:] == tuple(int(v) for v in g["decode_out_hw"])
    import hashlib as _h
    assert _h.sha256(dec.tobytes()).digest() == g["decode_out_sha256"].tobytes()
    ops.check_device_errors()


def test_pyramid_arguments_are_validated(ops):
    """Shapes the kernels cannot see (one index per box, one (B, C) for the four levels, the preallocated gradient
    pyramid of the fused backward) are rejected on the host instead of being read out of bounds."""
    fms = [torch.zeros((2, 8, s, s), device="cuda") for s in (16, 8, 4, 2)]
    boxes = torch.tensor([[0.1, 0.1, 0.6, 0.6], [0.2, 0.3, 0.9, 0.8], [0.0, 0.0, 1.0, 1.0]], device="cuda")
    ind = torch.tensor([0, 1, 1], dtype=torch.int32, device="cuda")
    assert ops.pyramid_roi_align(fms, boxes, ind, 7, (64, 64, 3)).shape == (3, 8, 7, 7)
    with pytest.raises(ValueError):
        ops.pyramid_roi_align(fms, boxes, ind[:2], 7, (64, 64, 3))                       # an index short
    with pytest.raises(ValueError):
        ops.pyramid_roi_align(fms, boxes, ind.reshape(3, 1), 7, (64, 64, 3))
    with pytest.raises(TypeError):
        ops.pyramid_roi_align(fms, boxes, ind.long(), 7, (64, 64, 3))                    # __init__.py:34-35: int32
    with pytest.raises(ValueError):
        ops.pyramid_roi_align(fms[:3] + [torch.zeros((2, 4, 2, 2), device="cuda")], boxes, ind, 7, (64, 64, 3))
    with pytest.raises(ValueError):
        ops.pyramid_roi_align(fms[:3] + [torch.zeros((1, 8, 2, 2), device="cuda")], boxes, ind, 7, (64, 64, 3))
    with pytest.raises(ValueError):
        ops.pyramid_roi_align(fms[:3], boxes, ind, 7, (64, 64, 3))
    with pytest.raises(ValueError):
        ops.pyramid_roi_align_pair([cl(f) for f in fms], boxes, ind[:1], (7, 14), (64, 64, 3))
    g7, g14 = cl(torch.zeros((3, 8, 7, 7), device="cuda")), cl(torch.zeros((3, 8, 14, 14), device="cuda"))
    shapes = [tuple(f.shape) for f in fms]
    with pytest.raises(ValueError):
        ops.pyramid_roi_align_backward_pair(g7, g14, shapes, boxes, ind[:2], (64, 64, 3))
    with pytest.raises(ValueError):
        ops.pyramid_roi_align_backward_pair(g7, g14, shapes, boxes, ind, (64, 64, 3), out=[cl(f) for f in fms[:3]])
    with pytest.raises(ValueError):                                                      # NCHW gradient pyramid
        ops.pyramid_roi_align_backward_pair(g7, g14, shapes, boxes, ind, (64, 64, 3), out=[f.clone() for f in fms])
    out = ops.pyramid_roi_align_backward_pair(g7, g14, shapes, boxes, ind, (64, 64, 3), out=[cl(f) for f in fms])
    assert all(float(g.abs().max()) == 0.0 for g in out)
    ops.check_device_errors()


def test_deterministic_gather_plans(ops):
    """ops.set_deterministic(True): two backward calls (two plans of the same boxes) give bit-identical gradients, the fused
    two-head backward as well; the values agree with the default plan's to rounding and stay within the oracle's tolerance."""
    C, N, B = 64, 600, 3
    fms = [cl(torch.randn(B, C, s, s, device="cuda")) for s in (64, 32, 16, 8)]
    boxes_np = np.concatenate([synth.random_rois(N // B, 900 + i) for i in range(B)], 0)
    boxes_np[N // 2:N // 2 + 100] = boxes_np[N // 2]          # a hundred RoIs on one spot: long item lists in a few units
    boxes, ind = dev(boxes_np), torch.arange(B, device="cuda", dtype=torch.int32).repeat_interleave(N // B)

    def grads(pool):
        leaves = [f.clone().requires_grad_(True) for f in fms]
        out = ops.pyramid_roi_align(leaves, boxes, ind, pool, (256, 256, 3))
        g = torch.randn(out.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)).contiguous(memory_format=torch.channels_last)
        out.backward(g)
        return [l.grad.clone() for l in leaves]

    try:
        ops.set_deterministic(True)
        for pool in (7, 14):
            a, b = grads(pool), grads(pool)
            for x, y in zip(a, b):
                assert torch.equal(x, y)
        det = grads(14)
    finally:
        ops.set_deterministic(False)
    plain = grads(14)
    for x, y in zip(det, plain):       # a hundred coinciding RoIs: sums of hundreds of terms, so relative to the largest value
        assert float((x - y).abs().max()) <= 1e-5 * max(1.0, float(x.abs().max()))


def test_cooperative_kernels_in_a_cuda_graph(ops):
    """The fixed-point nms and the hybrid proposal NMS are cooperative launches: captured in a CUDA graph and replayed they give
    what the eager launches give (a serving loop replays the proposal layer from a graph)."""
    from maskrcnn_b200 import _lib as L
    n = 3000
    d = dev(_dets(n, 31))
    keep = torch.empty(n, dtype=torch.int64, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    ws = torch.empty(L.lib.mrcnn_nms_workspace_bytes(n), dtype=torch.uint8, device="cuda")
    anchors = synth.pyramid_anchors((256, 256))
    rcs, rbs = zip(*[synth.rpn_outputs(anchors, 40 + i, image=256.0, n_clusters=6) for i in range(3)])
    rc, rb, an = dev(np.stack(rcs)), dev(np.stack(rbs)), dev(anchors)
    want_keep = ops.nms(d, 0.6)
    want_rois, want_counts = ops.proposal_layer(rc, rb, an, 3000, 500, 0.7, image_hw=(256, 256))
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        L.check(L.lib.mrcnn_nms(d.data_ptr(), n, 0.6, keep.data_ptr(), cnt.data_ptr(), ws.data_ptr(), ws.numel(),
                                torch.cuda.current_stream().cuda_stream))
        rois, counts = ops.proposal_layer(rc, rb, an, 3000, 500, 0.7, image_hw=(256, 256))
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3):
        keep.zero_(); cnt.zero_(); rois.zero_(); counts.zero_()
        g.replay()
    torch.cuda.synchronize()
    k = int(cnt.item())
    assert k == want_keep.numel() and torch.equal(keep[:k], want_keep)
    assert torch.equal(rois, want_rois) and torch.equal(counts, want_counts)
