"""The reference's UNMODIFIED model.py executing on the GPU on top of this repo's drop-in (north_star: "model.py,
predict.py and train.sh use it as a drop-in").  baseline/_ref/ holds the reference's four Python files and the demo image
(copied by __graft_entry__.build(), git-ignored, shipped to the GPU box); tools/refmodel.py loads them.

Two ways of dropping in, both exercised for inference (predict.py:43-60 -> MaskRCNN.detect, model.py:1095-1138) and for a
training step (MaskRCNN.train_epoch, model.py:1579-1637: extract, mrn_samples, the five losses, backward):

  (i)  package only: model.py imports this repo's `maskrcnn` (nms + CropFunction) and runs ITS OWN roi_align / rpn_refine /
       mrn_refine / mrn_samples Python around them (NCHW feature maps, per-level crops, per-class NMS calls);
  (ii) maskrcnn_b200.patch(model, data): the fused layers, with the network in torch.channels_last.

Every call that crosses the operator boundary is recorded (the tensors the real network produced) and replayed through
the oracle on the CPU: selections and forward interpolation bit-exact, gradients <= 1e-5 of max |reference|.  Comparing
whole-network outputs between CPU and GPU is NOT a parity statement (cuDNN and MKL-DNN convolutions differ in the last
bits and random-init scores are saturated with ties), so parity is asserted per operator on identical inputs."""
import os
import sys
import types

import numpy as np
import pytest
import torch

import oracle
from helpers import SOFTMAX_TOL, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from tools import refmodel  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refmodel.available(), reason="baseline/_ref (the reference's model.py) is not here")]

SEED = 2026
BWD_TOL = 1e-5
T = lambda t: t.detach().cpu().numpy().copy()  # noqa: E731


@pytest.fixture(scope="module")
def ops():
    import maskrcnn_b200
    return maskrcnn_b200


def recording_package(product, rec):
    """A module with the reference package's two names (c++ext/maskrcnn/__init__.py) that calls `product` (this repo's
    `maskrcnn`) and records every call's tensors."""
    mod = types.ModuleType("maskrcnn")

    def nms(dets, threshold):
        keep = product.nms(dets, threshold)
        rec["nms"].append((T(dets), float(threshold), T(keep)))
        return keep

    class CropFunction(object):
        def __init__(self, crop_height, crop_width, extrapolation_value=0):
            self.args = (crop_height, crop_width, extrapolation_value)
            self.inner = product.CropFunction(crop_height, crop_width, extrapolation_value)

        def __call__(self, image, boxes, box_ind):
            e = {"args": self.args, "image": T(image), "boxes": T(boxes), "ind": T(box_ind), "is_cuda": image.is_cuda and out_is_cuda(image)}
            x = image
            if torch.is_grad_enabled() and image.requires_grad:
                x = image * 1                      # a private autograd edge: its gradient is this call's grad_image alone
                x.register_hook(lambda g, e=e: e.__setitem__("grad_image", T(g)))
            out = self.inner(x, boxes, box_ind)
            if out.requires_grad:
                out.register_hook(lambda g, e=e: e.__setitem__("grad_out", T(g)))
            e["out"] = T(out)
            rec["crop"].append(e)
            return out

    mod.nms, mod.CropFunction = nms, CropFunction
    return mod


def out_is_cuda(t):
    return t.device.type == "cuda"


def check_package_calls(rec, on_gpu=True):
    """Replays every recorded nms / CropFunction call through the oracle."""
    n_bwd = 0
    for dets, thr, keep in rec["nms"]:
        np.testing.assert_array_equal(keep, oracle.nms(dets, thr))
    for e in rec["crop"]:
        assert e["is_cuda"] == on_gpu
        h, w, ev = e["args"]
        want = oracle.crop_forward(e["image"], e["boxes"], e["ind"], h, w, float(ev))
        np.testing.assert_array_equal(e["out"], want)
        if "grad_out" in e:
            assert "grad_image" in e
            wg = oracle.crop_backward(e["grad_out"], e["boxes"], e["ind"], e["image"].shape)
            assert rel_err(e["grad_image"], wg) <= BWD_TOL
            n_bwd += 1
    return n_bwd


def new_rec():
    return {"nms": [], "crop": [], "roi_align": [], "rpn_refine": [], "mrn_refine": [], "mrn_samples": [], "rpn_detect": []}


def wrap_fused(ref, ops, rec):
    """maskrcnn_b200.patch() and recorders around the fused layers it installed."""
    ops.patch(ref.model, ref.data)
    fused_roi_align, fused_samples = ref.model.roi_align, ref.model.mrn_samples
    M = ref.model.MaskRCNN
    fused_refine, fused_mrn, fused_detect = M.rpn_refine, M.mrn_refine, M.rpn_detect

    def roi_align(inputs, pool_size, image_shape):
        boxes, fms = inputs[0], list(inputs[1:5])
        e = {"boxes": T(boxes), "fms": [T(f) for f in fms], "pool": int(pool_size), "shape": [int(v) for v in image_shape],
             "channels_last": all(f.is_contiguous(memory_format=torch.channels_last) for f in fms)}
        if torch.is_grad_enabled() and any(f.requires_grad for f in fms):
            fms = [f * 1 for f in fms]
            e["grad_fms"] = [None] * 4
            for l, f in enumerate(fms):
                f.register_hook(lambda g, e=e, l=l: e["grad_fms"].__setitem__(l, T(g)))
        out = fused_roi_align([boxes] + fms, pool_size, image_shape)
        if out.requires_grad:
            out.register_hook(lambda g, e=e: e.__setitem__("grad_out", T(g)))
        e["out"] = T(out)
        rec["roi_align"].append(e)
        return out

    def rpn_refine(self, rpn_class, rpn_bbox):
        out = fused_refine(self, rpn_class, rpn_bbox)
        rec["rpn_refine"].append((T(rpn_class), T(rpn_bbox), T(self.anchors), T(out)))
        return out

    def mrn_refine(self, rois, probs, deltas, window):
        out = fused_mrn(self, rois, probs, deltas, window)
        rec["mrn_refine"].append((T(rois), T(probs), T(deltas), np.asarray(window, np.float32),
                                  None if out[0] is None else tuple(T(o) for o in out)))
        return out

    def rpn_detect(self, fms):
        out = fused_detect(self, fms)
        with torch.no_grad():      # the reference's chain (model.py:624-641 per level, :1294-1304) on the same device
            lg, bb = [], []
            for p in fms:
                x = self.rpn.relu(self.rpn.conv_shared(self.rpn.padding(p)))
                lg.append(self.rpn.conv_class(x).permute(0, 2, 3, 1).contiguous().view(p.size(0), -1, 2))
                bb.append(self.rpn.conv_bbox(x).permute(0, 2, 3, 1).contiguous().view(p.size(0), -1, 4))
            lg, bb = torch.cat(lg, 1), torch.cat(bb, 1)
        rec["rpn_detect"].append((bool(torch.equal(out[0], lg)), bool(torch.equal(out[2], bb)),
                                  float((out[1] - torch.softmax(lg, 2)).abs().max())))
        return out

    def mrn_samples(rpn_rois, gt_class_ids, gt_boxes, gt_masks, config):
        state = torch.get_rng_state()
        out = fused_samples(rpn_rois, gt_class_ids, gt_boxes, gt_masks, config)
        after = torch.get_rng_state()
        torch.set_rng_state(state)                   # the oracle draws the reference's two permutations again
        want = oracle.mrn_samples(T(rpn_rois)[0], T(gt_class_ids)[0].astype(np.int32), T(gt_boxes)[0], T(gt_masks)[0],
                                  int(config.TRAIN_ROIS_PER_IMAGE), float(config.ROI_POSITIVE_RATIO),
                                  np.asarray(config.BBOX_STD_DEV, np.float32).reshape(4), tuple(config.MASK_SHAPE),
                                  lambda n: torch.randperm(n).numpy())
        torch.set_rng_state(after)
        rec["mrn_samples"].append(([T(o) for o in out], want))
        return out

    ref.model.roi_align, ref.model.mrn_samples = roi_align, mrn_samples
    M.rpn_refine, M.mrn_refine, M.rpn_detect = rpn_refine, mrn_refine, rpn_detect


def check_fused_calls(rec, cfg, expect_channels_last):
    size = int(cfg.IMAGE_SHAPE[0])
    n_bwd = 0
    for ok_logits, ok_bbox, softmax_err in rec["rpn_detect"]:
        assert ok_logits and ok_bbox and softmax_err <= SOFTMAX_TOL
    for rpn_class, rpn_bbox, anchors, out in rec["rpn_refine"]:
        want = oracle.proposal_layer(rpn_class[0], rpn_bbox[0], anchors, min(500, len(anchors)), int(cfg.RPN_NMS_MAX_ROIS_NUM),
                                     float(cfg.RPN_NMS_THRESHOLD), std=np.asarray(cfg.RPN_BBOX_STD_DEV, np.float32).reshape(4),
                                     height=float(size), width=float(size))
        assert out.shape == (1,) + want.shape
        np.testing.assert_array_equal(out[0], want)
    for e in rec["roi_align"]:
        assert e["channels_last"] == expect_channels_last
        area = float(e["shape"][0] * e["shape"][1])
        boxes = e["boxes"].reshape(-1, 4)
        want, _ = oracle.pyramid_roi_align_fwd(e["fms"], boxes, None, e["pool"], area)
        np.testing.assert_array_equal(e["out"], want)
        if "grad_out" in e:
            wg = oracle.pyramid_roi_align_bwd(e["grad_out"], [f.shape for f in e["fms"]], boxes, None, area)
            for l in range(4):
                assert e["grad_fms"][l] is not None
                assert rel_err(e["grad_fms"][l], wg[l]) <= BWD_TOL
            n_bwd += 1
    for rois, probs, deltas, window, out in rec["mrn_refine"]:
        want = oracle.detection_layer(rois.reshape(-1, 4), probs, deltas, window, float(cfg.DETECTION_MIN_CONFIDENCE or 0.0),
                                      float(cfg.DETECTION_NMS_THRESHOLD), int(cfg.DETECTION_MAX_INSTANCES),
                                      std=np.asarray(cfg.RPN_BBOX_STD_DEV, np.float32).reshape(4), height=float(size), width=float(size))
        if out is None:
            assert len(want) == 0
            continue
        class_ids, scores, boxes = out
        np.testing.assert_array_equal(boxes[0], want[:, :4])
        np.testing.assert_array_equal(scores[0], want[:, 4])
        np.testing.assert_array_equal(class_ids[0], want[:, 5].astype(np.int64))
    for got, want in rec["mrn_samples"]:
        for a, b in zip(got, want):
            np.testing.assert_array_equal(a, b)
    return n_bwd


def demo_image():
    p = refmodel.image_path()
    assert p is not None, "baseline/_ref/images/car58a54312d.jpg is missing"
    return refmodel.pil_imread(p)


# ------------------------------------------------------------------ inference: BASELINE configs[0]
def test_predict_flow_package_only(ops):
    """predict.py:43-60 on the GPU with only the `maskrcnn` package swapped: the reference's own Python issues one nms per
    class and one CropFunction per populated level; all of them land on this repo's kernels with NCHW tensors."""
    rec = new_rec()
    import maskrcnn as product
    ref = refmodel.load(recording_package(product, rec))
    refmodel.tolerate_empty_boxes(ref)
    cfg = refmodel.make_config(ref, gpu=True)
    model = refmodel.make_model(ref, cfg, SEED)
    with torch.no_grad():
        class_ids, scores, boxes, masks = model.detect(demo_image())
    assert class_ids is not None and len(class_ids) > 0
    assert len(rec["nms"]) >= 2 and len(rec["crop"]) >= 2           # RPN nms + per-class nms; 7x7 and 14x14 crops
    assert {e["args"][0] for e in rec["crop"]} == {7, 14}
    assert rec["nms"][0][0].shape[0] == 500                         # model.py:1345 hard-codes the pre-NMS 500
    check_package_calls(rec)
    ops.check_device_errors()


def test_predict_flow_patched_channels_last(ops):
    """The same flow with maskrcnn_b200.patch() and the network in channels_last: fused rpn_detect,
    rpn_refine, roi_align x2, mrn_refine, full_masks, decode_masks - every stage checked against the oracle on the tensors
    the network produced."""
    rec = new_rec()
    import maskrcnn as product
    ref = refmodel.load(product)
    wrap_fused(ref, ops, rec)
    cfg = refmodel.make_config(ref, gpu=True)
    model = refmodel.make_model(ref, cfg, SEED).to(memory_format=torch.channels_last)
    full_calls = []
    fused_full, fused_decode = ref.data.full_masks, ref.data.decode_masks

    def full_masks(class_ids, boxes, masks, h, w):
        out = fused_full(class_ids, boxes, masks, h, w)
        full_calls.append((T(class_ids), T(boxes), T(masks), h, w, T(out)))
        return out
    ref.data.full_masks = full_masks
    with torch.no_grad():
        class_ids, scores, boxes, masks = model.detect(demo_image())
    assert class_ids is not None and len(class_ids) > 0
    assert len(rec["rpn_refine"]) == 1 and len(rec["mrn_refine"]) == 1 and [e["pool"] for e in rec["roi_align"]] == [7, 14]
    check_fused_calls(rec, cfg, expect_channels_last=True)
    assert len(full_calls) == 1
    cid, bx, mk, h, w, out = full_calls[0]
    ok = ((bx[:, 2] - bx[:, 0]).astype(np.int64) > 0) & ((bx[:, 3] - bx[:, 1]).astype(np.int64) > 0)
    want = oracle.full_masks(cid[ok].astype(np.int64), bx[ok], mk[ok], h, w)
    np.testing.assert_array_equal(out[ok].astype(bool), want.astype(bool))
    assert not out[~ok].any()
    assert np.asarray(masks).shape[1:] == (1200, 1920)              # decode_masks: back to the original frame
    ops.check_device_errors()


# ------------------------------------------------------------------ training step
def _train_step(ref, cfg, model, rec):
    """One sample through MaskRCNN.train_epoch.  The ground truth is made of the network's own proposals for the image
    (refmodel.proposals_of: an untrained RPN proposes nothing that overlaps a random ground truth), found by a dry forward
    pass whose recorded calls are dropped."""
    probe = refmodel.train_inputs(ref, cfg, 5)
    gt = refmodel.gt_from_proposals(refmodel.proposals_of(model, probe[0]))
    for v in rec.values():
        del v[:]
    inputs = refmodel.train_inputs(ref, cfg, 5, gt_boxes=gt)
    opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=1e-3, momentum=0.9)
    torch.manual_seed(3)
    loss = model.train_epoch([inputs], opt, 1)
    assert np.isfinite(loss)
    return loss


def test_train_step_package_only(ops):
    """MaskRCNN.train_epoch (model.py:1579-1637) for one sample with only the package swapped: forward, the reference's own
    mrn_samples (28x28 mask-target crop through CropFunction), the five losses, loss.backward() through CropFunction's
    autograd for both heads and every populated level."""
    rec = new_rec()
    import maskrcnn as product
    ref = refmodel.load(recording_package(product, rec))
    cfg = refmodel.make_config(ref, gpu=True, train=True)
    model = refmodel.make_model(ref, cfg, SEED)
    _train_step(ref, cfg, model, rec)
    assert {e["args"][0] for e in rec["crop"]} == {7, 14, 28}
    assert check_package_calls(rec) >= 2                            # at least one level per head received a gradient
    ops.check_device_errors()


def test_train_step_patched_channels_last(ops):
    """The same training step with patch() + channels_last: fused proposal layer, detection-target layer, and ONE RoIAlign
    node per head whose backward is the gather kernel; gradients on P2..P5 against the oracle's scatter."""
    rec = new_rec()
    import maskrcnn as product
    ref = refmodel.load(product)
    wrap_fused(ref, ops, rec)
    cfg = refmodel.make_config(ref, gpu=True, train=True)
    model = refmodel.make_model(ref, cfg, SEED).to(memory_format=torch.channels_last)
    _train_step(ref, cfg, model, rec)
    assert [e["pool"] for e in rec["roi_align"]] == [7, 14] and len(rec["mrn_samples"]) == 1
    assert check_fused_calls(rec, cfg, expect_channels_last=True) == 2
    ops.check_device_errors()
