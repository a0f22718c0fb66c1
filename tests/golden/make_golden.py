"""Generates tests/golden/golden_v1.npz by EXECUTING THE REFERENCE in the build container:
the reference's compiled CPU extension (oracle/_ref) and its unmodified model.py (oracle/reference.py).
Run:  python tests/golden/make_golden.py      (needs /root/reference; the GPU box only reads the .npz)

Every array named `*_in_*` is an input, `*_out_*` is what the reference returned for it.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from maskrcnn_b200 import synth  # noqa: E402
from oracle import reference  # noqa: E402


def main():
    ref = reference.load()
    C = reference.ref_C()
    g = {}
    rng = np.random.default_rng(2026)

    # --- nms (nms_cpu.cpp) ---
    for tag, n, thr in (("a", 300, 0.5), ("b", 130, 0.7), ("c", 64, 0.3)):
        b = synth.random_rois(n, 100 + n, image=512.0, min_size=8, max_size=300) * 512.0
        b[n // 2:] = b[: n - n // 2] + rng.uniform(-5, 5, (n - n // 2, 4)).astype(np.float32)
        dets = np.concatenate([b, synth.unique_scores(n, n)[:, None]], 1).astype(np.float32)
        g[f"nms_{tag}_in_dets"] = dets
        g[f"nms_{tag}_in_thr"] = np.float32(thr)
        g[f"nms_{tag}_out_keep"] = C.nms(torch.from_numpy(dets), thr).numpy()

    # --- crop_and_resize fwd/bwd (crop_cpu.cpp) ---
    for tag, (B, Cc, H, W, N, ch, cw, ev) in (("a", (2, 3, 24, 20, 12, 7, 7, 0.0)), ("b", (3, 1, 64, 64, 6, 28, 28, 0.0)),
                                              ("c", (1, 2, 9, 9, 5, 1, 3, -2.0))):
        img = rng.standard_normal((B, Cc, H, W), dtype=np.float32)
        boxes = synth.random_rois(N, 7 + N, image=64.0, min_size=4, max_size=60)
        boxes[0] += 0.3
        boxes[1] -= 0.2
        boxes[2] = [0.0, 0.0, 1.0, 1.0]
        ind = rng.integers(0, B, N).astype(np.int32)
        out = torch.zeros(1)
        with reference.quiet_stdout():
            C.crop_forward(torch.from_numpy(img), torch.from_numpy(boxes), torch.from_numpy(ind), ev, ch, cw, out)
        go = rng.standard_normal(tuple(out.shape), dtype=np.float32)
        gi = torch.empty(img.shape)
        C.crop_backward(torch.from_numpy(go), torch.from_numpy(boxes), torch.from_numpy(ind), gi)
        g[f"crop_{tag}_in_image"] = img
        g[f"crop_{tag}_in_boxes"] = boxes
        g[f"crop_{tag}_in_ind"] = ind
        g[f"crop_{tag}_in_ev"] = np.float32(ev)
        g[f"crop_{tag}_out_crops"] = out.numpy()
        g[f"crop_{tag}_in_grads"] = go
        g[f"crop_{tag}_out_gimage"] = gi.numpy()

    # --- PyramidROIAlign fwd + bwd through model.roi_align (model.py:276-393) ---
    size, Cc, N = 512, 2, 48
    fms = synth.feature_pyramid(1, Cc, 31, image=size)
    boxes = synth.random_rois(N, 32, image=float(size), min_size=8, max_size=size * 0.9)
    for pool in (7, 14):
        t = [torch.from_numpy(f).clone().requires_grad_(True) for f in fms]
        out = ref.model.roi_align([torch.from_numpy(boxes).unsqueeze(0)] + t, pool, [size, size, 3])
        go = rng.standard_normal(tuple(out.shape), dtype=np.float32)
        out.backward(torch.from_numpy(go))
        g[f"pyr{pool}_out"] = out.detach().numpy()
        g[f"pyr{pool}_in_grads"] = go
        for l in range(4):
            g[f"pyr{pool}_out_gfm{l}"] = t[l].grad.numpy()
    for l in range(4):
        g[f"pyr_in_fm{l}"] = fms[l]
    g["pyr_in_boxes"] = boxes
    g["pyr_in_image_size"] = np.int32(size)

    # --- proposal layer through MaskRCNN.rpn_refine (model.py:1307-1382); fork hard-codes pre_nms=500 ---
    anchors = synth.pyramid_anchors((256, 256))
    rc, rb = synth.rpn_outputs(anchors, 11, image=256.0, n_clusters=6)
    cfg = types.SimpleNamespace(RPN_NMS_MAX_ROIS_NUM=200, RPN_NMS_THRESHOLD=0.7, RPN_BBOX_STD_DEV=[0.1, 0.1, 0.2, 0.2],
                                IMAGE_SHAPE=np.array([256, 256, 3]), GPU_COUNT=0)
    stub = types.SimpleNamespace(config=cfg, anchors=torch.from_numpy(anchors))
    rois = ref.model.MaskRCNN.rpn_refine(stub, torch.from_numpy(rc).unsqueeze(0), torch.from_numpy(rb).unsqueeze(0))
    g["prop_in_rpn_class"] = rc
    g["prop_in_rpn_bbox"] = rb
    g["prop_in_image_size"] = np.int32(256)
    g["prop_in_limits"] = np.array([500, 200], np.int64)
    g["prop_out_rois"] = rois[0].numpy()

    # --- detection layer through MaskRCNN.mrn_refine (model.py:1389-1487) ---
    N, NC = 200, 81
    rois = synth.random_rois(N, 41)
    probs, deltas = synth.head_outputs(N, NC, 42)
    window = np.array([0, 0, 1024, 1024], np.float32)
    cfg = types.SimpleNamespace(RPN_BBOX_STD_DEV=np.array([0.1, 0.1, 0.2, 0.2]), IMAGE_SHAPE=np.array([1024, 1024, 3]),
                                GPU_COUNT=0, DETECTION_MIN_CONFIDENCE=0, DETECTION_NMS_THRESHOLD=0.3,
                                DETECTION_MAX_INSTANCES=100)
    ci, sc, bx = ref.model.MaskRCNN.mrn_refine(types.SimpleNamespace(config=cfg), torch.from_numpy(rois).unsqueeze(0),
                                               torch.from_numpy(probs), torch.from_numpy(deltas), window)
    g["det_in_rois"] = rois
    g["det_in_probs"] = probs.astype(np.float16).astype(np.float32) if False else probs
    g["det_in_deltas"] = deltas
    g["det_in_window"] = window
    g["det_out"] = np.concatenate([bx[0].numpy(), sc[0].numpy()[:, None], ci[0].numpy()[:, None].astype(np.float32)], 1)

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
