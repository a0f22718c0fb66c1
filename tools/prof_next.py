"""Runs the SURVEY 8(f) kernels a few times each (for ncu): proposal layer with both NMS algorithms, rpn_pack, full_masks,
decode_masks, detection layer with both NMS algorithms.  usage: prof_next.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import maskrcnn_b200 as m
from maskrcnn_b200 import synth

dev = "cuda"
IMAGE = 1024
anchors = synth.pyramid_anchors((IMAGE, IMAGE))
rcs, rbs = zip(*[synth.rpn_outputs(anchors, 1235 + i) for i in range(2)])
rc = torch.from_numpy(np.stack([rcs[i % 2] for i in range(8)])).to(dev)
rb = torch.from_numpy(np.stack([rbs[i % 2] for i in range(8)])).to(dev)
an = torch.from_numpy(anchors).to(dev)
for algo in ("lazy", "mask"):
    m.set_proposal_nms(algo)
    for _ in range(2):
        rois, counts = m.proposal_layer(rc, rb, an, 6000, 1000, 0.7)
m.set_proposal_nms("auto")
torch.cuda.synchronize()
g = torch.Generator(device=dev)
g.manual_seed(9)
sides = [IMAGE // s for s in (4, 8, 16, 32, 64)]
cls_l = [torch.randn((8, 6, s, s), device=dev, generator=g) for s in sides]
box_l = [torch.randn((8, 12, s, s), device=dev, generator=g) for s in sides]
for _ in range(2):
    m.rpn_pack(cls_l, box_l)
cls, boxes, masks = synth.mask_head_outputs(100, 81, 41, image=IMAGE)
cls_d, boxes_d, masks_d = (torch.from_numpy(a).to(dev) for a in (cls, boxes, masks))
for _ in range(2):
    pasted = m.full_masks(cls_d, boxes_d, masks_d, IMAGE, IMAGE)
    dec = m.decode_masks(pasted, IMAGE / 1920.0, (640, IMAGE))
B, N, NC = 64, 1000, 81
rois = torch.from_numpy(np.stack([synth.random_rois(N, 300 + i) for i in range(B)])).to(dev)
probs = torch.softmax(3 * torch.randn(B, N, NC, device=dev, generator=g), -1)
deltas = 0.1 * torch.randn(B, N, NC, 4, device=dev, generator=g)
win = torch.tensor([[0, 0, IMAGE, IMAGE]], dtype=torch.float32, device=dev).repeat(B, 1)
for algo in ("lazy", "mask"):
    m.set_detection_nms(algo)
    for _ in range(2):
        dets, dc = m.detection_layer(rois, probs, deltas, win, 0.0, 0.3, 100)
m.set_detection_nms("auto")
torch.cuda.synchronize()
print("done", counts.tolist(), int(pasted.sum()), tuple(dec.shape), dc[:4].tolist())
