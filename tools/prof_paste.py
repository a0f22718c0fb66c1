"""Runs the mask paste-back and the RPN head plumbing a few times at bench size (for ncu).  usage: prof_paste.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import maskrcnn_b200 as m
from maskrcnn_b200 import synth

dev = "cuda"
D, IMAGE = 800, 1024
cls, boxes, masks = synth.mask_head_outputs(100, 81, 41, image=IMAGE)
cls_d = torch.from_numpy(np.tile(cls, D // 100)).to(dev)
boxes_d = torch.from_numpy(np.tile(boxes, (D // 100, 1))).to(dev)
masks_d = torch.from_numpy(masks).to(dev).repeat(D // 100, 1, 1, 1)
sides = [IMAGE // s for s in (4, 8, 16, 32, 64)]
cls_l = [torch.randn((8, 6, s, s), device=dev) for s in sides]
box_l = [torch.randn((8, 12, s, s), device=dev) for s in sides]
for _ in range(3):
    m.full_masks(cls_d, boxes_d, masks_d, IMAGE, IMAGE)
    with torch.no_grad():
        m.rpn_pack(cls_l, box_l)
torch.cuda.synchronize()
if len(sys.argv) > 1 and sys.argv[1] == "time":
    out = torch.empty((D, IMAGE, IMAGE), dtype=torch.bool, device=dev)
    from maskrcnn_b200 import _lib as L
    s = torch.cuda.current_stream().cuda_stream
    ws = torch.empty(L.lib.mrcnn_full_masks_workspace_bytes(D, 28, 28, IMAGE, IMAGE), dtype=torch.uint8, device=dev)
    f = lambda: L.check(L.lib.mrcnn_full_masks(cls_d.data_ptr(), boxes_d.data_ptr(), masks_d.data_ptr(), D, 81, 28, 28, IMAGE, IMAGE,
                                               out.data_ptr(), ws.data_ptr(), ws.numel(), s))
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20 * 1e-3
    print("full_masks %.1f us  %.0f GB/s" % (t * 1e6, D * IMAGE * IMAGE / t / 1e9))
    boxes_d.zero_()
    for _ in range(3):
        f()
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20 * 1e-3
    print("full_masks, every box empty %.1f us  %.0f GB/s" % (t * 1e6, D * IMAGE * IMAGE / t / 1e9))
    z = lambda: out.zero_()
    z()
    e0.record()
    for _ in range(20):
        z()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20 * 1e-3
    print("memset of the same buffer %.1f us  %.0f GB/s" % (t * 1e6, D * IMAGE * IMAGE / t / 1e9))
print("done")
