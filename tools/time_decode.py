"""Times data.decode_masks' kernels on the bench case (100 pasted masks, 640x1024 window -> 1200x1920) and on smooth
'blob' masks (one disc per mask: few edge tiles).  usage: time_decode.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import maskrcnn_b200 as m
from maskrcnn_b200 import synth

IMAGE = 1024
cls, boxes, masks = synth.mask_head_outputs(100, 81, 41, image=IMAGE)
pasted = m.full_masks(*(torch.from_numpy(a).cuda() for a in (cls, boxes, masks)), IMAGE, IMAGE)
yy, xx = torch.meshgrid(torch.arange(IMAGE, device="cuda"), torch.arange(IMAGE, device="cuda"), indexing="ij")
g = torch.Generator(device="cuda"); g.manual_seed(1)
cy, cx, r = (torch.rand(100, generator=g, device="cuda") * s + o for s, o in ((400, 300), (600, 200), (150, 30)))
blobs = ((yy[None] - cy[:, None, None]) ** 2 + (xx[None] - cx[:, None, None]) ** 2) < (r ** 2)[:, None, None]
for name, src in (("bench masks (noisy checkerboards)", pasted), ("blob masks (one disc each)", blobs.contiguous())):
    f = lambda: m.decode_masks(src, IMAGE / 1920.0, (640, IMAGE))
    for _ in range(3):
        out = f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        f()
    b.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b) / 20 * 1e-3
    by = 100 * (640 * 1024 + 1200 * 1920)
    print("%-36s %.1f us  %.0f GB/s  checksum %d" % (name, t * 1e6, by / t / 1e9, int(out.long().sum())))
