"""Times the detection layer (BASELINE configs[4] sizes: 64 images x 1000 RoIs x 81 classes, NMS 0.3, top-100) with both NMS
algorithms.  usage: time_detection.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import maskrcnn_b200 as m
from maskrcnn_b200 import synth

B, N, NC = 64, 1000, 81
rois = torch.from_numpy(np.stack([synth.random_rois(N, 300 + i) for i in range(B)])).cuda()
g = torch.Generator(device="cuda")
g.manual_seed(7)
probs = torch.softmax(3 * torch.randn(B, N, NC, device="cuda", generator=g), -1)
deltas = 0.1 * torch.randn(B, N, NC, 4, device="cuda", generator=g)
win = torch.tensor([[0, 0, 1024, 1024]], dtype=torch.float32, device="cuda").repeat(B, 1)
for algo in ("lazy", "mask"):
    m.set_detection_nms(algo)
    f = lambda: m.detection_layer(rois, probs, deltas, win, 0.0, 0.3, 100)
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        f()
    b.record()
    torch.cuda.synchronize()
    print("%-4s NMS: detection layer %.1f us / 64 images" % (algo, a.elapsed_time(b) / 50 * 1e3))
