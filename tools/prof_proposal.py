"""Runs the batched proposal layer (BASELINE configs[1]) and the detection layer (configs[4]) a few times (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import maskrcnn_b200 as m
from maskrcnn_b200 import synth

dev = "cuda"
anchors = synth.pyramid_anchors((1024, 1024))
rcs, rbs = zip(*[synth.rpn_outputs(anchors, 1235 + i) for i in range(8)])
rc, rb, an = torch.from_numpy(np.stack(rcs)).to(dev), torch.from_numpy(np.stack(rbs)).to(dev), torch.from_numpy(anchors).to(dev)
for _ in range(3):
    rois, counts = m.proposal_layer(rc, rb, an, 6000, 1000, 0.7)
torch.cuda.synchronize()
print("kept", counts.tolist())
B, N, NC = 64, 1000, 81
rois = torch.from_numpy(np.stack([synth.random_rois(N, 300 + i) for i in range(B)])).to(dev)
g = torch.Generator(device=dev); g.manual_seed(7)
probs = torch.softmax(3 * torch.randn(B, N, NC, device=dev, generator=g), -1)
deltas = 0.1 * torch.randn(B, N, NC, 4, device=dev, generator=g)
win = torch.tensor([[0, 0, 1024, 1024]], dtype=torch.float32, device=dev).repeat(B, 1)
for _ in range(3):
    dets, dc = m.detection_layer(rois, probs, deltas, win, 0.0, 0.3, 100)
torch.cuda.synchronize()
print("dets", dc[:8].tolist())
