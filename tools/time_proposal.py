"""Proposal layer (configs[1]: 261,888 anchors, 6000 -> NMS 0.7 -> 1000, batch 8) with each NMS algorithm, CUDA events.
MRCNN_PROPOSAL_PREFIX_PCT sets the hybrid's prefix (percent of post_nms).  python tools/time_proposal.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maskrcnn_b200 as m
from maskrcnn_b200 import synth
anchors = synth.pyramid_anchors((1024, 1024))
def inputs(converge):
    rcs, rbs = zip(*[synth.rpn_outputs(anchors, 1235 + i, converge=converge) for i in range(2 if converge else 8)])
    n = len(rcs)
    return (torch.from_numpy(np.stack([rcs[i % n] for i in range(8)])).cuda(), torch.from_numpy(np.stack([rbs[i % n] for i in range(8)])).cuda())
an = torch.from_numpy(anchors).cuda()
def time_op(f, iters=30):
    for _ in range(5): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
for name, conv in (("bench inputs", 0.0), ("converged inputs", 0.9)):
    rc, rb = inputs(conv)
    for algo in ("hybrid", "lazy", "mask"):
        m.set_proposal_nms(algo)
        t = time_op(lambda: m.proposal_layer(rc, rb, an, 6000, 1000, 0.7))
        print("%-18s %-7s %8.1f us per batch of 8   kept %.0f" % (name, algo, t, float(m.proposal_layer(rc, rb, an, 6000, 1000, 0.7)[1].float().mean())))
m.set_proposal_nms("auto")
