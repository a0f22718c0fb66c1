"""Times the batched proposal layer (BASELINE configs[1]) with both NMS algorithms on the bench inputs and on inputs whose
top boxes converge on a few objects (heavy suppression).  usage: time_proposal.py [ncu]  (ncu: a few untimed calls only)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import maskrcnn_b200 as m
from maskrcnn_b200 import synth

dev = "cuda"
anchors = synth.pyramid_anchors((1024, 1024))
an = torch.from_numpy(anchors).to(dev)


def inputs(converge):
    rcs, rbs = zip(*[synth.rpn_outputs(anchors, 1235 + i, converge=converge) for i in range(2)])
    return (torch.from_numpy(np.stack([rcs[i % 2] for i in range(8)])).to(dev),
            torch.from_numpy(np.stack([rbs[i % 2] for i in range(8)])).to(dev))


for converge in (0.0, 0.9):
    rc, rb = inputs(converge)
    for algo in ("lazy", "mask"):
        m.set_proposal_nms(algo)
        f = lambda: m.proposal_layer(rc, rb, an, 6000, 1000, 0.7)
        for _ in range(3):
            rois, counts = f()
        torch.cuda.synchronize()
        if len(sys.argv) > 1:
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            f()
        e1.record()
        torch.cuda.synchronize()
        print("converge %.1f  %-4s NMS: %.1f us per batch of 8, kept %s" % (converge, algo, e0.elapsed_time(e1) / 20 * 1e3, counts[:2].tolist()))
print("done")
