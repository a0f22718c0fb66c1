"""Generates tests/golden/golden_rpnhead_v1.npz by EXECUTING THE REFERENCE's unmodified RPN module (model.py:573-653) and the
concatenation of MaskRCNN.rpn_detect (model.py:1294-1304) on the CPU in the build container.
Run:  python tests/golden/make_golden_rpnhead.py   (needs /root/reference).
`in_logits_<l>` / `in_bbox_<l>` = the outputs of the module's own conv_class / conv_bbox per level (the inputs of the
plumbing this repo replaces), `out_*` = what RPN.forward + torch.cat returned."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference  # noqa: E402

SIDES = (24, 12, 6, 3, 2)
BATCH, DEPTH, K = 2, 8, 3


def main():
    ref = reference.load()
    torch.manual_seed(77)
    rpn = ref.model.RPN(K, 1, DEPTH)
    feats = [torch.randn(BATCH, DEPTH, s, s + 1) * 2.0 for s in SIDES]
    g = {}
    with torch.no_grad():
        outs = [rpn(f) for f in feats]
        logits, cls, bbox = [torch.cat(list(o), dim=1) for o in zip(*outs)]
        for l, f in enumerate(feats):
            x = rpn.relu(rpn.conv_shared(rpn.padding(f)))
            g[f"in_logits_{l}"] = rpn.conv_class(x).numpy()
            g[f"in_bbox_{l}"] = rpn.conv_bbox(x).numpy()
    g["out_logits"], g["out_class"], g["out_bbox"] = logits.numpy(), cls.numpy(), bbox.numpy()
    print("anchors", logits.shape[1])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_rpnhead_v1.npz"), **g)


if __name__ == "__main__":
    main()
