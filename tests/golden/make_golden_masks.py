"""Generates tests/golden/golden_masks_v1.npz by EXECUTING THE REFERENCE's unmodified data.full_masks (data.py:287-314,
so the installed Pillow's resample) in the build container.  Run:  python tests/golden/make_golden_masks.py
(needs /root/reference).  `<tag>_in_*` inputs, `<tag>_out_bits` = np.packbits of the reference's bool [D,H,W] result."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from maskrcnn_b200 import synth  # noqa: E402
from oracle import reference  # noqa: E402

# tag: (detections, classes, image height, image width, min box, max box, seed)
CASES = {"a": (12, 81, 256, 256, 4, 200, 31), "b": (6, 5, 200, 333, 2, 40, 32), "c": (8, 81, 1024, 1024, 16, 800, 33)}


def inputs(tag):
    d, nc, h, w, lo, hi, seed = CASES[tag]
    cls, boxes, masks = synth.mask_head_outputs(d, nc, seed, image=min(h, w), min_size=lo, max_size=hi)
    if tag == "b":  # boxes that leave the image, fractional coordinates, unchanged width / height (pass skipped), saturated values
        boxes[0] = [-7.0, 300.0, 25.0, 350.0]
        boxes[1] = [10.5, 20.25, 38.5, 90.75]
        boxes[2] = [3.0, 5.0, 31.0, 120.0]
        boxes[3] = [100.0, 7.0, 190.0, 35.0]
        masks[4] = masks[4] * 1.5 - 0.25
    return cls, boxes, masks, h, w


def main():
    ref = reference.load()
    g = {}
    for tag in CASES:
        cls, boxes, masks, h, w = inputs(tag)
        out = ref.data.full_masks(torch.from_numpy(cls), torch.from_numpy(boxes), torch.from_numpy(masks), h, w).numpy()
        g[f"{tag}_in_cls"], g[f"{tag}_in_boxes"], g[f"{tag}_in_hw"] = cls, boxes, np.int32([h, w])
        sel = masks[np.arange(len(cls)), cls]   # only the selected class plane is read: keep the fixture small
        g[f"{tag}_in_masks_sel"] = sel
        g[f"{tag}_out_bits"] = np.packbits(out)
        print(tag, out.shape, "set pixels", int(out.sum()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_masks_v1.npz"), **g)


if __name__ == "__main__":
    main()
