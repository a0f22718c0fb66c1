"""Generates tests/golden/golden_rpn_v1.npz by EXECUTING THE REFERENCE's unmodified data.rpn_samples (data.py:449-591)
in the build container.  Run:  python tests/golden/make_golden_rpn.py   (needs /root/reference).
`<tag>_in_*` inputs (incl. the numpy seed), `<tag>_out_*` what the reference returned."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from maskrcnn_b200 import synth  # noqa: E402
from oracle import reference  # noqa: E402

CASES = {"a": (256, 7, 0, 64, 21), "b": (256, 10, 2, 256, 22), "c": (128, 2, 0, 32, 23)}  # image, n_gt, n_crowd, T, seed


def main():
    ref = reference.load()
    g = {}
    for tag, (image, n_gt, n_crowd, T, seed) in CASES.items():
        anchors = synth.pyramid_anchors((image, image)).astype(np.float64)
        cls, gt = synth.rpn_target_inputs(n_gt, seed, image=image, n_crowd=n_crowd)
        cfg = types.SimpleNamespace(RPN_TRAIN_ANCHORS_PER_IMAGE=T, RPN_BBOX_STD_DEV=np.array([0.1, 0.1, 0.2, 0.2]))
        np.random.seed(700 + seed)
        match, bbox = ref.data.rpn_samples(anchors, cls, gt, cfg)
        g[f"{tag}_in_image"], g[f"{tag}_in_cls"], g[f"{tag}_in_gt"] = np.int32(image), cls, gt
        g[f"{tag}_in_T"], g[f"{tag}_in_seed"] = np.int32(T), np.int32(700 + seed)
        g[f"{tag}_out_match"], g[f"{tag}_out_bbox"] = match.astype(np.int8), bbox
        print(tag, "anchors", len(anchors), "pos", int((match == 1).sum()), "neg", int((match == -1).sum()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_rpn_v1.npz"), **g)


if __name__ == "__main__":
    main()
