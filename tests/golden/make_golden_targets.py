"""Generates tests/golden/golden_targets_v1.npz by EXECUTING THE REFERENCE's unmodified model.mrn_samples
(model.py:396-576) on CPU in the build container.  Run:  python tests/golden/make_golden_targets.py
(needs /root/reference; the GPU box only reads the .npz).

`<tag>_in_*` are inputs (incl. the torch seed and the two permutations torch.randperm produced under it, in the
reference's call order), `<tag>_out_*` what the reference returned."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from maskrcnn_b200 import synth  # noqa: E402
from oracle import reference  # noqa: E402

CASES = {  # tag: (n_rois, n_gt, n_crowd, n_pad, train_rois, image, seed)
    "a": (400, 10, 0, 0, 200, 128, 11),
    "b": (300, 9, 2, 2, 512, 128, 12),
    "c": (96, 4, 1, 0, 64, 96, 13),
}


def main():
    ref = reference.load()
    g = {}
    for tag, (n_rois, n_gt, n_crowd, n_pad, train_rois, image, seed) in CASES.items():
        rois, cls, gt, masks = synth.target_inputs(n_rois, n_gt, seed, image=image, n_crowd=n_crowd, n_pad=n_pad)
        cfg = types.SimpleNamespace(GPU_COUNT=0, TRAIN_ROIS_PER_IMAGE=train_rois, ROI_POSITIVE_RATIO=0.33,
                                    BBOX_STD_DEV=np.array([0.1, 0.1, 0.2, 0.2]), MASK_SHAPE=[28, 28])
        drawn = []
        real = torch.randperm

        def spy(n, *a, **k):
            p = real(n, *a, **k)
            drawn.append(p.numpy().copy())
            return p
        torch.manual_seed(500 + seed)
        torch.randperm = spy
        try:
            out = ref.model.mrn_samples(torch.from_numpy(rois)[None], torch.from_numpy(cls)[None], torch.from_numpy(gt)[None],
                                        torch.from_numpy(masks)[None], cfg)
        finally:
            torch.randperm = real
        assert len(drawn) == 2 and out[0].shape[0] > 10
        g[f"{tag}_in_rois"], g[f"{tag}_in_cls"], g[f"{tag}_in_gt"] = rois, cls, gt
        g[f"{tag}_in_masks"] = masks.astype(np.uint8)       # binary: stored compactly
        g[f"{tag}_in_train_rois"] = np.int32(train_rois)
        g[f"{tag}_in_seed"] = np.int32(500 + seed)
        g[f"{tag}_in_perm_pos"], g[f"{tag}_in_perm_neg"] = drawn
        g[f"{tag}_out_rois"], g[f"{tag}_out_cls"] = out[0].numpy(), out[1].numpy()
        g[f"{tag}_out_deltas"], g[f"{tag}_out_masks"] = out[2].numpy(), out[3].numpy().astype(np.uint8)
        print(tag, "rows", out[0].shape[0], "positives", int((out[1] > 0).sum()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_targets_v1.npz"), **g)


if __name__ == "__main__":
    main()
