"""Generates tests/golden/golden_pyr_wide_v1.npz by EXECUTING THE REFERENCE in the build container: the unmodified
model.roi_align (model.py:276-393) over the reference's compiled CPU extension (oracle/_ref), forward and autograd
backward, on a pyramid with C = 40 channels - a multiple of 4, so that on the GPU this reference-held vector reaches the
128-bit channel-vectorised kernels (golden_v1.npz's pyramid has C = 2, which only the strided kernels accept) with a
partially filled 64-channel chunk.  Run:  python tests/golden/make_golden_pyr_wide.py   (needs /root/reference).

The feature maps are 32^2 .. 4^2 while image_shape says 1024 x 1024: roi_align uses image_shape only for the level rule
(model.py:331-336), so the RoIs (16 .. 724 px of 1024) populate all four levels and the fixture stays small.  Inputs are
float16-representable and stored as float16; outputs are the reference's float32."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from maskrcnn_b200 import synth  # noqa: E402
from oracle import reference  # noqa: E402


def main():
    ref = reference.load()
    rng = np.random.default_rng(40)
    C, N, image = 40, 24, 1024
    sides = (32, 16, 8, 4)
    fms = [rng.standard_normal((1, C, s, s)).astype(np.float16) for s in sides]
    boxes = synth.random_rois(N, 41)
    boxes[0] = [0.0, 0.0, 1.0, 1.0]             # the whole image: every sample on the border rows / columns
    boxes[1] = [0.25, 0.25, 0.25, 0.75]         # zero height
    boxes[2] += 0.4                             # partly outside: extrapolated bins (crop_cpu.cpp:63-74)
    g = {"in_boxes": boxes, "in_image_shape": np.array([image, image, 3], np.int64)}
    for l, f in enumerate(fms):
        g[f"in_fm{l}"] = f
    for pool in (7, 14):
        t = [torch.from_numpy(f.astype(np.float32)).requires_grad_(True) for f in fms]
        with reference.quiet_stdout():
            out = ref.model.roi_align([torch.from_numpy(boxes).unsqueeze(0)] + t, pool, [image, image, 3])
        go = rng.standard_normal(tuple(out.shape)).astype(np.float16)
        out.backward(torch.from_numpy(go.astype(np.float32)))
        g[f"pool{pool}_out"] = out.detach().numpy()
        g[f"pool{pool}_in_grads"] = go
        for l in range(4):
            g[f"pool{pool}_out_gfm{l}"] = t[l].grad.numpy()
    lv = np.clip(np.round(4 + np.log2(np.sqrt((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]) + 1e-30) * image / 224.0)), 2, 5)
    print("levels", np.bincount(lv.astype(int), minlength=6)[2:])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_pyr_wide_v1.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
