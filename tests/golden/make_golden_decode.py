"""Generates tests/golden/golden_decode_v1.npz by EXECUTING THE REFERENCE's data.decode_masks (data.py:265-284; PIL
'1' -> 'L', torchvision CenterCrop + Resize through the installed Pillow) in the build container.
Run:  python tests/golden/make_golden_decode.py   (needs /root/reference).
Per case `<tag>_in_bits` = np.packbits of the bool masks [D,H,W], `<tag>_in_geom` = (D, H, W, window y1, x1, y2, x2),
`<tag>_in_scale` (float64), `<tag>_out` = the uint8 [D,nh,nw] the reference returned."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import reference  # noqa: E402
from test_oracle_vs_ref import _decode_case  # noqa: E402

CASES = {"a": (3, 256, 256, (48, 0, 208, 256), 256 / 1920, 11),     # predict.py geometry (1920x1200 frame) at 256
         "b": (4, 200, 200, (13, 0, 186, 200), 0.4161, 12),         # CenterCrop's half-even origin differs from the window's
         "c": (4, 128, 160, (0, 0, 128, 160), 2.0, 13),             # downscale by 2
         "d": (3, 120, 90, (10, 5, 111, 86), 3.3, 14),              # downscale by 3.3
         "e": (2, 96, 96, (0, 16, 96, 80), 0.75, 15)}


def main():
    d = reference.load().data
    g = {}
    for tag, (D, H, W, window, scale, seed) in CASES.items():
        m = _decode_case(D, H, W, seed)
        out = d.decode_masks(torch.from_numpy(m), scale, d.Box.fromlist(list(window))).numpy()
        g[f"{tag}_in_bits"] = np.packbits(m)
        g[f"{tag}_in_geom"] = np.array([D, H, W] + list(window), np.int64)
        g[f"{tag}_in_scale"] = np.float64(scale)
        g[f"{tag}_out"] = out
        print(tag, m.shape, "->", out.shape, "interpolated px", int(((out > 0) & (out < 255)).sum()))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_decode_v1.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
