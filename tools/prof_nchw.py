"""Runs the NCHW-pyramid RoIAlign kernels a few times at the configs[3] geometry (for ncu).  usage: prof_nchw.py <pool>"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from maskrcnn_b200 import _lib as L, synth  # noqa: E402

pool = int(sys.argv[1]) if len(sys.argv) > 1 else 14
dev = "cuda"
B, R, C = 16, 512, 256
g = torch.Generator(device=dev)
g.manual_seed(1)
fm = [torch.randn((B, C, h, w), device=dev, generator=g) for h, w in bench.LEVEL_HW]
gfm = [torch.empty_like(f) for f in fm]
boxes = torch.from_numpy(np.concatenate([synth.random_rois(R, 100 + i) for i in range(B)], 0)).to(dev)
ind = torch.arange(B, dtype=torch.int32, device=dev).repeat_interleave(R)
N = B * R
Hs, Ws = L.i4([h for h, _ in bench.LEVEL_HW]), L.i4([w for _, w in bench.LEVEL_HW])
st = torch.cuda.current_stream().cuda_stream
out = torch.empty((N, C, pool, pool), device=dev)
grad = torch.randn((N, C, pool, pool), device=dev, generator=g)
for _ in range(2):
    L.check(L.lib.mrcnn_pyramid_roi_align_forward(L.vp4([f.data_ptr() for f in fm]), Hs, Ws, B, C, L.NCHW, boxes.data_ptr(), ind.data_ptr(), N, pool,
                                                  1024.0 * 1024.0, out.data_ptr(), L.NCHW, None, st))
    L.check(L.lib.mrcnn_pyramid_roi_align_backward(grad.data_ptr(), L.NCHW, Hs, Ws, B, C, boxes.data_ptr(), ind.data_ptr(), N, pool, 1024.0 * 1024.0,
                                                   L.vp4([f.data_ptr() for f in gfm]), L.NCHW, 1, None, L.BWD_AUTO, None, 0, st))
torch.cuda.synchronize()
print("done")
