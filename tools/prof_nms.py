"""Runs the standalone nms drop-in a few times (for the ncu launch list).  usage: prof_nms.py [N]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maskrcnn_b200 as m  # noqa: E402
from maskrcnn_b200 import synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
rng = np.random.default_rng(11)
b = synth.random_rois(N, 11, image=1024.0, min_size=16, max_size=500) * 1024.0
b[N // 2:] = b[:N - N // 2] + rng.uniform(-8, 8, (N - N // 2, 4)).astype(np.float32)
d5 = torch.from_numpy(np.concatenate([b, np.sort(synth.unique_scores(N, 11))[::-1][:, None]], 1).astype(np.float32)).cuda()
for _ in range(3):
    k = m.nms(d5, 0.7)
torch.cuda.synchronize()
print("kept", k.numel())
