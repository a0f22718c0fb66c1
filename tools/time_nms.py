"""Kernel time of the standalone nms (C ABI, no count read-back) for 6000 clustered boxes; MRCNN_NMS_THREADS / MRCNN_NMS_STAGED select
sweep variants.  python tools/time_nms.py [N]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskrcnn_b200 import _lib as L, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
rng = np.random.default_rng(11)
b = synth.random_rois(N, 11, image=1024.0, min_size=16, max_size=500) * 1024.0
b[N // 2:] = b[:N - N // 2] + rng.uniform(-8, 8, (N - N // 2, 4)).astype(np.float32)
d5 = torch.from_numpy(np.concatenate([b, np.sort(synth.unique_scores(N, 11))[::-1][:, None]], 1).astype(np.float32)).cuda()
keep = torch.empty(N, dtype=torch.int64, device="cuda")
cnt = torch.empty(1, dtype=torch.int32, device="cuda")
ws = torch.empty(L.lib.mrcnn_nms_workspace_bytes(N), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
f = lambda: L.check(L.lib.mrcnn_nms(d5.data_ptr(), N, 0.7, keep.data_ptr(), cnt.data_ptr(), ws.data_ptr(), ws.numel(), st))  # noqa: E731
for _ in range(5):
    f()
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    f()
e.record()
torch.cuda.synchronize()
print("threads=%s staged=%s  N=%d kept=%d  %.1f us" % (os.environ.get("MRCNN_NMS_THREADS", "dflt"), os.environ.get("MRCNN_NMS_STAGED", "dflt"), N,
                                                      int(cnt.item()), a.elapsed_time(e) / 50 * 1e3))
