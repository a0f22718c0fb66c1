"""The cooperative kernels (fixed-point nms, hybrid proposal NMS) inside a CUDA graph: capture, replay, compare with eager launches."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maskrcnn_b200 as m
from maskrcnn_b200 import _lib as L, synth

N = 3000
rng = np.random.default_rng(3)
b = synth.random_rois(N, 3, image=1024.0, min_size=16, max_size=400) * 1024.0
b[N // 2:] = b[:N - N // 2] + rng.uniform(-8, 8, (N - N // 2, 4)).astype(np.float32)
d5 = torch.from_numpy(np.concatenate([b, synth.unique_scores(N, 3)[:, None]], 1).astype(np.float32)).cuda()
keep = torch.empty(N, dtype=torch.int64, device="cuda"); cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
ws = torch.empty(L.lib.mrcnn_nms_workspace_bytes(N), dtype=torch.uint8, device="cuda")
anchors = synth.pyramid_anchors((256, 256))
rcs, rbs = zip(*[synth.rpn_outputs(anchors, 40 + i, image=256.0, n_clusters=6) for i in range(3)])
rc, rb, an = torch.from_numpy(np.stack(rcs)).cuda(), torch.from_numpy(np.stack(rbs)).cuda(), torch.from_numpy(anchors).cuda()
want_keep = m.nms(d5, 0.6)
want_rois, want_counts = m.proposal_layer(rc, rb, an, 3000, 500, 0.7, image_hw=(256, 256))
torch.cuda.synchronize()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    L.check(L.lib.mrcnn_nms(d5.data_ptr(), N, 0.6, keep.data_ptr(), cnt.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    rois, counts = m.proposal_layer(rc, rb, an, 3000, 500, 0.7, image_hw=(256, 256))
torch.cuda.current_stream().wait_stream(s)
for _ in range(3):
    keep.zero_(); cnt.zero_(); rois.zero_(); counts.zero_()
    g.replay()
torch.cuda.synchronize()
k = int(cnt.item())
ok = k == want_keep.numel() and torch.equal(keep[:k], want_keep) and torch.equal(rois, want_rois) and torch.equal(counts, want_counts)
print("graph replay of the cooperative kernels matches eager:", ok, "| kept", k, "| proposals", counts.tolist())
sys.exit(0 if ok else 1)
