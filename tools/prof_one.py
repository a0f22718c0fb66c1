"""Runs one RoIAlign variant a few times (for ncu).  usage: prof_one.py <fwd|bwd|pair> <pool> <nchw|nhwc> [gather]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from maskrcnn_b200 import _lib as L

kind, pool, lay = sys.argv[1], int(sys.argv[2]), sys.argv[3]
gather = len(sys.argv) > 4 and sys.argv[4] == "gather"
perimage = len(sys.argv) > 4 and sys.argv[4] == "perimage"
ft = len(sys.argv) > 4 and sys.argv[4] == "ft"
wl = bench.Workload(torch, torch.device("cuda", 0))
cl = torch.channels_last
lay_id = L.NHWC if lay == "nhwc" else L.NCHW
buf = {7: (wl.out7, wl.g7, wl.gfm7), 14: (wl.out14, wl.g14, wl.gfm14)}[pool]
o, g, gf = buf
if lay == "nhwc":
    o, g = o.contiguous(memory_format=cl), g.contiguous(memory_format=cl)
ws = torch.empty(L.lib.mrcnn_pyramid_roi_align_backward_workspace_bytes(wl.Hs, wl.Ws, wl.batch, wl.N, pool), dtype=torch.uint8, device="cuda")
for _ in range(3):
    if kind == "pair":
        wl.fwd_pair()
    elif kind == "fwd":
        L.check(L.lib.mrcnn_pyramid_roi_align_forward(L.vp4([f.data_ptr() for f in wl.fm]), wl.Hs, wl.Ws, wl.batch, bench.CHANNELS, L.NHWC,
                                                      wl.boxes.data_ptr(), wl.ind.data_ptr(), wl.N, pool, wl.area, o.data_ptr(), lay_id, None, wl._s()))
    else:
        L.check(L.lib.mrcnn_pyramid_roi_align_backward(g.data_ptr(), lay_id, wl.Hs, wl.Ws, wl.batch, bench.CHANNELS, wl.boxes.data_ptr(),
                                                       wl.ind.data_ptr(), wl.N, pool, wl.area, L.vp4([x.data_ptr() for x in gf]), L.NHWC, 1,
                                                       wl.offsets if perimage else None, L.BWD_GATHER if gather else L.BWD_SCATTER, ws.data_ptr() if gather else None, ws.numel() if gather else 0, wl._s()))
torch.cuda.synchronize()
print("done")
