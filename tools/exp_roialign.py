"""Scratch timing of RoIAlign variants on the bench workload (not part of the product)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from maskrcnn_b200 import _lib as L

wl = bench.Workload(torch, torch.device("cuda", 0))
cl = torch.channels_last
g7c, g14c = wl.g7.contiguous(memory_format=cl), wl.g14.contiguous(memory_format=cl)
o7c, o14c = wl.out7.contiguous(memory_format=cl), wl.out14.contiguous(memory_format=cl)
wl.g7n, wl.g14n = wl.g7.contiguous(), wl.g14.contiguous()
o7n, o14n = wl.out7.contiguous(), wl.out14.contiguous()


ws = torch.empty(L.lib.mrcnn_pyramid_roi_align_backward_workspace_bytes(wl.Hs, wl.Ws, wl.batch, wl.N, 14), dtype=torch.uint8, device="cuda")


def bwd(pool, g, gl, gfm, offs, gather=False, ft=False):
    L.check(L.lib.mrcnn_pyramid_roi_align_backward(g.data_ptr(), gl, wl.Hs, wl.Ws, wl.batch, bench.CHANNELS, wl.boxes.data_ptr(),
                                                   wl.ind.data_ptr(), wl.N, pool, wl.area, L.vp4([x.data_ptr() for x in gfm]),
                                                   L.NHWC, 1, offs, L.BWD_GATHER if gather else L.BWD_SCATTER, ws.data_ptr() if gather else None, ws.numel() if gather else 0, wl._s()))


def fwd(pool, o, ol):
    L.check(L.lib.mrcnn_pyramid_roi_align_forward(L.vp4([f.data_ptr() for f in wl.fm]), wl.Hs, wl.Ws, wl.batch, bench.CHANNELS, L.NHWC,
                                                  wl.boxes.data_ptr(), wl.ind.data_ptr(), wl.N, pool, wl.area, o.data_ptr(), ol, None, wl._s()))


def zero_only():
    for g in wl.gfm14:
        g.zero_()


cases = {
    "fwd7  nchw": lambda: fwd(7, o7n, L.NCHW), "fwd7  nhwc": lambda: fwd(7, o7c, L.NHWC),
    "fwd14 nchw": lambda: fwd(14, o14n, L.NCHW), "fwd14 nhwc": lambda: fwd(14, o14c, L.NHWC),
    "bwd7  nchw all": lambda: bwd(7, wl.g7n, L.NCHW, wl.gfm7, None), "bwd7  nchw per-image": lambda: bwd(7, wl.g7n, L.NCHW, wl.gfm7, wl.offsets),
    "bwd7  nhwc all": lambda: bwd(7, g7c, L.NHWC, wl.gfm7, None), "bwd7  nhwc per-image": lambda: bwd(7, g7c, L.NHWC, wl.gfm7, wl.offsets),
    "bwd14 nchw all": lambda: bwd(14, wl.g14n, L.NCHW, wl.gfm14, None), "bwd14 nchw per-image": lambda: bwd(14, wl.g14n, L.NCHW, wl.gfm14, wl.offsets),
    "bwd14 nhwc all": lambda: bwd(14, g14c, L.NHWC, wl.gfm14, None), "bwd14 nhwc per-image": lambda: bwd(14, g14c, L.NHWC, wl.gfm14, wl.offsets),
    "bwd7  nhwc GATHER": lambda: bwd(7, g7c, L.NHWC, wl.gfm7, None, True), "bwd14 nhwc GATHER": lambda: bwd(14, g14c, L.NHWC, wl.gfm14, None, True),
    "torch zero_ of one pyramid": zero_only,
}
only = sys.argv[1].lower() if len(sys.argv) > 1 else None
for name, fn in cases.items():
    if only and only not in name.lower():
        continue
    t = wl.time_op(fn, iters=20)
    print("%-28s %8.1f us" % (name, t * 1e6))
