"""Both heads' forward: two launches (7x7, 14x14) against the fused launch, configs[3] geometry; checks bit-identity."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

wl = bench.Workload(torch, torch.device("cuda", 0))
L = wl.L
o7 = torch.empty_like(wl.out7)
o14 = torch.empty_like(wl.out14)


def pair():
    L.check(L.lib.mrcnn_pyramid_roi_align_forward_pair(L.vp4([f.data_ptr() for f in wl.fm]), wl.Hs, wl.Ws, wl.batch, bench.CHANNELS,
                                                       wl.boxes.data_ptr(), wl.ind.data_ptr(), wl.N, wl.area, o7.data_ptr(), o14.data_ptr(), wl._s()))


def two():
    wl.fwd(7, wl.out7)
    wl.fwd(14, wl.out14)


t2 = wl.time_op(two, iters=30, warm=5)
tp = wl.time_op(pair, iters=30, warm=5)
print("two launches %.4f ms   fused %.4f ms   identical: %s %s" % (t2 * 1e3, tp * 1e3, bool(torch.equal(o7, wl.out7)), bool(torch.equal(o14, wl.out14))))
