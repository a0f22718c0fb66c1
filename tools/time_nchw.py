"""NCHW-pyramid RoIAlign timings (the literal drop-in's layouts): configs[2] forward (1000 RoIs, one image) and the configs[3]
training geometry (16 x 512 RoIs), CUDA events, next to the algorithmic bytes.  python tools/time_nchw.py"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from maskrcnn_b200 import _lib as L, roofline, synth  # noqa: E402


def time_op(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def main():
    dev = "cuda"
    hbm, _ = bench.hbm_peak()
    B, R, C = 16, 512, 256
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    fm = [torch.randn((B, C, h, w), device=dev, generator=g) for h, w in bench.LEVEL_HW]
    gfm = [torch.empty_like(f) for f in fm]
    boxes_np = np.concatenate([synth.random_rois(R, 100 + i) for i in range(B)], 0)
    ind_np = np.repeat(np.arange(B, dtype=np.int32), R)
    boxes, ind = torch.from_numpy(boxes_np).to(dev), torch.from_numpy(ind_np).to(dev)
    N = B * R
    Hs, Ws = L.i4([h for h, _ in bench.LEVEL_HW]), L.i4([w for _, w in bench.LEVEL_HW])
    st = torch.cuda.current_stream().cuda_stream
    area = 1024.0 * 1024.0
    pf, pg = L.vp4([f.data_ptr() for f in fm]), L.vp4([f.data_ptr() for f in gfm])
    offs = L.i32_array([i * R for i in range(B + 1)])
    offs_p = ctypes.cast(offs, ctypes.c_void_p)
    res = {}
    pyr = B * bench.PYR_ELEMS_PER_IMAGE
    for pool in (7, 14):
        out = torch.empty((N, C, pool, pool), device=dev)
        grad = torch.randn((N, C, pool, pool), device=dev, generator=g)
        U, _ = roofline.unique_taps(boxes_np, ind_np, pool, (1024, 1024), bench.LEVEL_HW, B)
        fb = roofline.roialign_fwd_bytes(N, C, pool, U)
        bb = roofline.roialign_bwd_bytes(N, C, pool, pyr)
        t = time_op(lambda: L.check(L.lib.mrcnn_pyramid_roi_align_forward(pf, Hs, Ws, B, C, L.NCHW, boxes.data_ptr(), ind.data_ptr(), N, pool, area,
                                                                          out.data_ptr(), L.NCHW, None, st)))
        res["fwd%d" % pool] = {"ms": t * 1e3, "frac_of_hbm": fb / t / 1e9 / hbm}
        t = time_op(lambda: L.check(L.lib.mrcnn_pyramid_roi_align_backward(grad.data_ptr(), L.NCHW, Hs, Ws, B, C, boxes.data_ptr(), ind.data_ptr(), N,
                                                                           pool, area, pg, L.NCHW, 1, None, L.BWD_SCATTER, None, 0, st)))
        res["bwd%d_scatter" % pool] = {"ms": t * 1e3, "frac_of_hbm": bb / t / 1e9 / hbm}
        ws = torch.empty(L.lib.mrcnn_pyramid_roi_align_backward_workspace_bytes_ex(Hs, Ws, B, C, N, pool, L.NCHW), dtype=torch.uint8, device=dev)
        t = time_op(lambda: L.check(L.lib.mrcnn_pyramid_roi_align_backward(grad.data_ptr(), L.NCHW, Hs, Ws, B, C, boxes.data_ptr(), ind.data_ptr(), N,
                                                                           pool, area, pg, L.NCHW, 1, None, L.BWD_GATHER, ws.data_ptr(), ws.numel(), st)))
        res["bwd%d_gather" % pool] = {"ms": t * 1e3, "frac_of_hbm": bb / t / 1e9 / hbm}
        del ws
        t = time_op(lambda: L.check(L.lib.mrcnn_pyramid_roi_align_backward(grad.data_ptr(), L.NCHW, Hs, Ws, B, C, boxes.data_ptr(), None, N,
                                                                           pool, area, pg, L.NCHW, 1, offs_p, L.BWD_AUTO, None, 0, st)))
        res["bwd%d_image_by_image" % pool] = {"ms": t * 1e3, "frac_of_hbm": bb / t / 1e9 / hbm}
        # configs[2]: 1000 RoIs on image 0
        b1_np = synth.random_rois(1000, 1234)
        b1 = torch.from_numpy(b1_np).to(dev)
        o1 = torch.empty((1000, C, pool, pool), device=dev)
        U1, _ = roofline.unique_taps(b1_np, None, pool, (1024, 1024), bench.LEVEL_HW, 1)
        f1 = roofline.roialign_fwd_bytes(1000, C, pool, U1)
        p1 = L.vp4([f[:1].data_ptr() for f in fm])
        t = time_op(lambda: L.check(L.lib.mrcnn_pyramid_roi_align_forward(p1, Hs, Ws, 1, C, L.NCHW, b1.data_ptr(), None, 1000, pool, area,
                                                                          o1.data_ptr(), L.NCHW, None, st)), iters=50)
        res["config2_fwd%d" % pool] = {"us": t * 1e6, "frac_of_hbm": f1 / t / 1e9 / hbm}
        del out, grad
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
