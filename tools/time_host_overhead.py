"""Host-side cost of the small drop-in calls (configs[2]: 1000 RoIs on one image - kernels of 20-60 us): wall-clock per call
when calls are issued back to back without synchronising (the host is the bottleneck when this exceeds the kernel time), and
the kernel time from CUDA events.  Run on the GPU box:  python tools/time_host_overhead.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maskrcnn_b200 as m  # noqa: E402
from maskrcnn_b200 import _lib as L, synth  # noqa: E402


def host_us(fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6


def main():
    dev = "cuda"
    torch.manual_seed(0)
    fms_cl = [torch.randn(1, 256, s, s, device=dev).contiguous(memory_format=torch.channels_last) for s in (256, 128, 64, 32)]
    fms_nc = [f.contiguous() for f in fms_cl]
    boxes = torch.from_numpy(synth.random_rois(1000, 1234)).to(dev)
    Hs, Ws = L.i4([256, 128, 64, 32]), L.i4([256, 128, 64, 32])
    st = torch.cuda.current_stream().cuda_stream
    print("%-62s %10s %10s" % ("call", "issue us", "total us"))
    for pool in (7, 14):
        for name, fms, lay in (("channels-last pyramid", fms_cl, L.NHWC), ("NCHW pyramid", fms_nc, L.NCHW)):
            for ocl in (True, False):
                out = torch.empty((1000, 256, pool, pool), device=dev, memory_format=torch.channels_last if ocl else torch.contiguous_format)
                pt = L.vp4([f.data_ptr() for f in fms])
                a = host_us(lambda: L.check(L.lib.mrcnn_pyramid_roi_align_forward(pt, Hs, Ws, 1, 256, lay, boxes.data_ptr(), None, 1000, pool,
                                                                                    1024.0 * 1024.0, out.data_ptr(), L.NHWC if ocl else L.NCHW, None, st)))
                b = host_us(lambda: m.pyramid_roi_align(fms, boxes, None, pool, (1024, 1024, 3), out_channels_last=ocl))
                c = host_us(lambda: m.roi_align([boxes.unsqueeze(0)] + fms, pool, [1024, 1024, 3]))
                tag = "%dx%d %s, %s crops" % (pool, pool, name, "channels-last" if ocl else "NCHW")
                print("%-62s %10.1f %10.1f   C ABI" % (tag, a[0], a[1]))
                print("%-62s %10.1f %10.1f   ops.pyramid_roi_align" % ("", b[0], b[1]))
                print("%-62s %10.1f %10.1f   ops.roi_align (crops follow the pyramid)" % ("", c[0], c[1]))
    e = host_us(lambda: torch.empty((1000, 256, 7, 7), device=dev), 5000)
    print("%-62s %10.1f %10.1f" % ("torch.empty((1000,256,7,7))", e[0], e[1]))


if __name__ == "__main__":
    main()
