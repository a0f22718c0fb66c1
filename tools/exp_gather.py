"""Times the channels-last gather backward (14x14 and 7x7, planned) at the configs[3] geometry for MRCNN_GATHER_CTAS = persistent CTAs
per SM (0: one CTA per unit), one process per setting.  python tools/exp_gather.py [settings...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import torch
    import bench
    wl = bench.Workload(torch, torch.device("cuda", 0))
    wl.plan(14, wl.ws, torch.cuda.current_stream())
    wl.plan(7, wl.ws7, torch.cuda.current_stream())
    t14 = wl.time_op(lambda: wl.bwd_planned(14, wl.g14, wl.gfm14, wl.ws), iters=30, warm=5)
    t7 = wl.time_op(lambda: wl.bwd_planned(7, wl.g7, wl.gfm7, wl.ws7), iters=30, warm=5)
    print("ctas/sm %-4s bwd14 %.4f ms   bwd7 %.4f ms" % (os.environ.get("MRCNN_GATHER_CTAS", "dflt"), t14 * 1e3, t7 * 1e3), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        for v in (sys.argv[1:] or ["0", "16", "20", "24", "32", "64"]):
            subprocess.call([sys.executable, os.path.abspath(__file__), "one"], env=dict(os.environ, MRCNN_GATHER_CTAS=v))
