"""Times the channels-last 14x14 forward at the configs[3] geometry for the kernel variant MRCNN_FWD14 selects (one process per
variant: the choice is fixed at first use).  python tools/exp_fwd14.py  -> runs itself once per variant."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import torch
    import bench
    wl = bench.Workload(torch, torch.device("cuda", 0))
    t = wl.time_op(lambda: wl.fwd(14, wl.out14), iters=30, warm=5)
    t7 = wl.time_op(lambda: wl.fwd(7, wl.out7), iters=30, warm=5)
    ref = os.environ.get("EXP_REF")
    chk = ""
    if ref:
        if os.path.exists(ref):
            chk = " identical_to_col=%s" % bool(torch.equal(torch.load(ref), wl.out14.cpu()))
        else:
            torch.save(wl.out14.cpu(), ref)
    print("%-5s fwd14 %.4f ms   fwd7 %.4f ms%s" % (os.environ.get("MRCNN_FWD14", "dflt"), t * 1e3, t7 * 1e3, chk), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        ref = "/tmp/exp_fwd14_ref.pt"
        if os.path.exists(ref):
            os.remove(ref)
        for v in (sys.argv[1:] or ["col", "row", "tma"]):
            subprocess.call([sys.executable, os.path.abspath(__file__), "one"], env=dict(os.environ, MRCNN_FWD14=v, EXP_REF=ref))
