"""Experiment: cudaLimitMaxL2FetchGranularity (32 / 64 / 128 bytes) against the sparse-access kernels (mask-target crop, NCHW
pyramid RoIAlign) and the headline step.  usage: exp_l2gran.py <bytes|default>"""
import os, runpy, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.cuda.init()
torch.zeros(1, device="cuda")
from cuda import cudart
lim = cudart.cudaLimit.cudaLimitMaxL2FetchGranularity
print("granularity before:", cudart.cudaDeviceGetLimit(lim))
if sys.argv[1] != "default":
    print("set ->", cudart.cudaDeviceSetLimit(lim, int(sys.argv[1])))
print("granularity now:", cudart.cudaDeviceGetLimit(lim))
import bench
wl = bench.Workload(torch, torch.device("cuda", 0))
t = wl.time_op(wl.mask_targets, iters=50)
print("mask targets %.2f us" % (t * 1e6))
step = bench.capture_step(torch, wl) or wl.step
t = wl.time_op(step, iters=20)
print("headline step %.4f ms" % (t * 1e3))
del wl, step
torch.cuda.empty_cache()
sys.argv = ["time_nchw.py"]
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "time_nchw.py"), run_name="__main__")
