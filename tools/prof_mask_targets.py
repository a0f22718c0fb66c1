"""Runs the 28x28 mask-target crop of the bench step a few times (for ncu) and prints its unique-pixel and unique-sector bytes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench

wl = bench.Workload(torch, torch.device("cuda", 0))
for _ in range(3):
    wl.mask_targets()
torch.cuda.synchronize()
t = wl.time_op(wl.mask_targets, iters=50)
print("mask targets: %d crops, %.2f us" % (wl.mt.shape[0], t * 1e6))
print("bytes", bench.mask_target_bytes(wl))
