"""Experiment: the step's four independent kernels (two-head forward, mask targets, two gather backwards) on one stream against
two / three concurrent lanes inside the step's CUDA graph.  python tools/exp_overlap.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

wl = bench.Workload(torch, torch.device("cuda", 0))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def lanes2():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur)
    with torch.cuda.stream(s1):
        wl.plan(14, wl.ws, s1); wl.plan(7, wl.ws7, s1)
        wl.bwd_planned(14, wl.g14, wl.gfm14, wl.ws)
        wl.bwd_planned(7, wl.g7, wl.gfm7, wl.ws7)
    wl.fwd_pair(); wl.mask_targets()
    cur.wait_stream(s1)

def lanes2b():   # backward 14 beside the forward, backward 7 after both
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur)
    with torch.cuda.stream(s1):
        wl.plan(14, wl.ws, s1); wl.plan(7, wl.ws7, s1)
        wl.bwd_planned(14, wl.g14, wl.gfm14, wl.ws)
    wl.fwd_pair(); wl.mask_targets()
    cur.wait_stream(s1)
    wl.bwd_planned(7, wl.g7, wl.gfm7, wl.ws7)

def lanes3():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        wl.plan(14, wl.ws, s1)
        wl.bwd_planned(14, wl.g14, wl.gfm14, wl.ws)
    with torch.cuda.stream(s2):
        wl.plan(7, wl.ws7, s2)
        wl.bwd_planned(7, wl.g7, wl.gfm7, wl.ws7)
        wl.mask_targets()
    wl.fwd_pair()
    cur.wait_stream(s1); cur.wait_stream(s2)

def graph_of(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    cap = torch.cuda.Stream(); cap.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=cap):
        fn()
    torch.cuda.current_stream().wait_stream(cap)
    return g.replay

import numpy as np
ref = None
for name, fn in (("one stream (the bench step)", wl.step), ("two lanes: fwd+targets | bwd14, bwd7", lanes2), ("two lanes, bwd7 after the join", lanes2b), ("three lanes", lanes3)):
    r = graph_of(fn)
    for g in wl.gfm14 + wl.gfm7: g.zero_()
    wl.out14.zero_()
    r(); torch.cuda.synchronize()
    sig = (float(wl.gfm14[0].double().sum()), float(wl.gfm7[1].double().sum()), float(wl.out14.double().sum()), float(wl.mt.double().sum()))
    if ref is None: ref = sig
    ts = [wl.time_op(r, iters=50) * 1e3 for _ in range(3)]
    print("%-40s %s ms  same results: %s" % (name, " ".join("%.4f" % t for t in ts), sig == ref))
