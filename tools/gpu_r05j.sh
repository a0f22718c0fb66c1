#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -m gpu -k "pyramid or backward or host_train_step or config3 or deterministic or crop" 2>&1 | tail -3 > gpurun_out/r05j_tests.log
MRCNN_B200_DEBUG=1 timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -m gpu -k "pyramid or backward or config3 or deterministic" 2>&1 | tail -3 >> gpurun_out/r05j_tests.log
timeout 300 python - > gpurun_out/r05j_time.log 2>&1 <<'PY'
import os, sys; sys.path.insert(0, '.')
import torch, bench
wl = bench.Workload(torch, torch.device("cuda", 0))
wl.plan(14, wl.ws, torch.cuda.current_stream()); wl.plan(7, wl.ws7, torch.cuda.current_stream())
t14 = min(wl.time_op(lambda: wl.bwd_planned(14, wl.g14, wl.gfm14, wl.ws), iters=30) for _ in range(3))
t7 = min(wl.time_op(lambda: wl.bwd_planned(7, wl.g7, wl.gfm7, wl.ws7), iters=30) for _ in range(3))
step = bench.capture_step(torch, wl)
ts = min(wl.time_op(step, iters=30) for _ in range(3))
print("gather14 %.4f ms  gather7 %.4f ms  step %.4f ms" % (t14 * 1e3, t7 * 1e3, ts * 1e3))
PY
cat gpurun_out/r05j_tests.log gpurun_out/r05j_time.log
