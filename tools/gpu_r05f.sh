#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "deterministic or pyramid or backward" 2>&1 | tail -4 > gpurun_out/r05f_tests.log
MRCNN_B200_DEBUG=1 timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "deterministic or backward" 2>&1 | tail -3 >> gpurun_out/r05f_tests.log
MRCNN_DETERMINISTIC=1 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -m gpu -k "pyramid or backward or host_train_step or config3" 2>&1 | tail -3 >> gpurun_out/r05f_tests.log
MRCNN_DETERMINISTIC=1 timeout 300 python - > gpurun_out/r05f_time.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
import torch, bench
wl = bench.Workload(torch, torch.device("cuda", 0))
def plans():
    wl.plan(14, wl.ws, torch.cuda.current_stream()); wl.plan(7, wl.ws7, torch.cuda.current_stream())
print("deterministic plans (both heads): %.1f us" % (wl.time_op(plans) * 1e6))
step = bench.capture_step(torch, wl)
print("step with deterministic plans: %.4f ms" % (wl.time_op(step, iters=30) * 1e3))
PY
timeout 300 python - >> gpurun_out/r05f_time.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
import torch, bench
wl = bench.Workload(torch, torch.device("cuda", 0))
def plans():
    wl.plan(14, wl.ws, torch.cuda.current_stream()); wl.plan(7, wl.ws7, torch.cuda.current_stream())
print("default plans (both heads): %.1f us" % (wl.time_op(plans) * 1e6))
PY
cat gpurun_out/r05f_tests.log gpurun_out/r05f_time.log
