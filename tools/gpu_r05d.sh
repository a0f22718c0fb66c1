#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -m gpu -k "proposal" 2>&1 | tail -3 > gpurun_out/r05d_tests.log
timeout 300 python tools/time_proposal.py > gpurun_out/r05d_time.log 2>&1
cat gpurun_out/r05d_tests.log gpurun_out/r05d_time.log
