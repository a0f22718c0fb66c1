#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "deterministic" 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r05g_tests.log
cat gpurun_out/r05g_tests.log
