#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "cuda_graph" 2>&1 | tail -4 > gpurun_out/r06d.log
MRCNN_B200_DEBUG=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "cuda_graph" 2>&1 | tail -2 >> gpurun_out/r06d.log
cat gpurun_out/r06d.log
