#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "nms_edge" 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r06a_tests.log
MRCNN_B200_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "nms_edge" 2>&1 | tail -3 >> gpurun_out/r06a_tests.log
MRCNN_NMS_PUB=0 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "nms_edge" 2>&1 | tail -3 >> gpurun_out/r06a_tests.log
cat gpurun_out/r06a_tests.log
