#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02x_pytest.log
tail -5 gpurun_out/r02x_pytest.log
MRCNN_B200_DEBUG=1 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x -k "pyramid or crop or roi_align or fullsize or reference_model" > gpurun_out/r02x_pytest_debug.log 2>&1; tail -2 gpurun_out/r02x_pytest_debug.log
python tools/time_nchw.py > gpurun_out/r02x_nchw.log 2>&1; tail -50 gpurun_out/r02x_nchw.log
