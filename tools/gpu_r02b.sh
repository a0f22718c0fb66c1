#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -k "crop or pyramid or roi_align or reference_model or fullsize or binding or adjoint or hypothesis or smoke" > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -40 gpurun_out/r02b_pytest.log
python tools/time_nchw.py > gpurun_out/r02b_nchw.log 2>&1; cat gpurun_out/r02b_nchw.log | tail -40
