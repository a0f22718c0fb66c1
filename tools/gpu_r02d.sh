#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x -k "crop or pyramid or roi_align or reference_model or fullsize or binding or adjoint or hypothesis" > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -15 gpurun_out/r02d_pytest.log
python tools/time_nchw.py > gpurun_out/r02d_nchw.log 2>&1; cat gpurun_out/r02d_nchw.log | tail -40
python tools/prof_nchw.py 14 > gpurun_out/r02d_plain.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:"roialign_fwd_nchw|roialign_bwd_nchw|zero_levels" -s 3 -c 3 -o gpurun_out/r02d_nchw14 python tools/prof_nchw.py 14 > gpurun_out/r02d_ncu.log 2>&1; tail -2 gpurun_out/r02d_ncu.log
