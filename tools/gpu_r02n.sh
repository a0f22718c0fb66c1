#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_pytest.log
tail -6 gpurun_out/r02n_pytest.log
MRCNN_B200_DEBUG=1 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r02n_pytest_debug.log 2>&1; echo "debug pytest rc=$?" >> gpurun_out/r02n_pytest_debug.log
tail -6 gpurun_out/r02n_pytest_debug.log
python tools/prof_nms.py > gpurun_out/r02n_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02n_nms_launches.csv python tools/prof_nms.py > gpurun_out/r02n_ncu.log 2>&1; tail -2 gpurun_out/r02n_ncu.log
