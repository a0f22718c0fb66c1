#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_nms.py 6000 > gpurun_out/r04b_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_active.max,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r04b_nms_launches.csv python tools/prof_nms.py 6000 > gpurun_out/r04b_ncu.log 2>&1
tail -2 gpurun_out/r04b_ncu.log
