#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nms_mask_lower -s 2 -c 1 -o gpurun_out/r04l_mask -f python tools/prof_nms.py 6000 > gpurun_out/r04l_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nms_fixpoint_pub -s 2 -c 1 -o gpurun_out/r04l_fix -f python tools/prof_nms.py 6000 >> gpurun_out/r04l_ncu.log 2>&1
tail -3 gpurun_out/r04l_ncu.log
