#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_mask_targets.py > gpurun_out/r04n_plain.log 2>&1 || { tail -5 gpurun_out/r04n_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:crop_plane_fwd -s 2 -c 1 -o gpurun_out/r04n_mt -f python tools/prof_mask_targets.py > gpurun_out/r04n_ncu.log 2>&1
tail -3 gpurun_out/r04n_plain.log; tail -2 gpurun_out/r04n_ncu.log
