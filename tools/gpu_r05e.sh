#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/exp_overlap.py > gpurun_out/r05e.log 2>&1
tail -8 gpurun_out/r05e.log
