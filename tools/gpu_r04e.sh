#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_nms.py > gpurun_out/r04e_trace.log 2>&1
cat gpurun_out/r04e_trace.log
