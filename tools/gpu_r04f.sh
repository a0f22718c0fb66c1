#!/bin/bash
mkdir -p gpurun_out
for b in 0 32 64 128 256 512; do MRCNN_NMS_BACKOFF=$b timeout 120 python tools/time_nms.py 6000 2>&1 | tail -1 | sed "s/^/backoff=$b /" >> gpurun_out/r04f_time.log; done
MRCNN_NMS_BACKOFF=128 timeout 300 python tools/trace_nms.py 2>&1 | tail -14 | cut -c1-400 > gpurun_out/r04f_trace.log
cat gpurun_out/r04f_time.log gpurun_out/r04f_trace.log
