#!/bin/bash
mkdir -p gpurun_out
MRCNN_NMS_BACKOFF=65536 timeout 300 python tools/trace_nms.py 2>&1 | tail -17 | cut -c1-420 > gpurun_out/r04h_trace.log
cat gpurun_out/r04h_trace.log
