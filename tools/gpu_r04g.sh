#!/bin/bash
mkdir -p gpurun_out
for m in 0 65536 131072 196608 262144; do MRCNN_NMS_BACKOFF=$m timeout 120 python tools/time_nms.py 6000 2>&1 | tail -1 | sed "s/^/mode=$m /" >> gpurun_out/r04g_time.log; done
cat gpurun_out/r04g_time.log
