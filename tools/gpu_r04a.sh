#!/bin/bash
# fixed-point nms route: parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "nms" 2>&1 | tail -5 > gpurun_out/r04a_tests.log
for n in 500 1000 2000 6000 12000; do
  for r in serial fixpoint; do
    MRCNN_NMS_SWEEP=$r timeout 120 python tools/time_nms.py $n 2>&1 | tail -1 | sed "s/^/route=$r /" >> gpurun_out/r04a_time.log
  done
done
for t in 256 1024; do MRCNN_NMS_SWEEP=fixpoint MRCNN_NMS_THREADS=$t timeout 120 python tools/time_nms.py 6000 2>&1 | tail -1 | sed "s/^/route=fixpoint /" >> gpurun_out/r04a_time.log; done
MRCNN_B200_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "nms and not routes" 2>&1 | tail -3 >> gpurun_out/r04a_tests.log
cat gpurun_out/r04a_tests.log gpurun_out/r04a_time.log
