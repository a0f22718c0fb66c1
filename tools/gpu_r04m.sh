#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_binding.py -q -x -m gpu -k "nms" 2>&1 | tail -3 > gpurun_out/r04m_tests.log
MRCNN_B200_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "nms and not routes" 2>&1 | tail -3 >> gpurun_out/r04m_tests.log
MRCNN_NMS_PUB=0 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "nms and not routes" 2>&1 | tail -3 >> gpurun_out/r04m_tests.log
for n in 500 1000 2000 6000 8000; do timeout 120 python tools/time_nms.py $n 2>&1 | tail -1 | sed "s/^/pub /">> gpurun_out/r04m_time.log; MRCNN_NMS_PUB=0 timeout 120 python tools/time_nms.py $n 2>&1 | tail -1 | sed "s/^/barrier /" >> gpurun_out/r04m_time.log; done
for n in 500 1000; do MRCNN_NMS_SWEEP=serial timeout 120 python tools/time_nms.py $n 2>&1 | tail -1 | sed "s/^/serial /">> gpurun_out/r04m_time.log; done
for t in 256 1024; do MRCNN_NMS_THREADS=$t timeout 120 python tools/time_nms.py 6000 2>&1 | tail -1 >> gpurun_out/r04m_time.log; done
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_active.max,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r04m_nms_launches.csv python tools/prof_nms.py 6000 > gpurun_out/r04m_ncu.log 2>&1
cat gpurun_out/r04m_tests.log gpurun_out/r04m_time.log
