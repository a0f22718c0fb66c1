#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -m gpu -k "proposal or rpn_refine or golden_detect or flow" 2>&1 | tail -5 > gpurun_out/r05a_tests.log
MRCNN_B200_DEBUG=1 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -m gpu -k "proposal" 2>&1 | tail -3 >> gpurun_out/r05a_tests.log
timeout 300 python tools/time_proposal.py > gpurun_out/r05a_time.log 2>&1
for pct in 110 150; do echo "pct $pct" >> gpurun_out/r05a_time.log; MRCNN_PROPOSAL_PREFIX_PCT=$pct timeout 300 python tools/time_proposal.py 2>&1 | grep hybrid >> gpurun_out/r05a_time.log; done
cat gpurun_out/r05a_tests.log gpurun_out/r05a_time.log
