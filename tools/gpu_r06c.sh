#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/check_graph_capture.py > gpurun_out/r06c.log 2>&1; echo "rc=$?" >> gpurun_out/r06c.log
tail -5 gpurun_out/r06c.log
