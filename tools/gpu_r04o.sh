#!/bin/bash
mkdir -p gpurun_out
for g in default 32 128; do
  echo "=== $g" >> gpurun_out/r04o.log
  timeout 400 python tools/exp_l2gran.py $g 2>&1 | grep -v "^$" | tail -25 >> gpurun_out/r04o.log
done
cat gpurun_out/r04o.log
