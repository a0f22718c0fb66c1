#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r03c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03c_pytest.log
tail -4 gpurun_out/r03c_pytest.log
MRCNN_B200_DEBUG=1 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -k "pair or fused or fullsize" > gpurun_out/r03c_pytest_debug.log 2>&1; tail -1 gpurun_out/r03c_pytest_debug.log
bash tools/collect_profiles.sh r02 > gpurun_out/r03c_collect.log 2>&1; tail -2 gpurun_out/r03c_collect.log
( time python bench.py --steps 50 --warmup 5 > gpurun_out/r03c_bench.json 2> gpurun_out/r03c_bench.err ) 2>&1 | grep real; tail -c 400 gpurun_out/r03c_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r03c_bench.json'))
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'step_frac', d['roofline']['step_frac'], 'top', d['roofline']['kernel'], d['roofline']['frac'])
for k,v in d['roofline']['kernels'].items(): print('  ', k[:60], round(v['ms'],4), round(v['frac'],3))
print(d['roofline']['step_with_one_forward_per_head'], d['eager_step'])
"
