#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02z_pytest.log
tail -4 gpurun_out/r02z_pytest.log
bash tools/collect_profiles.sh r02 > gpurun_out/r02z_collect.log 2>&1; tail -3 gpurun_out/r02z_collect.log
python tools/prof_nchw.py 14 > gpurun_out/r02z_plain.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:"roialign_fwd_nchw|roialign_bwd_nchw" -s 2 -c 2 -o gpurun_out/r02_nchw14 python tools/prof_nchw.py 14 > gpurun_out/r02z_ncu.log 2>&1
python bench.py --steps 50 --warmup 5 > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; echo "bench rc=$?"; tail -c 500 gpurun_out/r02z_bench.err
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/r02z_bench_ref.json 2> gpurun_out/r02z_bench_ref.err; echo "ref rc=$?"
ls gpurun_out | grep "^r02_" | head -30
