set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/s3_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/s3_pytest.log
python bench.py > $O/s3_bench.json 2> $O/s3_bench.err; echo "bench rc=$?"; tail -2 $O/s3_bench.err
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"full_masks|rpn_pack" -c 4 -o $O/s3_paste python tools/prof_paste.py > $O/s3_paste.log 2>&1; tail -1 $O/s3_paste.log
