#!/bin/bash
mkdir -p gpurun_out
( time python bench.py --steps 50 --warmup 5 > gpurun_out/r06b_bench.json 2> gpurun_out/r06b_bench.err ) 2>&1 | grep real; tail -c 300 gpurun_out/r06b_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r06b_bench.json"))
r = d["roofline"]
print(d["ms_per_step"], r["frac"], r["step_frac"], r["step_frac_footprint_once"], {k: r["kernels"]["roialign_fwd_nhwc_pair_kernel<7+14,nhwc>"].get(k) for k in ("ms", "frac", "footprint_once_MB", "frac_footprint_once", "ncu_dram_MB")})
print(d["rpn_nms"]["images_per_s"], d["clocks"])
PY
