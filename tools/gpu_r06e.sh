#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/r06e_bench.json 2> gpurun_out/r06e_bench.err; echo rc=$?
python - <<'PY'
import json
d = json.load(open("gpurun_out/r06e_bench.json"))
r = d["roofline"]
print(d["steps"], d["ms_per_step"], d["value"], r["kernel"], r["frac"], r["step_frac"], r["step_frac_footprint_once"], d["rpn_nms"]["images_per_s"], d["e2e"]["value"], d["clocks"])
PY
