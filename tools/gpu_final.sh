#!/bin/bash
# round-end validation: whole GPU suite (release and -DMRCNN_DEBUG builds), smoke, both bench arms
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest.log; tail -3 gpurun_out/final_pytest.log
MRCNN_B200_DEBUG=1 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/final_pytest_debug.log 2>&1; echo "debug rc=$?" >> gpurun_out/final_pytest_debug.log; tail -2 gpurun_out/final_pytest_debug.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
( time python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err ) 2>&1 | grep real; tail -c 300 gpurun_out/final_bench.err
( time python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.load(open("gpurun_out/final_bench.json"))
print("steps", d["steps"], "ms_per_step", d["ms_per_step"], "value", d["value"], "step_frac", d["roofline"]["step_frac"], "top", d["roofline"]["kernel"], d["roofline"]["frac"])
print("e2e", d["e2e"]["value"], d["e2e_fused_backward"]["value"], "clocks", d["clocks"])
a = d["also"]
for k in ("roialign_fwd_7x7", "roialign_fwd_14x14"):
    print(k, round(a[k]["us"], 1), round(a[k]["frac_of_hbm"], 3), "nchw crops", round(a[k]["through_ops_nchw_crops"]["us"], 1), round(a[k]["through_ops_nchw_crops"]["frac_of_hbm"], 3))
print("nchw step", a["nchw_pyramid"]["train_step_nchw_pyramid"])
print("fused step", a["train_step_fused_backward"]["ms_per_step"], a["train_step_fused_backward"]["rois_per_s"])
print("nms", a["nms_standalone"])
print("sharded", {k: d["detection_path_sharded"][k] for k in ("ms_per_64_images", "ms_rank_local_part", "ms_all_gather", "g_invariant")})
print("rpn", d["rpn_nms"]["images_per_s"])
r = json.load(open("gpurun_out/final_bench_ref.json")); print("ref", r["value"], r["steps"])
PY
