#!/bin/bash
# round 2, first GPU call: whole GPU suite (new full-size / reference-model / binding tests included), bench both arms, host overhead
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -30 gpurun_out/r02a_pytest.log
python tools/time_host_overhead.py > gpurun_out/r02a_host.log 2>&1; tail -30 gpurun_out/r02a_host.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02a_bench.err
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02a_bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d.get("e2e", {}).get("value"))
print("roofline", {k: d["roofline"][k] for k in ("kernel", "frac", "step_frac")})
for k, v in d["roofline"]["kernels"].items(): print("  ", k, round(v["ms"], 4), round(v["frac"], 3))
a = d.get("also", {})
for k in ("roialign_fwd_7x7", "roialign_fwd_14x14", "nchw_pyramid", "nms_standalone", "predict_flow", "cpu_baselines_other_configs"):
    print(k, json.dumps(a.get(k))[:1500])
print("sharded", json.dumps(d.get("detection_path_sharded"))[:800])
r = json.load(open("gpurun_out/r02a_bench_ref.json"))
print("ref", r["value"], r["steps"], r["cpu_baseline"]["sample"][:200])
PY
