#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
tail -8 gpurun_out/r02i_pytest.log
python tools/time_host_overhead.py > gpurun_out/r02i_host.log 2>&1; grep -E "channels-last pyramid|ops\." gpurun_out/r02i_host.log | head -16
python bench.py --steps 20 --warmup 5 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc=$?"; tail -c 800 gpurun_out/r02i_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02i_bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d.get("e2e", {}).get("value"), d["config"].get("launch"))
print("eager", d.get("eager_step"))
print("roofline", {k: d["roofline"][k] for k in ("kernel", "frac", "step_frac")})
for k, v in d["roofline"]["kernels"].items(): print("  ", k, round(v["ms"], 4), round(v["frac"], 3))
a = d.get("also", {})
for k in ("roialign_fwd_7x7", "roialign_fwd_14x14"):
    print(k, a[k]["us"], a[k]["frac_of_hbm"], {n: (round(v["us"],1), round(v["frac_of_hbm"],3)) for n, v in a[k]["direct_abi"].items()})
print("nchw", json.dumps(a.get("nchw_pyramid"))[:1200])
print("nms", a.get("nms_standalone"))
print("predict", json.dumps(a.get("predict_flow"))[:900])
PY
