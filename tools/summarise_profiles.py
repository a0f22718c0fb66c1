"""Turns the captures tools/collect_profiles.sh left in gpurun_out/ into the committed text summaries under
profiles/ (launch list + per-kernel ncu --set full counters) and profiles/<round>_traffic.json, which bench.py reads
for roofline.traffic.  usage: python tools/summarise_profiles.py r01"""
import csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# 1. launch list
src = os.path.join(G, "%s_launches.csv" % R)
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
agg, order = {}, []
for r in rows[1:]:
    k = r[ci["Kernel Name"]].split("(")[0][:80]
    if k not in agg:
        agg[k] = [0, 0.0]
        order.append(k)
    agg[k][0] += 1
    agg[k][1] += float(r[ci["Metric Value"]])
mine = sum(v[1] for k, v in agg.items() if "mrcnn::" in k)
with open(os.path.join(P, "%s_launches_summary.txt" % R), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 3 --warmup 3 --no-extras\n")
    f.write("# (6 steps in total; per-launch times are cold-cache and serialised - compare SHARES with bench.py's roofline.kernels)\n")
    f.write("%-82s %6s %12s %10s %8s\n" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k in order:
        share = agg[k][1] / mine if "mrcnn::" in k else float("nan")
        f.write("%-82s %6d %12.1f %10.1f %8.3f\n" % (k, agg[k][0], agg[k][1] / 1e3, agg[k][1] / 1e3 / agg[k][0], share))
shutil.copy(src, os.path.join(P, "%s_launches.csv" % R))
nxt = os.path.join(G, "%s_next_launches.csv" % R)       # launch list of tools/prof_next.py (SURVEY 8f kernels)
if os.path.exists(nxt):
    shutil.copy(nxt, os.path.join(P, "%s_next_launches.csv" % R))

# 2. per-kernel full captures
traffic = {}
for name in sorted(os.listdir(G)):
    if not (name.startswith(R + "_") and name.endswith(".ncu-rep")):
        continue
    rep = os.path.join(G, name)
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(P, name.replace(".ncu-rep", "_ncu_summary.txt")), "w").write(txt.replace(G + "/", "gpurun_out/"))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    col = {h: i for i, h in enumerate(rr[0])}
    units = rr[1]

    def to_bytes(v, u):
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    for r in rr[2:]:
        kn = r[col["Kernel Name"]].split("(")[0]
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        us = float(r[col["gpu__time_duration.sum"]]) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}[units[col["gpu__time_duration.sum"]]]
        key = name.replace(".ncu-rep", "").replace(R + "_", "") + ":" + kn
        if key not in traffic:   # first launch of each kernel in each capture
            traffic[key] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes": rd + wr, "ncu_time_us": us}
json.dump(traffic, open(os.path.join(P, "%s_traffic.json" % R), "w"), indent=1, sort_keys=True)
print(open(os.path.join(P, "%s_launches_summary.txt" % R)).read())
print(json.dumps(traffic, indent=1, sort_keys=True))
