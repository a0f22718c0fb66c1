"""Per-source-line instruction / stall-sample totals of one kernel from an .ncu-rep captured with
--import-source on (code built with -lineinfo).  usage: ncu_lines.py <rep> <kernel regex> [min_share]"""
import csv, io, subprocess, sys

rep, kern = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname, hdr, ci = "", None, None
agg = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        ci = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        ie = int(r[ci["Instructions Executed"]]); smp = int(r[ci["# Samples"]])
    except ValueError:
        continue
    key = (fname, int(r[0]))
    a = agg.setdefault(key, [0, 0, r[1].strip()[:100]])
    a[0] += ie; a[1] += smp
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
for (f, ln), (ie, smp, src) in sorted(agg.items()):
    if ie >= tot_i * min_share or smp >= tot_s * min_share:
        print("%-14s %5d  inst %5.1f%%  samples %5.1f%%  %s" % (f, ln, 100.0 * ie / tot_i, 100.0 * smp / tot_s, src))
