"""Per-kernel counts of the SASS mnemonics that show which hardware paths libmrcnn_b200.so uses (cuobjdump -sass; no GPU
needed).  usage: python tools/sass_evidence.py  ->  profiles/r02_sass_evidence.txt"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATS = {"UBLKCP (cp.async.bulk: 1-D TMA bulk copy)": r"UBLKCP", "SYNCS (mbarrier)": r"SYNCS", "UCGABAR (cluster barrier)": r"UCGABAR",
        "cluster window (distributed shared memory: MAPA / .CLUSTER)": r"MAPA|\.CLUSTER",
        "LDGSTS (cp.async)": r"LDGSTS", "FFMA2 / FADD2 / FMUL2 (packed fp32x2)": r"FFMA2|FADD2|FMUL2", "SHFL": r"SHFL",
        "MATCH.ANY": r"MATCH\.ANY", "VOTE": r"VOTE", "RED (reduction without return)": r"\bRED\.", "ATOMS (shared-memory atomic)": r"ATOMS",
        "MUFU.RCP": r"MUFU\.RCP", "128-bit / streaming global stores": r"STG\.E\.(EF\.)?128|STG\.E\.EF", "128-bit global loads": r"LDG\.E\.[A-Z.]*128",
        "LDS.128": r"LDS\.128", "PRMT (byte permute)": r"PRMT", "fp64 arithmetic": r"\bD(ADD|MUL|FMA)\b"}


def main():
    lib = os.path.join(ROOT, "maskrcnn_b200", "libmrcnn_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    parts = re.split(r"\n\s*Function : ", txt)[1:]
    rows = []
    for part in parts:
        name = part.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", "-p", name], capture_output=True, text=True).stdout.strip() or name
        rows.append((dem.replace("mrcnn::", ""), {k: len(re.findall(v, part)) for k, v in PATS.items()}))
    out = os.path.join(ROOT, "profiles", "r02_sass_evidence.txt")
    with open(out, "w") as f:
        f.write("# cuobjdump -sass maskrcnn_b200/libmrcnn_b200.so (sm_100a): per kernel, how many instructions of each kind\n"
                "# (only the mnemonics that show which hardware path a kernel uses).  Regenerate: python tools/sass_evidence.py\n\n")
        for dem, c in sorted(rows):
            nz = ["%s: %d" % (k, c[k]) for k in PATS if c[k]]
            f.write("%s\n    %s\n" % (dem[:160], "; ".join(nz) if nz else "-"))
    print("wrote", out, len(rows), "kernels")


if __name__ == "__main__":
    main()
