"""Per-instruction stall breakdown of one kernel in an .ncu-rep (source page): prints the hottest SASS lines with their
dominant stall reasons, plus totals per stall reason.  Usage: python tools/ncu_stalls.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]
ix = {k: i for i, k in enumerate(hdr)}
data = rows[2:]
stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
tot = Counter()
for r in data:
    for k in stall_cols:
        tot[k] += int(r[ix[k]] or 0)
allS = sum(tot.values())
print("total samples", allS)
for k, v in tot.most_common():
    print("  %-24s %6d %5.1f%%" % (k, v, 100.0 * v / allS))
print("instructions executed (warp):", sum(int(r[ix["Instructions Executed"]]) for r in data))
ranked = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:top]
for i in sorted(ranked):
    r = data[i]
    st = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stall_cols), reverse=True)[:3]
    print("%5d %6s  %-60s %s" % (i, r[ix["# Samples"]], r[ix["Source"]].strip()[:60], " ".join("%s:%d" % (k, v) for v, k in st if v)))
