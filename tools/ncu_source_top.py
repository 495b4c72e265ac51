"""Top instructions by stall samples of one kernel in an ncu report's source page (CSV export).
Usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_source_top.py src.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r][0]
hdr = rows[hi]
data, seen = [], set()
for r in rows[hi + 1:]:
    if len(r) == len(hdr) and r[0] not in seen:
        seen.add(r[0])
        data.append(r)
I = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[I[k]])
    except ValueError:
        return 0.0


tot = sum(int(r[I["# Samples"]]) for r in data)
ex = sum(int(r[I["Instructions Executed"]]) for r in data)
print("instructions", len(data), "samples", tot, "warp-instructions executed", ex)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(f(r, s) for r in data) for s in stalls}
print({k[6:]: int(v) for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > 0})
print("shared wavefronts", int(sum(f(r, "L1 Wavefronts Shared") for r in data)), "ideal",
      int(sum(f(r, "L1 Wavefronts Shared Ideal") for r in data)))
for r in sorted(data, key=lambda r: -int(r[I["# Samples"]]))[:n_top]:
    top = sorted(((f(r, s), s) for s in stalls), reverse=True)[:2]
    print("%6s %9s  %-74s %s" % (r[I["# Samples"]], r[I["Instructions Executed"]], r[I["Source"]][:74],
                                [(s[6:], int(v)) for v, s in top]))
