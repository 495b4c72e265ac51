"""Turns the ncu artefacts under gpurun_out/ into the committed summaries under profiles/.

  r01_launches.csv  (ncu --metrics gpu__time_duration.sum over `bench.py`)  -> per-kernel share of the step
  r01_full.ncu-rep  (ncu --set full over tools/profile_once.py --batch 256) -> per-kernel roofline inputs
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)


def short(name):
    m = re.search(r"(conv_tc_kernel<[^>]*>|[a-z0-9_]+_kernel)", name)
    s = m.group(1) if m else name
    return s.replace("(int)", "").replace("(bool)", "").replace(" ", "")


# ---- launch list -------------------------------------------------------------------------------
lpath = os.path.join(ROOT, "gpurun_out", tag + "_launches.csv")
if os.path.exists(lpath):
    rows = [r for r in csv.reader(open(lpath)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    iK, iV, iM = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr or len(r) != len(hdr) or r[iM] != "gpu__time_duration.sum":
            continue
        k = short(r[iK])
        t = float(r[iV].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(v[1] for v in agg.values())
    with open(os.path.join(out_dir, tag + "_launches_summary.md"), "w") as f:
        f.write("# %s — ncu launch list of `python bench.py --steps 3 --warmup 3` (batch 256, fp16 path)\n\n" % tag)
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 300`; times are cold-cache and\n"
                "serialised (compare SHARES with bench.py's `kernels_ms_per_step`, not absolutes).\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f %% |\n" % (k, n, t / 1e3, 100 * t / total))
    print("wrote launches summary,", len(agg), "kernels")

# ---- full capture ------------------------------------------------------------------------------
rep = os.path.join(ROOT, "gpurun_out", tag + "_full.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    want = collections.OrderedDict([
        ("gpu__time_duration.sum", "us"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__inst_executed.avg.per_cycle_active", "IPC (SM)"),
        ("dram__bytes_read.sum", "dram read"),
        ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"),
        ("smsp__inst_executed.sum", "warp instr"),
    ])
    units = rows[1]
    seen, traffic = set(), {}
    lines = []
    for r in rows[2:]:
        k = short(r[h.index("Kernel Name")])
        if k in seen:
            continue
        seen.add(k)
        vals = []
        for m in want:
            i = h.index(m)
            v = r[i].replace(",", "")
            u = units[i]
            vals.append((want[m], v, u))
        lines.append((k, vals))
        def to_bytes(v, u):
            mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            return float(v) * mul
        rd = to_bytes(r[h.index("dram__bytes_read.sum")].replace(",", ""), units[h.index("dram__bytes_read.sum")])
        wr = to_bytes(r[h.index("dram__bytes_write.sum")].replace(",", ""), units[h.index("dram__bytes_write.sum")])
        traffic[k] = rd + wr
    with open(os.path.join(out_dir, tag + "_ncu_full_summary.md"), "w") as f:
        f.write("# %s — `ncu --set full --clock-control none --import-source on` over tools/profile_once.py --batch 256\n\n" % tag)
        f.write("One launch = one half-batch of 128 images (the engine splits 256 into two streams). First occurrence of\n"
                "every kernel; the report itself stays in gpurun_out/ (45 MB).\n\n")
        for k, vals in lines:
            f.write("## `%s`\n\n" % k)
            for name, v, u in vals:
                f.write("* %s: %s %s\n" % (name, v, u))
            f.write("\n")
    # map to the engine's kernel names used by bench.py
    layer_of = {"conv_tc_kernel<1,16,31,1,2,0,8,0>": "conv0_tc", "conv_tc_kernel<1,32,41,1,1,0,32,0>": "conv1_tc",
                "conv_tc_kernel<4,32,41,1,0,0,32,0>": "conv2_tc", "conv_tc_kernel<4,32,41,1,0,0,32,1>": "conv3_tc",
                "conv_tc_kernel<4,64,42,1,0,0,64,0>": "conv4_tc", "conv_tc_kernel<8,64,42,1,0,0,64,1>": "conv5_tc",
                "conv_tc_kernel<8,64,0,2,0,0,64,0>": "conv6_tc", "conv_tc_kernel<16,16,42,1,0,0,16,0>": "conv7_tc"}
    tj = {}
    for k, b in traffic.items():
        name = layer_of.get(k, k)
        tj[name] = {"dram_bytes_per_launch_b128": b, "ncu_kernel": k}
    json.dump(tj, open(os.path.join(out_dir, "ncu_traffic.json"), "w"), indent=1)
    print("wrote full summary,", len(lines), "kernels")
