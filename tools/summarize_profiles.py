"""Turns the ncu artefacts under gpurun_out/ (tools/run_profiles.sh) into the committed summaries under profiles/.

  <tag>_launches.csv   (ncu --metrics gpu__time_duration.sum over `bench.py`)   -> per-kernel share of the step
  <tag>_full.ncu-rep   (ncu --set full, 16-bit path, tools/profile_once.py --batch 256)
  <tag>_fp32.ncu-rep   (ncu --set full, fp32 CUDA-core path, batch 8)
  <tag>_front.ncu-rep  (ncu --set full, front-end kernels + the 300 / 600 variants)
plus a SASS opcode listing of the shipped library (tcgen05 / TMA evidence).   Usage: summarize_profiles.py [tag]
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ncu kernel name -> the engine's name for it (bench.py: kernels_ms_per_step)
ENGINE_NAME = {"prep_u8_kernel": "prep_u8", "conv_tc_kernel<1,16,31,1,2,0,8,0>": "conv0_tc",
               "conv_tc_kernel<1,32,41,1,1,0,32,0>": "conv1_tc", "block2_fused_kernel<0>": "block2_tc",
               "conv_tc_kernel<4,32,41,1,0,0,32,0>": "conv2_tc", "conv_tc_kernel<4,32,41,1,0,0,32,1>": "conv3_tc",
               "conv_tc_kernel<4,64,42,1,0,0,64,0>": "conv4_tc", "conv_tc_kernel<8,64,42,1,0,0,64,1>": "conv5_tc",
               "conv_tc_kernel<8,64,0,2,0,0,64,0>": "conv6_tc", "conv_tc_kernel<16,16,42,2,0,0,16,0>": "conv7_tc",
               "tail_fused_kernel": "tail_fused"}


def short(name):
    m = re.search(r"(conv_tc_kernel<[^>]*>|block2_fused_kernel<[^>]*>|[a-z0-9_]+_kernel(?:<[^>]*>)?)", name)
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
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400`; times are cold-cache and\n"
                "serialised (compare SHARES with bench.py's `kernels_ms_per_step`, not absolutes).\n\n")
        f.write("| kernel | engine name | launches | total us | share |\n|---|---|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %s | %d | %.1f | %.1f %% |\n" % (k, ENGINE_NAME.get(k, ""), n, t / 1e3, 100 * t / total))
    print("wrote launches summary,", len(agg), "kernels")

# ---- full captures -----------------------------------------------------------------------------
WANT = collections.OrderedDict([
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC (SM)"),
    ("sm__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active"),
    ("launch__registers_per_thread", "registers"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
])
CAPTURES = [("full", "16-bit tensor-core path, `tools/profile_once.py --batch 256` (one launch = one half-batch of 128 images)"),
            ("fp32", "fp32 CUDA-core path, `tools/profile_once.py --batch 8 --precision fp32`"),
            ("fp32tc", "fp32-class tensor-core path (split-fp16 layers), `tools/profile_once.py --batch 256 --precision fp32tc`"),
            ("front", "front-end kernels and the 300 / 600 variants, `tools/profile_front.py`")]
traffic = {}
for suffix, what in CAPTURES:
    rep = os.path.join(ROOT, "gpurun_out", "%s_%s.ncu-rep" % (tag, suffix))
    csv_path = os.path.join(ROOT, "gpurun_out", "%s_%s.raw.csv" % (tag, suffix))  # exported on the box (run_profiles.sh)
    if os.path.exists(csv_path):
        raw = open(csv_path).read()
    elif os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        continue
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    seen, lines = set(), []

    def to_bytes(v, u):
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    for r in rows[2:]:
        k = short(r[h.index("Kernel Name")])
        if k in seen:
            continue
        seen.add(k)
        vals = []
        for m in WANT:
            if m in h:
                i = h.index(m)
                vals.append((WANT[m], r[i].replace(",", ""), units[i]))
        lines.append((k, vals))
        rd = to_bytes(r[h.index("dram__bytes_read.sum")].replace(",", ""), units[h.index("dram__bytes_read.sum")])
        wr = to_bytes(r[h.index("dram__bytes_write.sum")].replace(",", ""), units[h.index("dram__bytes_write.sum")])
        if suffix == "full":
            traffic[ENGINE_NAME.get(k, k)] = {"dram_bytes_per_launch_b128": rd + wr, "ncu_kernel": k}
    with open(os.path.join(out_dir, "%s_ncu_%s_summary.md" % (tag, suffix)), "w") as f:
        f.write("# %s — `ncu --set full --clock-control none --import-source on`: %s\n\n" % (tag, what))
        f.write("First occurrence of every kernel; the .ncu-rep itself stays in gpurun_out/ (tens of MB).\n\n")
        for k, vals in lines:
            f.write("## `%s`%s\n\n" % (k, " (= %s)" % ENGINE_NAME[k] if k in ENGINE_NAME else ""))
            for name, v, u in vals:
                f.write("* %s: %s %s\n" % (name, v, u))
            f.write("\n")
    print("wrote %s summary, %d kernels" % (suffix, len(lines)))
if traffic:
    json.dump(traffic, open(os.path.join(out_dir, "ncu_traffic.json"), "w"), indent=1)

# ---- SASS opcode listing of the shipped library --------------------------------------------------
lib = os.path.join(ROOT, "roomnet_b200", "libroomnet.so")
if os.path.exists(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    per_fn, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per_fn.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            op = m.group(1)
            if op in ("UTCHMMA", "UTCBAR", "UTMALDG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "FHFMA",
                      "FHADD", "FADD2", "FFMA2", "FMUL2", "HFMA2", "HADD2", "ELECT", "ACQBULK", "NANOSLEEP"):
                cur[op + ("" if op not in ("UTMALDG", "LDTM", "STTM") else (m.group(2) or ""))] += 1
    total = collections.Counter()
    for c in per_fn.values():
        total.update(c)
    with open(os.path.join(out_dir, tag + "_sass_opcodes.md"), "w") as f:
        f.write("# %s — Blackwell-specific SASS in roomnet_b200/libroomnet.so (`cuobjdump -sass`)\n\n" % tag)
        f.write("UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor load, UBLKCP = TMA bulk copy, UTMAPF = TMA\n"
                "prefetch, SYNCS = mbarrier ops, FHFMA/FHADD = mixed-precision fma/add, FADD2/FFMA2/FMUL2 = packed fp32.\n\n")
        f.write("Whole library: " + ", ".join("%s %d" % kv for kv in sorted(total.items(), key=lambda kv: -kv[1])) + "\n\n")
        f.write("| kernel | opcode counts |\n|---|---|\n")
        for fn, c in per_fn.items():
            if not c or not any(k.startswith(("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP")) for k in c):
                continue
            name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
            f.write("| `%s` | %s |\n" % (short(name), ", ".join("%s %d" % kv for kv in sorted(c.items(), key=lambda kv: -kv[1]))))
    print("wrote SASS opcode listing,", len(per_fn), "functions")
