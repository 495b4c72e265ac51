"""Static opcode histogram of the innermost loops of block2_fused_kernel<false> that contain TMEM loads (the epilogue
loops): a quick CPU-side proxy for instructions per epilogue iteration.
Usage: python tools/sass_loop_hist.py [object] [--dump N]   (N = index of the loop to list)"""
import re
import subprocess
import sys
from collections import Counter

args = [a for a in sys.argv[1:] if not a.startswith("--")]
obj = args[0] if args else "/root/repo/build/kernels_block2.o"
func = "block2_fused_kernelILb0E"
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
ops, on = [], False
for line in txt.splitlines():
    if "Function" in line:
        on = func in line
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if on and m:
        ops.append((int(m.group(1), 16), m.group(2).strip()))
addr2idx = {a: i for i, (a, _) in enumerate(ops)}
loops = []
for i, (a, o) in enumerate(ops):
    m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", o)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr2idx:
            j = addr2idx[tgt]
            n_ld = sum("LDTM" in x for _, x in ops[j:i + 1])
            if n_ld:
                loops.append((j, i, n_ld))
# keep innermost loops only
inner = [l for l in loops if not any(o is not l and o[0] >= l[0] and o[1] <= l[1] for o in loops)]
def opname(o):
    o = re.sub(r"^@!?U?P\d+\s+", "", o)
    return o.split()[0] if o.startswith(("IMAD", "LDL", "STL")) else o.split()[0].split(".")[0]
for k, (j, i, n_ld) in enumerate(inner):
    c = Counter(opname(o) for _, o in ops[j:i + 1])
    print("loop %d: %d instructions, %d LDTM (= %d layer steps) -> %.0f per layer step" % (k, i - j + 1, n_ld, n_ld // 2, (i - j + 1) / (n_ld / 2)))
    print("  ", sorted(c.items(), key=lambda x: -x[1])[:40])
if "--dump" in sys.argv:
    k = int(sys.argv[sys.argv.index("--dump") + 1])
    j, i, _ = inner[k]
    for q in range(j, i + 1):
        print("%5d %s" % (q, ops[q][1]))
