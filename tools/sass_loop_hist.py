"""Static opcode histogram of the epilogue loop of block2_fused_kernel<false> (from the first LDTM to the last STG):
a quick CPU-side proxy for instructions per epilogue iteration.  Usage: python tools/sass_loop_hist.py [object]"""
import re
import subprocess
import sys
from collections import Counter

obj = sys.argv[1] if len(sys.argv) > 1 else "/root/repo/build/kernels_block2.o"
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
ops, on = [], False
for line in txt.splitlines():
    if "Function" in line:
        on = "block2_fused_kernelILb0E" in line
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if on and m:
        ops.append(m.group(2).strip())
ld = [i for i, o in enumerate(ops) if "LDTM" in o]
st = [i for i, o in enumerate(ops) if "STG" in o]
a, b = ld[0] - 30, st[-1] + 8
c = Counter()
for o in ops[a:b]:
    o = re.sub(r"^@!?U?P\d+\s+", "", o)
    c[o.split()[0] if o.startswith(("IMAD", "LDL", "STL")) else o.split()[0].split(".")[0]] += 1
print("loop region: %d instructions" % (b - a))
print(sorted(c.items(), key=lambda x: -x[1])[:45])
if "--dump" in sys.argv:
    for i in range(a, b):
        print(i, ops[i])
