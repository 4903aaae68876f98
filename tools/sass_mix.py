"""Instruction mix per kernel from `cuobjdump -sass` of a cubin / .so: counts of the FP32/FP64/MUFU/
memory opcodes that matter for this path. Usage: python tools/sass_mix.py <file> [name-filter]"""
import collections
import re
import subprocess
import sys

OPS = ["FFMA", "FMUL", "FADD", "FMNMX", "MUFU", "DFMA", "DMUL", "DADD", "LDS", "STS", "LDG", "STG", "LDL", "STL", "SHFL", "BAR"]
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
flt = sys.argv[2] if len(sys.argv) > 2 else ""
name, counts, total = None, collections.Counter(), 0
def flush():
    if name and flt in name:
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        d = re.sub(r"pas::\(anonymous namespace\)::|void ", "", d).split("(")[0]
        print(f"{d[:48]:48s} n={total:6d} " + " ".join(f"{o}={counts[o]}" for o in OPS if counts[o]))
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, counts, total = m.group(1), collections.Counter(), 0
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        total += 1
        op = m.group(1)
        for o in OPS:
            if op == o or op.startswith(o + "."):
                counts[o] += 1
flush()
