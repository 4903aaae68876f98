"""Static opcode histogram of one kernel: python tools/sass_ops.py <obj/.so> <mangled-name-substring> [top]"""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
flt = sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
name, c = None, collections.Counter()
def flush():
    if name and flt in name:
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"pas::\(anonymous namespace\)::|void ", "", dem).split("(")[0]
        print(dem[:110], "total", sum(c.values()))
        print("   " + " ".join(f"{k}={v}" for k, v in c.most_common(top)))
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush(); name, c = m.group(1), collections.Counter(); continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m: c[m.group(1)] += 1
flush()
