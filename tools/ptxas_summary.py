"""Summarises the ptxas -v logs of the last build (registers, spills, shared memory per kernel)."""
import os
import re
import subprocess

build = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                     "precomputed_atmospheric_scattering_b200", "csrc", "build")
for f in sorted(os.listdir(build)):
    if not f.endswith(".ptxas.log"):
        continue
    txt = open(os.path.join(build, f)).read()
    for b in re.split(r"ptxas info\s+: Compiling entry function '", txt)[1:]:
        name = b.split("'")[0]
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = dem.replace("pas::(anonymous namespace)::", "").replace("void ", "")
        dem = dem.split("(PasGeometry")[0].split("(float const")[0]
        regs = re.search(r"Used (\d+) registers", b)
        spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b)
        smem = re.search(r"(\d+) bytes smem", b)
        print(f"{dem[:70]:70s} regs={regs.group(1) if regs else '?':>3} "
              f"spill={spill.group(1) if spill else '?'}/{spill.group(2) if spill else '?'} "
              f"smem={smem.group(1) if smem else 0}")
