"""profiles/ncu_hot_kernels.json from an `ncu --set full` raw CSV export: per hot pass (bench.py phase
key) the DRAM traffic per launch and the pipe utilisations bench.py quotes next to its live timings.
Usage: ncu -i rep.ncu-rep --page raw --csv > raw.csv; python tools/ncu_to_json.py raw.csv out.json SOURCE"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
units = rows[1]


def key_of(name):
    if "density_kernel" in name:
        m = re.search(r"density_kernel\w*<\d+, (\d)", name)
        return "scattering_density_2" if m and m.group(1) == "1" else "scattering_density_n"
    if "multiple_scattering" in name:
        return "multiple_scattering"
    if "single_scattering" in name:
        return "single_scattering"
    return None


def val(r, name, scale=1.0):
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]]
    if u == "Mbyte":
        v *= 1e6
    elif u == "Kbyte":
        v *= 1e3
    elif u == "Gbyte":
        v *= 1e9
    elif u == "ms":
        v *= 1.0
    elif u == "us":
        v *= 1e-3
    return v * scale


acc = {}
for r in rows[2:]:
    k = key_of(r[col["Kernel Name"]])
    if k is None:
        continue
    e = acc.setdefault(k, {"launches": 0, "kernel": re.sub(r"\(.*", "", r[col["Kernel Name"]].replace("void unnamed>::", ""))})
    e["launches"] += 1
    for out, name in (("ms_under_ncu", "gpu__time_duration.sum"),
                      ("dram_read_bytes", "dram__bytes_read.sum"), ("dram_write_bytes", "dram__bytes_write.sum"),
                      ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                      ("fma_pipe_cycles_active_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                      ("xu_pipe_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                      ("l1_data_pipe_wavefronts_pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                      ("l2_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                      ("dram_throughput_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                      ("warp_instructions", "smsp__inst_executed.sum"),
                      ("registers_per_thread", "launch__registers_per_thread")):
        e[out] = e.get(out, 0.0) + val(r, name)
out = {"source": sys.argv[3] if len(sys.argv) > 3 else sys.argv[1], "passes": {}}
for k, e in acc.items():
    n = e.pop("launches")
    kern = e.pop("kernel")
    d = {m: round(v / n, 3) for m, v in e.items()}
    d["traffic_bytes"] = round(d.pop("dram_read_bytes") + d.pop("dram_write_bytes"))
    d["kernel"], d["launches_averaged"] = kern, n
    out["passes"][k] = d
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
