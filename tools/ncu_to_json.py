"""profiles/ncu_hot_kernels.json from an `ncu --set full` raw CSV export: per hot pass (bench.py phase
key) the DRAM traffic per launch, the pipe utilisations and the EXECUTED instruction census (thread
instructions per SASS opcode -> executed fp32 flops and MUFU ops per launch, L1 data-pipe wavefronts
per SM) that bench.py divides by its live CUDA-event timings.
Usage: ncu -i rep.ncu-rep --page raw --csv --print-metric-instances details > raw.csv
       python tools/ncu_to_json.py raw.csv out.json SOURCE"""
import csv
import json
import re
import sys

csv.field_size_limit(1 << 30)
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
    if "ray_setup" in name:
        return "ray_setup"
    return None


def val(r, name, scale=1.0):
    v = float(r[col[name]].replace(",", "").split(" (")[0])  # instanced metrics: 'aggregate (id: value; ...)'
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


# fp32 flops per thread instruction of the FMA pipe (FFMA2 / FMUL2 / FADD2: Blackwell packed fp32)
FLOPS = {"FFMA": 2, "FFMA2": 4, "FMUL": 1, "FMUL2": 2, "FADD": 1, "FADD2": 2}
OPCODES = ("FFMA", "FFMA2", "FMUL", "FMUL2", "FADD", "FADD2", "MUFU", "LDS", "STS", "LDG", "STG", "LDL", "STL",
           "SHFL", "BAR", "DFMA", "DMUL", "DADD", "UBLKCP", "UTMALDG", "SYNCS")


def opcode_counts(cell):
    """'19934058680 (FADD: 6557616128; FFMA2: ...)' -> {opcode: thread instructions}."""
    m = re.search(r"\((.*)\)", cell)
    if not m:
        return None
    out = {}
    for part in m.group(1).split(";"):
        if ":" in part:
            k, v = part.split(":")
            out[k.strip()] = int(v.strip())
    return out


acc = {}
census = {}
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
    for out, name in (("l1_wavefronts_per_sm", "l1tex__data_pipe_lsu_wavefronts.avg"),
                      ("l1_wavefronts_shared_per_sm", "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg"),
                      ("sm_cycles_elapsed", "sm__cycles_elapsed.avg")):
        full = [h for h in hdr if h == name or h.endswith("." + name)]
        if full:
            e[out] = e.get(out, 0.0) + val(r, full[0])
    if "sass__thread_inst_executed_true_per_opcode" in col:
        ops = opcode_counts(r[col["sass__thread_inst_executed_true_per_opcode"]])
        if ops:
            c = census.setdefault(k, {})
            for op, n in ops.items():
                c[op] = c.get(op, 0) + n
out = {"source": sys.argv[3] if len(sys.argv) > 3 else sys.argv[1], "passes": {}}
for k, e in acc.items():
    n = e.pop("launches")
    kern = e.pop("kernel")
    d = {m: round(v / n, 3) for m, v in e.items()}
    d["traffic_bytes"] = round(d.pop("dram_read_bytes") + d.pop("dram_write_bytes"))
    d["kernel"], d["launches_averaged"] = kern, n
    if k in census:
        ops = {op: census[k].get(op, 0) / n for op in OPCODES if census[k].get(op, 0)}
        d["thread_instructions"] = {op: round(v) for op, v in ops.items()}
        d["fp32_flop_executed"] = round(sum(FLOPS[op] * v for op, v in ops.items() if op in FLOPS))
        d["mufu_executed"] = round(ops.get("MUFU", 0))
    out["passes"][k] = d
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
