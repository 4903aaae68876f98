# Copies the outputs of `TAG=$1 NCU=1 SANITIZE=1 REFARM=1 bash tools/gpu_round2.sh` (merged back into gpurun_out/)
# into profiles/ and regenerates the ncu census / summary / JSON and the SASS / ptxas summaries (no GPU needed).
set -e
TAG=${1:-r2f}
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv --print-metric-instances details > /tmp/${TAG}_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/${TAG}_raw.csv profiles/r2_ncu_full_census.csv > profiles/r2_ncu_full_summary.txt 2>&1
python tools/ncu_to_json.py /tmp/${TAG}_raw.csv profiles/ncu_hot_kernels.json "profiles/r2_ncu_full_census.csv (ncu --set full --clock-control none --print-metric-instances details, tools/ncu_target.py = one precompute of the bench workload, B200; round-2 kernels)" > /dev/null
cp gpurun_out/${TAG}_launches.csv profiles/r2_launches.csv
cp gpurun_out/${TAG}_bench.json profiles/r2_bench.json
cp gpurun_out/${TAG}_bench_reference.json profiles/r2_bench_reference.json
cp gpurun_out/${TAG}_pytest_gpu.log profiles/r2_pytest_gpu.log
cp gpurun_out/${TAG}_smoke.log profiles/r2_smoke.log
for t in memcheck racecheck synccheck initcheck; do cp gpurun_out/${TAG}_sanitizer_$t.log profiles/r2_sanitizer_$t.log; done
OBJ=precomputed_atmospheric_scattering_b200/csrc/build/kernel_raymarch.o
{ head -1 profiles/r2_sass_ops_raymarch.txt
  for k in "multiple_scattering_rows_kernelILi15" "single_scattering_kernelILi15ELi256ELi2ELi256ELb1" "ray_setup_kernelILi15"; do python tools/sass_ops.py $OBJ $k 24; done
  echo "# TMA / bulk-copy opcodes (UBLKCP / UTMALDG / LDGSTS) in kernel_raymarch.o:"
  cuobjdump -sass $OBJ | grep -c "UBLKCP\|UTMALDG\|LDGSTS" || true; } > /tmp/so.txt
cp /tmp/so.txt profiles/r2_sass_ops_raymarch.txt
{ echo "# ptxas -v summary of the last build (tools/ptxas_summary.py): registers, spill stores/loads in bytes, static shared memory"; python tools/ptxas_summary.py | grep -v "^#"; } > /tmp/ptx2.txt
cp /tmp/ptx2.txt profiles/r2_ptxas.txt
python - <<PY
import json
d = json.load(open("profiles/r2_bench.json"))
print(d["value"], d["e2e"]["value"], d["gpu_launches"], d["parity"]["ok"])
print({k: (v["ms"], v.get("frac_fp32_peak"), v.get("l1_wavefront_frac")) for k, v in d["roofline"]["kernels"].items()})
PY
