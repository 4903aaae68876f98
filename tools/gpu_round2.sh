# Round-2 check on one B200: parity suite, smoke, bench (with CPU baseline + parity key), reference arm,
# sanitizers on the bench kernel instantiations. TAG names the outputs. Optional: NCU=1 adds the ncu
# launch list of the bench command and the full-set capture of the hot kernels.
set -x
TAG=${TAG:-r2}
mkdir -p gpurun_out
nvidia-smi -L | head -3
python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/${TAG}_pytest_gpu.log; tail -8 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "parity", d["parity"])
    print({k: (v["ms"], v.get("frac_fp32_peak"), v.get("l1_wavefront_frac")) for k, v in d["roofline"]["kernels"].items()})
    print("roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "peak", "frac", "canonical_frac")}, d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("no bench result:", e)
PY
if [ -n "$REFARM" ]; then
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; cut -c1-400 gpurun_out/${TAG}_bench_reference.json
fi
if [ -n "$SANITIZE" ]; then
for TOOL in memcheck racecheck synccheck initcheck; do
  timeout 600 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_target.py > gpurun_out/${TAG}_sanitizer_$TOOL.log 2>&1
  echo "== $TOOL rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok" gpurun_out/${TAG}_sanitizer_$TOOL.log | tail -3
done
fi
if [ -n "$NCU" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bench.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:'density_kernel|multiple_scattering|single_scattering|ray_setup' -s 5 -c 8 -o gpurun_out/${TAG}_prof python tools/ncu_target.py > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
fi
ls -la gpurun_out | tail -12
