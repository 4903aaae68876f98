mkdir -p gpurun_out
for C in 1 3 5; do
  python bench.py --config $C --steps 10 --warmup 3 > gpurun_out/r2_config$C.json 2> gpurun_out/r2_config$C.err; echo "config $C rc=$?"; cut -c1-300 gpurun_out/r2_config$C.json; tail -2 gpurun_out/r2_config$C.err
done
timeout 900 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_config4.json 2> gpurun_out/r2_config4.err; echo "config 4 rc=$?"; cut -c1-300 gpurun_out/r2_config4.json; tail -2 gpurun_out/r2_config4.err
python - <<PY
import json
for c in (1,3,4,5):
    try:
        d=json.load(open(f"gpurun_out/r2_config{c}.json"))
        print(c, d["value"], d["e2e"]["value"], d.get("phases_ms"), d.get("parity",{}) and d["parity"].get("ok"), d.get("cpu_baseline",{}).get("value"))
    except Exception as e: print(c, "no result", e)
PY
