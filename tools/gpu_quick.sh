# Quick A/B on one B200: selected parity tests + a short bench without the CPU baseline. TAG names the outputs;
# PYTEST_K selects tests (default: the Earth golden / digest / KAT / chained small runs).
TAG=${TAG:-quick}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "${PYTEST_K:-golden or digest or kat or chained or bench_product or channel_counts or teacher}" 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log; tail -6 gpurun_out/${TAG}_pytest.log
python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "parity ok", d["parity"]["ok"], d["parity"]["max_floor"])
    print({k: v for k, v in d["phases_ms"].items()})
except Exception as e:
    print("no bench result:", e)
PY
