TAG=r2t REFARM=1 bash tools/gpu_round2.sh
timeout 600 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_config4.json 2> gpurun_out/r2t_config4.err; echo "config 4 rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2t_config4.json")); print(4, d["value"], d["e2e"]["value"], d.get("phases_ms"))
PY
