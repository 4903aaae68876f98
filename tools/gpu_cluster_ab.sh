mkdir -p gpurun_out
for CL in ${CLS:-2 1}; do
  echo "== cluster $CL"
  PAS_DENSITY_CLUSTER=$CL timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cl$CL.json 2> gpurun_out/bench_cl$CL.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_cl$CL.json").read().strip().splitlines()[-1])
print(d["value"], {k[:12]+k[-2:]: v["ms"] for k, v in d["roofline"]["kernels"].items()})
PY
done
