# bench at the given rank counts on one box (peer exchange unless PAS_EXCHANGE is set)
mkdir -p gpurun_out
for N in "$@"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}_${TAG:-x}.json 2> gpurun_out/bench_n${N}_${TAG:-x}.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n${N}_${TAG:-x}.json").read().strip().splitlines()[-1])
print("N=$N", d["value"], "e2e", d["e2e"]["value"], {k[:12]+k[-2:]: v["ms"] for k, v in d["roofline"]["kernels"].items()})
PY
done
