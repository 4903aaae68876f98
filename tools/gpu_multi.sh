# multi-GPU check on the GPUs of one box: 2-rank parity tests (peer + NCCL exchange), then the bench
# at the rank counts given as arguments, with both exchanges
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x ${PYTEST_K:+-k "$PYTEST_K"} 2>&1 | tail -15 | tee gpurun_out/multi_pytest.log
for N in "$@"; do
  for X in ${EXCHANGES:-peer nccl}; do
    echo "== N=$N exchange=$X"
    PAS_EXCHANGE=$X timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_$X.json 2> gpurun_out/bench_n${N}_$X.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_n${N}_$X.json"))
    print({k: d[k] for k in ("value", "n_gpus")}, "e2e", d["e2e"]["value"], {k: v["ms"] for k, v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("no result:", e)
PY
    tail -3 gpurun_out/bench_n${N}_$X.err
  done
done
