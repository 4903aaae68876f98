# Multi-GPU check on the GPUs of one box: optional 2-rank parity tests, then the bench at the rank counts
# given as arguments (peer exchange unless EXCHANGES says otherwise); prints the per-phase timings.
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ -n "$PYTEST" ]; then
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x ${PYTEST_K:+-k "$PYTEST_K"} 2>&1 | tail -15 | tee gpurun_out/${TAG:-multi}_pytest.log
fi
for N in "$@"; do
  for X in ${EXCHANGES:-symm}; do
    echo "== N=$N exchange=$X"
    PAS_EXCHANGE=$X timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-20} --warmup 5 > gpurun_out/${TAG:-multi}_bench_n${N}_$X.json 2> gpurun_out/${TAG:-multi}_bench_n${N}_$X.err
    echo "rc=$?"
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG:-multi}_bench_n${N}_$X.json"))
    print({k: d[k] for k in ("value", "n_gpus")}, "e2e", d["e2e"]["value"], "parity", d["parity"]["ok"], d["parity"]["ranks_ok"], d["parity"]["max_floor"])
    print(d["phases_ms"], "sum", round(sum(d["phases_ms"].values()), 4))
except Exception as e:
    print("no result:", e)
PY
    tail -3 gpurun_out/${TAG:-multi}_bench_n${N}_$X.err
  done
done
