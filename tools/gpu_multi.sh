# multi-GPU check on N GPUs of one box: 2-rank parity test + bench at N ranks
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
cat gpurun_out/bench_n$N.json | cut -c1-1500; tail -5 gpurun_out/bench_n$N.err
