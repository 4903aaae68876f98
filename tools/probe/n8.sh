TAG=r2l EXCHANGES=symm bash tools/gpu_multi2.sh 8
echo "=== density blocks of 256 threads"
PAS_DENSITY_THREADS=256 TAG=r2l_d256 EXCHANGES=symm bash tools/gpu_multi2.sh 8
echo "=== layers dealt round-robin"
PAS_LAYER_DEAL=rr TAG=r2l_rr EXCHANGES=symm bash tools/gpu_multi2.sh 8 4
echo "=== skew probe"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 tools/probe/skew_probe.py 2>&1 | grep -E "back-to-back|host barrier"
