# A/B of kernel variants on one B200: microbenchmark, parity suite with the new kernels, per-phase timings
mkdir -p gpurun_out
nvidia-smi -L | head -2
if [ -n "$UBENCH" ]; then
nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o /tmp/ffma2 tools/ubench/ffma2.cu && /tmp/ffma2 2>&1 | tee gpurun_out/ubench_ffma2.log
fi
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/variants_pytest.log
for v in ${VARIANTS:-"PAS_X=0"}; do
  echo "== $v"; env $v timeout 300 python tools/ncu_target.py; env $v timeout 300 python tools/ncu_target.py --rgb
done 2>&1 | tee gpurun_out/variants_timing.log
if [ -n "$NCU_K" ]; then
ncu --set full --clock-control none --import-source on -k regex:"$NCU_K" -s ${NCU_S:-3} -c ${NCU_C:-3} -o gpurun_out/${NCU_TAG:-prof} python tools/ncu_target.py > gpurun_out/${NCU_TAG:-prof}.log 2>&1
tail -2 gpurun_out/${NCU_TAG:-prof}.log
fi
