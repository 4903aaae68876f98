# A/B of library variants (variants/libpas_*.so, tools/build_variant.sh) on one B200: per-phase timings of
# the bench workload for each, the in-tree library first.
mkdir -p gpurun_out
echo "== in-tree"; python tools/ncu_target.py
for so in variants/libpas_*.so; do
  echo "== $so"; PAS_B200_LIB=$PWD/$so python tools/ncu_target.py
done 2>&1 | tee gpurun_out/${TAG:-ab}_timing.log
