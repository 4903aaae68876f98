# Builds variants/libpas_<NAME>.so: the listed kernel sources recompiled with extra nvcc flags, linked with
# the regular objects of the other sources (which `make -C csrc` must have built). A/B runs pick a variant
# with PAS_B200_LIB (model.py). Usage: tools/build_variant.sh NAME "-DPAS_X=1 ..." file.cu [file.cu ...]
set -e
NAME=$1; FLAGS=$2; shift 2
CSRC=precomputed_atmospheric_scattering_b200/csrc
mkdir -p variants $CSRC/build/var_$NAME
OBJS=""
for f in $(ls $CSRC/*.cu); do
  b=$(basename $f .cu)
  if echo " $* " | grep -q " $b.cu "; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --fmad=false -lineinfo -Xcompiler -fPIC -Xptxas -v -I/usr/include --expt-relaxed-constexpr $FLAGS -c $f -o $CSRC/build/var_$NAME/$b.o 2> $CSRC/build/var_$NAME/$b.ptxas.log &
    OBJS="$OBJS $CSRC/build/var_$NAME/$b.o"
  else
    OBJS="$OBJS $CSRC/build/$b.o"
  fi
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libpas_$NAME.so $OBJS -ldl
ls -la variants/libpas_$NAME.so
