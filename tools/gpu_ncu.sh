# ncu --set full of the hot kernels of one bench-workload precompute; $1 = kernel regex, $2 = output tag
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-4} -c ${4:-4} -o gpurun_out/$2 python tools/ncu_target.py > gpurun_out/$2.log 2>&1
tail -2 gpurun_out/$2.log
