# Round check on one B200: parity suite, smoke, bench (with CPU baseline), reference arm, ncu launch list
# of the bench command, ncu full-set capture of the hot kernels. TAG names the outputs.
set -x
TAG=${TAG:-r1}
mkdir -p gpurun_out
nvidia-smi -L | head -3
python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/${TAG}_pytest_gpu.log; tail -5 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-600 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; cut -c1-400 gpurun_out/${TAG}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bench.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:'density_kernel|multiple_scattering|single_scattering' -s 4 -c 7 -o gpurun_out/${TAG}_prof python tools/ncu_target.py > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out | tail -12
