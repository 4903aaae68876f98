set -x
mkdir -p gpurun_out
nvidia-smi -L | head -3
python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r1_pytest_gpu.log; tail -5 gpurun_out/r1_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; cat gpurun_out/r1_bench.json; tail -3 gpurun_out/r1_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches.csv python tools/ncu_target.py > gpurun_out/r1_ncu_target.log 2>&1
tail -2 gpurun_out/r1_ncu_target.log
ncu --set full --clock-control none --import-source on -k regex:'density_kernel|multiple_scattering|single_scattering' -s 4 -c 7 -o gpurun_out/r1_prof python tools/ncu_target.py > gpurun_out/r1_ncu_full.log 2>&1
tail -2 gpurun_out/r1_ncu_full.log
ls -la gpurun_out
