# render-row check: GPU render tests + 1080p timings
mkdir -p gpurun_out
python -m pytest tests/test_gpu_render.py -x -q -s 2>&1 | tail -40 > gpurun_out/render_tests.log; tail -40 gpurun_out/render_tests.log
python tools/render_bench.py 2>&1 | tail -12 | tee gpurun_out/render_bench.log
