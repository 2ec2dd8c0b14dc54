set -x
# v31: gpu tests (streaming reader/writer added), the full bench line (copy ceiling, config 5 segmented, single file)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_v31.log 2>&1; tail -5 gpurun_out/gpu_tests_v31.log
( time timeout 1500 python bench.py ) > gpurun_out/v31_bench.log 2> gpurun_out/v31_bench.err; tail -c 2500 gpurun_out/v31_bench.log; tail -5 gpurun_out/v31_bench.err
