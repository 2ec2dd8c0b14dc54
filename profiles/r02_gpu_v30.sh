set -x
# v30: gpu tests with the new API (compress2 + table, segmented parse, knobs read once), then the full bench line
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_v30.log 2>&1; tail -5 gpurun_out/gpu_tests_v30.log
( time timeout 1500 python bench.py ) > gpurun_out/v30_bench.log 2> gpurun_out/v30_bench.err; tail -c 1800 gpurun_out/v30_bench.log; tail -5 gpurun_out/v30_bench.err
