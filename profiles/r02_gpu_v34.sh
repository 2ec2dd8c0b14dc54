set -x
# v34: tests, full bench line, decode capture with source
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_v34.log 2>&1; tail -3 gpurun_out/gpu_tests_v34.log
( time timeout 1500 python bench.py ) > gpurun_out/v34_bench.log 2> gpurun_out/v34_bench.err; tail -c 1500 gpurun_out/v34_bench.log; tail -5 gpurun_out/v34_bench.err
ncu --set full --clock-control none --import-source on -k regex:decode_blocks -s 3 -c 1 -o gpurun_out/prof_decode_v34 timeout 900 python bench.py --comp-gib 0.25 --no-e2e --no-cpu --no-extra --steps 1 --warmup 3 > gpurun_out/ncu_dec_v34.log 2>&1
