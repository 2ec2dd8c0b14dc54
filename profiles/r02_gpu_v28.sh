set -x
# v28: the new bench.py contract end to end on one GPU (default flags), then the reference arm
( time timeout 1500 python bench.py ) > gpurun_out/v28_bench.log 2> gpurun_out/v28_bench.err; tail -c 600 gpurun_out/v28_bench.log; tail -5 gpurun_out/v28_bench.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 2 ) > gpurun_out/v28_ref.log 2> gpurun_out/v28_ref.err; tail -c 400 gpurun_out/v28_ref.log; tail -4 gpurun_out/v28_ref.err
