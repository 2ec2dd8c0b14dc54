set -x
# v33 (8 GPUs): the bench at N=8 — e2e against the copy ceiling measured in the same run, config 4 with the full exchange
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
( time timeout 1200 $TR 29513 bench.py --gpus 8 --steps 5 --warmup 3 ) > gpurun_out/v33_bench_n8.log 2> gpurun_out/v33_bench_n8.err; tail -c 3000 gpurun_out/v33_bench_n8.log; tail -8 gpurun_out/v33_bench_n8.err
