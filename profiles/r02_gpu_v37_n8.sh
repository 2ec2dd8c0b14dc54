set -x
# v37 (8 GPUs): frame gather / scatter over NCCL on its own (default channels, then more point-to-point channels), then the bench at N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
PROBE_GIB=2 timeout 300 $TR 29521 profiles/gather_probe.py 2> gpurun_out/v37_gather_a.err | tail -1 | tee gpurun_out/v37_gather_default.json
NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 PROBE_GIB=2 timeout 300 $TR 29522 profiles/gather_probe.py 2> gpurun_out/v37_gather_b.err | tail -1 | tee gpurun_out/v37_gather_p2p32.json
( time timeout 1200 $TR 29513 bench.py --gpus 8 --steps 5 --warmup 3 ) > gpurun_out/v37_bench_n8.log 2> gpurun_out/v37_bench_n8.err; tail -c 2500 gpurun_out/v37_bench_n8.log; tail -6 gpurun_out/v37_bench_n8.err
