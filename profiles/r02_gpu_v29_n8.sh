set -x
# v29 (8 GPUs): copy ceiling of the box with all ranks copying at once (with / without CPU affinity), then the bench at N=8
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1; nproc >> gpurun_out/topo_n8.txt; free -g >> gpurun_out/topo_n8.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
PROBE_AFFINITY=0 timeout 300 $TR 29511 profiles/pcie_probe_multi.py > gpurun_out/v29_pcie_n8.json 2> gpurun_out/v29_pcie_n8.err; tail -c 1500 gpurun_out/v29_pcie_n8.json
PROBE_AFFINITY=1 timeout 300 $TR 29512 profiles/pcie_probe_multi.py > gpurun_out/v29_pcie_n8_aff.json 2> gpurun_out/v29_pcie_n8_aff.err; tail -c 600 gpurun_out/v29_pcie_n8_aff.json
timeout 900 $TR 29513 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/v29_bench_n8.log 2> gpurun_out/v29_bench_n8.err; tail -c 1500 gpurun_out/v29_bench_n8.log; tail -5 gpurun_out/v29_bench_n8.err
