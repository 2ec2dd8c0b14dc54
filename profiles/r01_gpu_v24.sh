set -x
timeout 900 python bench.py > gpurun_out/v24_bench.log 2>&1; tail -c 300 gpurun_out/v24_bench.log
ncu --set full --clock-control none --import-source on -k regex:encode_blocks -s 3 -c 1 -o gpurun_out/prof_encode_v24 timeout 900 python bench.py --decomp-gib 0.25 --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_enc_v24.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"blocks_kernel|xxh32|frame_|stage_dict" -c 80 --csv --log-file gpurun_out/launches_v24.csv timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/launches_v24.log 2>&1
