set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/box_r02.txt; nproc >> gpurun_out/box_r02.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_v25.log 2>&1; tail -3 gpurun_out/gpu_tests_v25.log
ncu --set full --clock-control none --import-source on -k regex:encode_blocks -s 3 -c 1 -o gpurun_out/prof_encode_v25 timeout 900 python bench.py --decomp-gib 0.25 --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_enc_v25.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_blocks -s 3 -c 1 -o gpurun_out/prof_decode_v25 timeout 900 python bench.py --comp-gib 0.25 --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_dec_v25.log 2>&1
ls -la gpurun_out/*.ncu-rep
