set -x
nvidia-smi > gpurun_out/box_v16.txt; nproc >> gpurun_out/box_v16.txt; free -g >> gpurun_out/box_v16.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests_v16.log 2>&1; tail -3 gpurun_out/gpu_tests_v16.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v16.log 2>&1; tail -1 gpurun_out/smoke_v16.log
timeout 900 python bench.py > gpurun_out/v16_bench.log 2>&1; tail -c 400 gpurun_out/v16_bench.log
timeout 600 python bench.py --impl reference > gpurun_out/v16_reference.log 2>&1; tail -c 300 gpurun_out/v16_reference.log
ncu --set full --clock-control none --import-source on -k regex:decode_blocks -s 3 -c 1 -o gpurun_out/prof_decode_v16 timeout 600 python bench.py --no-compress --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_dec_v16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode_blocks -s 3 -c 1 -o gpurun_out/prof_encode_v16 timeout 900 python bench.py --decomp-gib 0.25 --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_enc_v16.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"blocks_kernel|xxh32|frame_|stage_dict" -c 80 --csv --log-file gpurun_out/launches_v16.csv timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/launches_v16.log 2>&1
timeout 900 python bench.py --extra --no-e2e --decomp-gib 0.25 --comp-gib 1 --steps 3 > gpurun_out/v16_extra.log 2>&1
