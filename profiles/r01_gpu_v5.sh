set -x
nvidia-smi > gpurun_out/box.txt; nproc >> gpurun_out/box.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests.log 2>&1
tail -5 gpurun_out/gpu_tests.log
( time python bench.py ) > gpurun_out/bench_v5.log 2>&1
tail -2 gpurun_out/bench_v5.log
ncu --set full --clock-control none --import-source on -k regex:decode_blocks -s 3 -c 1 -o gpurun_out/prof_decode_v5 python bench.py --no-compress --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_dec.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode_blocks -s 3 -c 1 -o gpurun_out/prof_encode_v5 python bench.py --decomp-gib 0.25 --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_enc.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v5.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/launches_v5.log 2>&1
ls -la gpurun_out
