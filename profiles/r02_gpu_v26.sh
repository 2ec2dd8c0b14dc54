set -x
# v26: parallel resolve + look-ahead prefetch in the encoder; A/B of the two knobs
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_v26.log 2>&1; tail -3 gpurun_out/gpu_tests_v26.log
for v in p0l0 p1l0 p0l1; do
  LZF_B200_LIB=build/v_$v.so timeout 600 python bench.py --decomp-gib 0.25 --no-e2e --no-cpu --steps 5 --warmup 3 > gpurun_out/v26_$v.log 2>&1; tail -c 1500 gpurun_out/v26_$v.log | grep -o '"compress": {"metric[^}]*' | head -1
done
timeout 600 python bench.py --decomp-gib 0.25 --no-e2e --no-cpu --steps 5 --warmup 3 > gpurun_out/v26_p1l1.log 2>&1; grep -o '"compress": {"metric[^}]*' gpurun_out/v26_p1l1.log | head -1
ncu --set full --clock-control none --import-source on -k regex:encode_blocks -s 3 -c 1 -o gpurun_out/prof_encode_v26 timeout 900 python bench.py --decomp-gib 0.25 --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_enc_v26.log 2>&1
