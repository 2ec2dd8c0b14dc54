set -x
# v27: tight walk + input stream prefetch (default build), A/B against the round-1 resolve loop, the walk alone, walk + prefetch + look-ahead
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_v27.log 2>&1; tail -3 gpurun_out/gpu_tests_v27.log
for v in base base_pf walk walk_pf_la; do
  LZF_B200_LIB=build/v_$v.so timeout 600 python bench.py --decomp-gib 0.25 --no-e2e --no-cpu --steps 5 --warmup 3 > gpurun_out/v27_$v.log 2>&1; grep -o '"compress": {"metric[^}]*' gpurun_out/v27_$v.log | head -1 | cut -c100-260
done
timeout 600 python bench.py --decomp-gib 0.25 --no-e2e --no-cpu --steps 5 --warmup 3 > gpurun_out/v27_walk_pf.log 2>&1; grep -o '"compress": {"metric[^}]*' gpurun_out/v27_walk_pf.log | head -1 | cut -c100-260
ncu --set full --clock-control none --import-source on -k regex:encode_blocks -s 3 -c 1 -o gpurun_out/prof_encode_v27 timeout 900 python bench.py --decomp-gib 0.25 --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_enc_v27.log 2>&1
