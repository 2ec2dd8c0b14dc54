set -x
# v35: results written by the kernels into pinned host memory (no small D2H copies behind the bulk ones), decoder gather <= 12
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_v35.log 2>&1; tail -3 gpurun_out/gpu_tests_v35.log
for cb in 0 268435456 1073741824; do
  if [ $cb = 0 ]; then unset LZF_B200_CHUNK_BYTES; else export LZF_B200_CHUNK_BYTES=$cb; fi
  timeout 600 python bench.py --no-compress --no-cpu --no-extra --steps 8 --warmup 3 > gpurun_out/v35_dec_cb$cb.log 2>&1
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/v35_dec_cb$cb.log') if l.startswith('{')][-1])
print('chunk $cb', 'decode', d['value'], 'e2e', d['e2e']['value'], 'ceiling', d['e2e']['ceiling_gbs'], d['e2e']['frac_of_ceiling'])
PY
done
unset LZF_B200_CHUNK_BYTES
( time timeout 1500 python bench.py ) > gpurun_out/v35_bench.log 2> gpurun_out/v35_bench.err; tail -c 800 gpurun_out/v35_bench.log; tail -4 gpurun_out/v35_bench.err
