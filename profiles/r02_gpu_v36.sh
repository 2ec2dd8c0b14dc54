set -x
# v36: A/B on one box: decoder gather limit 12 vs 20 (config 2 and full-size text decode), encoder walk with / without carried literals; e2e chunk sizes
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_v36.log 2>&1; tail -3 gpurun_out/gpu_tests_v36.log
for v in default v_dec_g20 v_enc_nocarry default; do
  if [ $v = default ]; then unset LZF_B200_LIB; else export LZF_B200_LIB=build/$v.so; fi
  timeout 600 python bench.py --no-e2e --no-cpu --no-extra --steps 5 --warmup 3 > gpurun_out/v36_$v.log 2>&1
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/v36_$v.log') if l.startswith('{')][-1])
print('AB $v', 'decode', round(d['value'],1), 'text decode', round(d['compress']['roundtrip_decompress']['value'],1), 'compress', round(d['compress']['value'],2), 'frames', round(d['compress']['frames_device']['compress_GiB_per_s'],2), round(d['compress']['frames_device']['decompress_GiB_per_s'],1))
PY
done
unset LZF_B200_LIB
for cb in 134217728 268435456; do
  LZF_B200_CHUNK_BYTES=$cb timeout 600 python bench.py --no-compress --no-cpu --no-extra --steps 8 --warmup 3 > gpurun_out/v36_dec_cb$cb.log 2>&1
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/v36_dec_cb$cb.log') if l.startswith('{')][-1])
print('CB $cb', 'e2e', d['e2e']['value'], 'ceiling', d['e2e']['ceiling_gbs'], d['e2e']['frac_of_ceiling'])
PY
done
