set -x
B="timeout 300 python bench.py --no-compress --no-cpu --no-e2e --steps 3"
$B > gpurun_out/v20_base.log 2>&1
for v in dec_w4_c9 dec_w4_c10; do
  LZF_B200_LIB=build/$v.so $B > gpurun_out/v20_$v.log 2>&1
  NB=4096 LZF_B200_LIB=build/$v.so timeout 300 python profiles/text_decode_probe.py > gpurun_out/v20_text_$v.log 2>&1
done
( time timeout 900 python -m pytest tests -m gpu -x -q -k "sliced" ) > gpurun_out/gpu_tests_v20.log 2>&1; tail -3 gpurun_out/gpu_tests_v20.log
for f in gpurun_out/v20_*.log; do python - "$f" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        j = json.loads(l); print('%-36s dec %.1f' % (sys.argv[1][11:], j['value']))
    elif l.startswith('text decode'): print('%-36s'%sys.argv[1][11:], l.strip())
PY
done
