set -x
B="timeout 300 python bench.py --no-compress --no-cpu --no-e2e --steps 3"
$B > gpurun_out/v13_new.log 2>&1
LZF_B200_LIB=build/head_v12.so $B > gpurun_out/v13_old.log 2>&1
NB=4096 timeout 300 python profiles/text_decode_probe.py > gpurun_out/v13_text_new.log 2>&1
NB=4096 LZF_B200_LIB=build/head_v12.so timeout 300 python profiles/text_decode_probe.py > gpurun_out/v13_text_old.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests_v13.log 2>&1; tail -3 gpurun_out/gpu_tests_v13.log
for f in gpurun_out/v13_*.log; do python - "$f" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        j = json.loads(l); print('%-26s dec %.1f' % (sys.argv[1][11:], j['value']))
    elif l.startswith('text decode'): print(sys.argv[1][11:], l.strip())
PY
done
