set -x
B="timeout 300 python bench.py --no-compress --no-cpu --no-e2e --steps 3"
$B > gpurun_out/v15_fused.log 2>&1
LZF_B200_LIB=build/v15_nofuse.so $B > gpurun_out/v15_nofuse.log 2>&1
NB=4096 timeout 300 python profiles/text_decode_probe.py > gpurun_out/v15_text_fused.log 2>&1
NB=4096 LZF_B200_LIB=build/v15_nofuse.so timeout 300 python profiles/text_decode_probe.py > gpurun_out/v15_text_nofuse.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests_v15.log 2>&1; tail -3 gpurun_out/gpu_tests_v15.log
timeout 600 python bench.py --no-cpu --steps 3 > gpurun_out/v15_bench.log 2>&1
for f in gpurun_out/v15_*.log; do python - "$f" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        j = json.loads(l); c = j.get('compress') or {}
        print('%-26s dec %.1f e2e %s | comp %s rt %s e2e %s' % (sys.argv[1][11:], j['value'], (j.get('e2e') or {}).get('value'), c.get('value'), (c.get('roundtrip_decompress') or {}).get('value'), (c.get('e2e') or {}).get('value')))
    elif l.startswith('text decode'): print(sys.argv[1][11:], l.strip())
PY
done
