set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests_v21.log 2>&1; tail -3 gpurun_out/gpu_tests_v21.log
timeout 900 python bench.py --no-cpu --steps 3 > gpurun_out/v21_bench.log 2>&1
python - gpurun_out/v21_bench.log <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        j = json.loads(l); c = j.get('compress') or {}
        print('%-26s dec %.1f e2e %s | comp %s rt %s e2e %s' % (sys.argv[1][11:], j['value'], (j.get('e2e') or {}).get('value'), c.get('value'), (c.get('roundtrip_decompress') or {}).get('value'), (c.get('e2e') or {}).get('value')))
PY
