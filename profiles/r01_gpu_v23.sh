set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests_v23.log 2>&1; tail -3 gpurun_out/gpu_tests_v23.log
timeout 600 python bench.py --decomp-gib 0.25 --no-e2e --no-cpu --steps 3 > gpurun_out/v23_comp.log 2>&1
python - gpurun_out/v23_comp.log <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        j = json.loads(l); c = j.get('compress') or {}
        print('comp %s rt %s' % (c.get('value'), (c.get('roundtrip_decompress') or {}).get('value')))
PY
