set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests_v19.log 2>&1; tail -3 gpurun_out/gpu_tests_v19.log
LZF_B200_TRACE=1 timeout 600 python bench.py --decomp-gib 0.25 --no-cpu --steps 3 > gpurun_out/v19_trace.log 2>&1
grep "lzf trace" gpurun_out/v19_trace.log | tail -2
python - gpurun_out/v19_trace.log <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        j = json.loads(l); c = j.get('compress') or {}
        print('%-26s comp %s e2e %s' % (sys.argv[1][11:], c.get('value'), (c.get('e2e') or {}).get('value')))
PY
