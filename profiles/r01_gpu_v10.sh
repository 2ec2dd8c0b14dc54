set -x
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests_v10.log 2>&1; tail -3 gpurun_out/gpu_tests_v10.log
timeout 900 python bench.py > gpurun_out/v10_bench.log 2>&1; tail -c 600 gpurun_out/v10_bench.log
for sl in 0 65536; do LZF_B200_FEED_SLICE=$sl timeout 600 python bench.py --decomp-gib 0.25 --no-cpu --steps 2 > gpurun_out/v10_feed$sl.log 2>&1; done
timeout 900 python bench.py --extra --no-e2e --decomp-gib 0.25 --comp-gib 1 --steps 3 > gpurun_out/v10_extra.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"blocks_kernel|xxh32|frame_|stage_dict" -c 150 --csv --log-file gpurun_out/launches_v10.csv timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --comp-gib 4 > gpurun_out/launches_v10.log 2>&1
for f in gpurun_out/v10_*.log; do python - "$f" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        j = json.loads(l); c = j.get('compress') or {}
        print('%-26s dec %.1f e2e %s | comp %s rt %s e2e %s' % (sys.argv[1][11:], j['value'], (j.get('e2e') or {}).get('value'), c.get('value'), (c.get('roundtrip_decompress') or {}).get('value'), (c.get('e2e') or {}).get('value')))
        if 'extra_configs' in j: print(json.dumps(j['extra_configs']))
PY
done
