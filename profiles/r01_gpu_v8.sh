set -x
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests_v8.log 2>&1; tail -3 gpurun_out/gpu_tests_v8.log
python bench.py > gpurun_out/v8_bench.log 2>&1; tail -c 1500 gpurun_out/v8_bench.log
for sl in 0 65536 1048576; do LZF_B200_FEED_SLICE=$sl python bench.py --decomp-gib 0.25 --no-cpu --steps 2 > gpurun_out/v8_feed$sl.log 2>&1; done
python bench.py --extra --no-e2e --decomp-gib 0.25 --comp-gib 1 --steps 3 > gpurun_out/v8_extra.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_blocks -s 3 -c 1 -o gpurun_out/prof_decode_v8 python bench.py --no-compress --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_dec_v8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode_blocks -s 3 -c 1 -o gpurun_out/prof_encode_v8 python bench.py --decomp-gib 0.25 --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_enc_v8.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"blocks_kernel|xxh32|frame_|stage_dict" -c 120 --csv --log-file gpurun_out/launches_v8.csv python bench.py --steps 2 --warmup 3 --no-cpu --comp-gib 4 > gpurun_out/launches_v8.log 2>&1
for f in gpurun_out/v8_*.log; do python - "$f" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        j = json.loads(l); c = j.get('compress') or {}
        print('%-44s dec %.1f e2e %s | comp %s rt %s e2e %s' % (sys.argv[1][11:], j['value'], (j.get('e2e') or {}).get('value'), c.get('value'), (c.get('roundtrip_decompress') or {}).get('value'), (c.get('e2e') or {}).get('value')))
        if 'extra_configs' in j: print(json.dumps(j['extra_configs']))
PY
done
