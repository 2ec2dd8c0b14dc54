set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests_v12.log 2>&1; tail -3 gpurun_out/gpu_tests_v12.log
B="timeout 300 python bench.py --no-compress --no-cpu --no-e2e --steps 3"
LZF_B200_LIB=build/dec_w2k.so $B > gpurun_out/v12_w2k.log 2>&1
LZF_B200_LIB=build/dec_w512.so $B > gpurun_out/v12_w512.log 2>&1
$B --no-xxh > gpurun_out/v12_noxxh.log 2>&1
timeout 300 python profiles/text_decode_probe.py > gpurun_out/v12_text.log 2>&1
LZF_B200_LIB=build/dec_w2k.so timeout 300 python profiles/text_decode_probe.py > gpurun_out/v12_text_w2k.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_blocks -s 3 -c 1 -o gpurun_out/prof_decode_v12 timeout 600 python bench.py --no-compress --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_dec_v12.log 2>&1
NB=1024 ncu --set full --clock-control none --import-source on -k regex:decode_blocks -s 2 -c 1 -o gpurun_out/prof_decode_text_v12 timeout 600 python profiles/text_decode_probe.py > gpurun_out/ncu_dec_text_v12.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"blocks_kernel|xxh32|frame_|stage_dict" -c 80 --csv --log-file gpurun_out/launches_v12.csv timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/launches_v12.log 2>&1
timeout 900 python bench.py > gpurun_out/v12_bench.log 2>&1
for f in gpurun_out/v12_*.log; do python - "$f" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        j = json.loads(l); c = j.get('compress') or {}
        print('%-26s dec %.1f e2e %s | comp %s rt %s e2e %s' % (sys.argv[1][11:], j['value'], (j.get('e2e') or {}).get('value'), c.get('value'), (c.get('roundtrip_decompress') or {}).get('value'), (c.get('e2e') or {}).get('value')))
    elif l.startswith('text decode'): print(sys.argv[1][11:], l.strip())
PY
done
