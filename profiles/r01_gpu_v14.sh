set -x
B="timeout 300 python bench.py --no-compress --no-cpu --no-e2e --steps 3"
for v in head_v12 v14_hash_only v14_hash_noprefetch v14_unroll_only; do
  LZF_B200_LIB=build/$v.so $B > gpurun_out/v14_$v.log 2>&1
  NB=4096 LZF_B200_LIB=build/$v.so timeout 300 python profiles/text_decode_probe.py > gpurun_out/v14_text_$v.log 2>&1
done
for f in gpurun_out/v14_*.log; do python - "$f" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        j = json.loads(l); print('%-36s dec %.1f' % (sys.argv[1][11:], j['value']))
    elif l.startswith('text decode'): print(sys.argv[1][11:], l.strip())
PY
done
