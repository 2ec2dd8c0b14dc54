set -x
B="timeout 300 python bench.py --no-compress --no-cpu --steps 3"
$B --no-e2e > gpurun_out/v11_base.log 2>&1
LZF_B200_LIB=build/dec_h1.so $B --no-e2e > gpurun_out/v11_h1.log 2>&1
LZF_B200_LIB=build/dec_h2.so $B --no-e2e > gpurun_out/v11_h2.log 2>&1
LZF_B200_LIB=build/dec_h2.so LZF_B200_DEC_CTAS_PER_SM=3 $B --no-e2e > gpurun_out/v11_h2_3ctas.log 2>&1
for cb in 134217728 268435456; do LZF_B200_CHUNK_BYTES=$cb $B > gpurun_out/v11_e2e_chunk$cb.log 2>&1; done
for f in gpurun_out/v11_*.log; do python - "$f" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        j = json.loads(l)
        print('%-32s dec %.1f e2e %s' % (sys.argv[1][11:], j['value'], (j.get('e2e') or {}).get('value')))
PY
done
