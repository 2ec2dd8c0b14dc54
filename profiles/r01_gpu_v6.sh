set -x
B="python bench.py --no-e2e --no-cpu --steps 3"
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests_v6.log 2>&1; tail -3 gpurun_out/gpu_tests_v6.log
python profiles/pcie_probe.py > gpurun_out/pcie_probe.json 2>&1; cat gpurun_out/pcie_probe.json
$B > gpurun_out/v6_base.log 2>&1; tail -1 gpurun_out/v6_base.log | cut -c1-300
LZF_B200_LIB=build/enc32.so $B --decomp-gib 0.25 > gpurun_out/v6_enc32.log 2>&1
LZF_B200_LIB=build/enc24.so $B --decomp-gib 0.25 > gpurun_out/v6_enc24.log 2>&1
LZF_B200_ENC_SMEM_WARPS=0 $B --decomp-gib 0.25 > gpurun_out/v6_enc28_allglobal.log 2>&1
LZF_B200_LIB=build/dec5.so $B --no-compress > gpurun_out/v6_dec5.log 2>&1
LZF_B200_DEC_CTAS_PER_SM=3 $B --no-compress > gpurun_out/v6_dec3ctas.log 2>&1
for cb in 134217728 268435456 1073741824; do LZF_B200_CHUNK_BYTES=$cb python bench.py --no-cpu --no-compress --steps 3 > gpurun_out/v6_e2e_chunk$cb.log 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lzf" -c 60 --csv --log-file gpurun_out/launches_v6.csv python bench.py --steps 2 --warmup 3 --no-cpu --comp-gib 4 > gpurun_out/launches_v6.log 2>&1
grep -h '^{' gpurun_out/v6_*.log | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l); c = j.get('compress') or {}
    print('dec %.1f e2e %s | comp %s rt %s' % (j['value'], (j.get('e2e') or {}).get('value'), c.get('value'), (c.get('roundtrip_decompress') or {}).get('value')))
"
