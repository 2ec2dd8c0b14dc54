set -x
B="python bench.py --no-e2e --no-cpu --steps 3 --decomp-gib 0.25"
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests_v7.log 2>&1; tail -3 gpurun_out/gpu_tests_v7.log
python bench.py --no-cpu --steps 3 > gpurun_out/v7_base_e2e.log 2>&1
LZF_B200_ENC_SMEM_WARPS=0 $B > gpurun_out/v7_enc28_g.log 2>&1
LZF_B200_ENC_SMEM_WARPS=13 $B > gpurun_out/v7_enc28_s13.log 2>&1
LZF_B200_ENC_SMEM_WARPS=0 LZF_B200_ENC_U32=1 $B > gpurun_out/v7_enc28_g_u32.log 2>&1
LZF_B200_LIB=build/enc32.so LZF_B200_ENC_SMEM_WARPS=0 $B > gpurun_out/v7_enc32_g.log 2>&1
LZF_B200_LIB=build/enc32.so LZF_B200_ENC_SMEM_WARPS=0 LZF_B200_ENC_U32=1 $B > gpurun_out/v7_enc32_g_u32.log 2>&1
for cb in 268435456 1073741824; do LZF_B200_CHUNK_BYTES=$cb python bench.py --no-cpu --no-compress --steps 3 > gpurun_out/v7_e2e_chunk$cb.log 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"blocks_kernel|xxh32|frame_|stage_dict" -c 80 --csv --log-file gpurun_out/launches_v7.csv python bench.py --steps 2 --warmup 3 --no-cpu --comp-gib 4 > gpurun_out/launches_v7.log 2>&1
for f in gpurun_out/v7_*.log; do python - "$f" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        j = json.loads(l); c = j.get('compress') or {}
        print('%-40s dec %.1f e2e %s | comp %s rt %s e2e %s' % (sys.argv[1], j['value'], (j.get('e2e') or {}).get('value'), c.get('value'), (c.get('roundtrip_decompress') or {}).get('value'), (c.get('e2e') or {}).get('value')))
PY
done
