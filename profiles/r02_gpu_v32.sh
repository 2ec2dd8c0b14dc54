set -x
# v32: (a) content-checksum kernel variants on long ranges, (b) decoder with / without the early vectorised gather
for v in default h_w2k_d2_a16 h_w4k_d3 h_w2k_d4 h_w4k_d2 h_w8k_d2; do
  if [ $v = default ]; then unset LZF_B200_LIB; else export LZF_B200_LIB=build/$v.so; fi
  timeout 300 python profiles/xxh_probe.py 2>&1 | tail -1 | tee -a gpurun_out/v32_xxh.jsonl
done
unset LZF_B200_LIB
for v in default v_dec_base default v_dec_base; do
  if [ $v = default ]; then unset LZF_B200_LIB; else export LZF_B200_LIB=build/$v.so; fi
  timeout 600 python bench.py --no-e2e --no-cpu --no-extra --comp-gib 1 --steps 10 --warmup 3 > gpurun_out/v32_dec_$v.log 2>&1
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/v32_dec_$v.log') if l.startswith('{')][-1])
print('$v', 'decode', d['value'], 'text decode', d['compress']['roundtrip_decompress']['value'])
PY
done
