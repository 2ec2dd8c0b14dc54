set -x
B="timeout 600 python bench.py --decomp-gib 0.25 --no-cpu --steps 2"
LZF_B200_TRACE=1 $B > gpurun_out/v18_base.log 2>&1
LZF_B200_TRACE=1 LZF_B200_LIB=build/enc64.so $B > gpurun_out/v18_enc64.log 2>&1
LZF_B200_TRACE=1 LZF_B200_LIB=build/enc64.so LZF_B200_ENC_ROOM_KB=36 $B > gpurun_out/v18_enc64_room.log 2>&1
LZF_B200_TRACE=1 LZF_B200_ENC_ROOM_KB=36 $B > gpurun_out/v18_base_room.log 2>&1
for f in gpurun_out/v18_*.log; do grep "lzf trace" $f | tail -1; python - "$f" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        j = json.loads(l); c = j.get('compress') or {}
        print('%-26s comp %s e2e %s' % (sys.argv[1][11:], c.get('value'), (c.get('e2e') or {}).get('value')))
PY
done
