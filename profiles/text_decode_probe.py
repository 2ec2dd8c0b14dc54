"""Decode of oracle-identical compressed TEXT blocks (4 MiB, config 3 round trip) for ncu captures:
    ncu --set full -k regex:decode_blocks -s 2 -c 1 -o gpurun_out/prof_decode_text python profiles/text_decode_probe.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lz_fear_b200 import _native as N
from lz_fear_b200 import workloads as W

B, nb = 4 << 20, int(os.environ.get("NB", "2048"))
ctx = N.Context(0)
src = W.TextSource(device="cuda")
data = torch.cat([src.make(16 * B) for _ in range(nb // 16)])
off = torch.arange(nb, device="cuda", dtype=torch.int64) * B
ln = torch.full((nb,), B, dtype=torch.int32, device="cuda")
comp = torch.empty(nb * B, dtype=torch.uint8, device="cuda")
clen = torch.zeros(nb, dtype=torch.int32, device="cuda")
st = torch.zeros(nb, dtype=torch.int32, device="cuda")
xx = torch.zeros(nb, dtype=torch.int32, device="cuda")
ctx.compress_blocks(data, off, ln, nb, comp, off, None, clen, st, None, None, max_block_len=B)
torch.cuda.synchronize()
back = torch.empty_like(data)
olen = torch.zeros_like(clen)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(5):
    if i == 2:
        e0.record()
    ctx.decompress_blocks(comp, off, clen, nb, back, off, ln, ln, olen, st, xx)
e1.record()
torch.cuda.synchronize()
assert int(st.abs().sum()) == 0 and torch.equal(back, data)
print("text decode: %.1f GiB/s (%d x 4 MiB blocks)" % (3 * nb * B / 2**30 / (e0.elapsed_time(e1) / 1e3), nb))
