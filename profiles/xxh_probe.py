"""Times lzf_xxh32_ranges (frame content checksums) on one long range, a few, and many: the chains are serial per range,
so a single 64 MiB frame is the worst case.  LZF_B200_LIB selects the build."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lz_fear_b200 import _native as N  # noqa: E402

ctx = N.Context(0)
dev = "cuda"
data = torch.randint(0, 256, (1 << 30,), dtype=torch.uint8, device=dev)
res = {"lib": os.environ.get("LZF_B200_LIB", "default")}
for name, nr, ln in (("1x64MiB", 1, 64 << 20), ("16x4MiB", 16, 4 << 20), ("16x64MiB", 16, 64 << 20), ("4096x256KiB", 4096, 256 << 10)):
    off = torch.arange(nr, device=dev, dtype=torch.int64) * ln
    lens = torch.full((nr,), ln, dtype=torch.int64, device=dev)
    out = torch.zeros(nr, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        ctx.xxh32_ranges(data, off, lens, nr, out, stream=s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ctx.xxh32_ranges(data, off, lens, nr, out, stream=s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    res[name] = {"ms": round(ms, 3), "GBps_per_range": round(ln / 1e9 / (ms / 1e3), 3)}
    if nr == 1:
        import xxhash
        assert int(out[0].item()) & 0xFFFFFFFF == xxhash.xxh32(data[:ln].cpu().numpy().tobytes()).intdigest()
print(json.dumps(res))
