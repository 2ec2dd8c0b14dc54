"""H2D ceilings for the sliced feed: one contiguous copy vs strided copies (width = slice, pitch = block) of the same
bytes from pinned memory, timed with CUDA events through the runtime API (cuda-python)."""
import json
import torch
from cuda.bindings import runtime as rt

B, rows = 4 << 20, 1024                      # 4 GiB
n = B * rows
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
err, stream = rt.cudaStreamCreate()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
ts = torch.cuda.ExternalStream(int(stream))
H2D = rt.cudaMemcpyKind.cudaMemcpyHostToDevice


def timed(fn):
    fn(); rt.cudaStreamSynchronize(stream)
    e0.record(ts); fn(); e1.record(ts); rt.cudaStreamSynchronize(stream)
    return n / 1e9 / (e0.elapsed_time(e1) / 1e3)


out = {"GiB": n / 2**30}
out["contiguous_GBps"] = timed(lambda: rt.cudaMemcpyAsync(d.data_ptr(), h.data_ptr(), n, H2D, stream))
for slice_ in (64 << 10, 256 << 10, 1 << 20):
    def f():
        for k in range(B // slice_):
            rt.cudaMemcpy2DAsync(d.data_ptr() + k * slice_, B, h.data_ptr() + k * slice_, B, slice_, rows, H2D, stream)
    out["slices_%dKiB_GBps" % (slice_ >> 10)] = timed(f)
print(json.dumps(out))
