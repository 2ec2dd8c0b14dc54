"""Host<->device copy ceilings of the box (pinned memory, CUDA events): H2D alone, D2H alone, both at once."""
import json
import torch

n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=4):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / 1e3


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d(); d2h()


for s in (s1, s2):
    s.wait_stream(torch.cuda.current_stream())
out = {"GiB": 1, "h2d_GBps": n / 1e9 / timed(h2d), "d2h_GBps": n / 1e9 / timed(d2h)}
t = timed(both)
out["both_each_GBps"] = n / 1e9 / t
out["both_sum_GBps"] = 2 * n / 1e9 / t
print(json.dumps(out))
