"""Dry run of bench.py's control flow WITHOUT a GPU (test infrastructure; slow, ~7 minutes per run).

bench.py itself refuses to run without a CUDA device.  This tool executes its unmodified logic on the CPU SIMT-emulator
build of the C ABI (tests/simt) with torch.cuda faked (events = wall clock, "cuda" tensors on the CPU, pinning a no-op)
and the sizes shrunk (64 KiB instead of 4 MiB blocks, 1 MiB instead of 64 MiB for the one-file section), so that every
section — e2e pipelines, copy ceilings, device frame calls, configs 4 and 5, the one-file and streaming sections, and at
N > 1 the NCCL sections over gloo (gather, the full config-4 exchange through sharding.frames_exchange) — runs once and
the JSON line is produced.  It checks bench.py's Python, not performance; the numbers it prints are meaningless.

    python tests/bench_dry_run.py                                   # N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \\
        tests/bench_dry_run.py --gpus 2                             # N = 2, gloo
"""
import contextlib
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "simt")])

import torch  # noqa: E402
from lz_fear_b200 import _native as N  # noqa: E402

N.load_library(os.path.join(ROOT, "tests", "simt", "libsimt_lzfear.so"))

_real_device = torch.device


class FakeEvent:
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self, *a):
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


class FakeStream:
    cuda_stream = 0


class DeviceShim:
    def __call__(self, *a, **k):
        return _real_device("cpu") if a and a[0] == "cuda" else _real_device(*a, **k)

    def __instancecheck__(self, inst):
        return isinstance(inst, _real_device)


torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.current_stream = lambda *a, **k: FakeStream()
torch.cuda.Event = FakeEvent
torch.cuda.Stream = lambda *a, **k: FakeStream()
torch.cuda.stream = lambda s: contextlib.nullcontext()
torch.cuda.empty_cache = lambda: None
torch.cuda.mem_get_info = lambda *a: (1 << 40, 1 << 40)
torch.cuda.memory_allocated = lambda *a: 0
torch.cuda.memory_reserved = lambda *a: 0
torch.Tensor.pin_memory = lambda self, *a, **k: self
torch.device = DeviceShim()

src = open(os.path.join(ROOT, "bench.py")).read()
for old, new in (("BLOCK3 = 4 << 20", "BLOCK3 = 65536"), ("make(64 << 20)", "make(1 << 20)"),
                 ('dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))',
                  'dist.init_process_group("gloo", timeout=datetime.timedelta(seconds=300))'),
                 ("ctx = N.Context(local_rank)", "ctx = N.Context(0)")):        # the emulator has one device
    assert src.count(old) == 1, old
    src = src.replace(old, new)
sys.argv = ["bench.py", "--steps", "1", "--warmup", "1", "--decomp-gib", "0.004", "--comp-gib", "0.0001", "--mixed-gib", "0.0001",
            "--lowent-gib", "0.0005"] + sys.argv[1:]
exec(compile(src, os.path.join(ROOT, "bench.py"), "exec"), {"__name__": "__main__", "__file__": os.path.join(ROOT, "bench.py")})
