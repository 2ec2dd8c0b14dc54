"""Frame-granular gather / scatter over NCCL (sharding.gather_bytes / scatter_bytes) on its own: every rank holds
~PROBE_GIB GiB of 'frames'; rank 0 gathers them and scatters them back.  Run under torchrun; NCCL_* tuning comes from the
environment of the launch."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lz_fear_b200 import sharding  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
gib = float(os.environ.get("PROBE_GIB", "2"))
n = int(gib * (1 << 30)) - 4096 * rank
payload = torch.randint(0, 256, (n,), dtype=torch.uint8, device=dev)
sizes = torch.tensor([n // 4, n - n // 4], dtype=torch.int64, device=dev)
per_rank = sharding.all_gather_sizes(sizes)
totals = [int(x.sum().item()) for x in per_rank]
archive = torch.empty(sum(totals), dtype=torch.uint8, device=dev) if rank == 0 else None
back = torch.empty(n, dtype=torch.uint8, device=dev)
for _ in range(2):
    sharding.gather_bytes(payload, per_rank, dst=0, out=archive)
    sharding.scatter_bytes(archive, totals, src=0, out=back)
torch.cuda.synchronize(); dist.barrier()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
K = 5
e0.record()
for _ in range(K):
    sharding.gather_bytes(payload, per_rank, dst=0, out=archive)
e1.record()
for _ in range(K):
    sharding.scatter_bytes(archive, totals, src=0, out=back)
e2.record()
torch.cuda.synchronize(); dist.barrier()
t = torch.tensor([e0.elapsed_time(e1) / K, e1.elapsed_time(e2) / K], dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert torch.equal(back, payload)
if rank == 0:
    remote = sum(totals) - totals[0]
    print(json.dumps({"n_gpus": world, "GiB_per_rank": gib, "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")},
                      "gather_ms": t[0].item(), "scatter_ms": t[1].item(),
                      "gather_GiB_per_s_into_rank0": remote / (1 << 30) / (t[0].item() / 1e3),
                      "scatter_GiB_per_s_out_of_rank0": remote / (1 << 30) / (t[1].item() / 1e3)}))
dist.destroy_process_group()
