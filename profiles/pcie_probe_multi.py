"""Host<->device copy ceiling of an N-GPU box with ALL ranks copying at once (run under torchrun).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 profiles/pcie_probe_multi.py

Every rank moves 1 GiB pinned -> device, device -> pinned, and both directions at once, all ranks released together by a
barrier; rank 0 prints per-rank and aggregate GB/s.  PROBE_AFFINITY=1 first binds the process (and therefore the pages
its pinned allocation touches) to the CPUs `nvidia-smi topo -m` lists as local to its GPU."""
import json
import os
import re
import subprocess
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))


def gpu_cpu_affinity(idx):
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
        for ln in out.splitlines():
            if ln.startswith("GPU%d" % idx + "\t") or ln.startswith("GPU%d " % idx):
                m = re.findall(r"(\d+-\d+(?:,\d+-\d+)*)", ln)
                if m:
                    cpus = set()
                    for part in m[0].split(","):
                        a, b = part.split("-")
                        cpus.update(range(int(a), int(b) + 1))
                    return cpus, ln.strip()
    except Exception:
        pass
    return None, None


aff_note = "none"
if os.environ.get("PROBE_AFFINITY") == "1":
    cpus, line = gpu_cpu_affinity(local)
    if cpus:
        os.sched_setaffinity(0, cpus)
        aff_note = "cpus %d-%d" % (min(cpus), max(cpus))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_in.fill_(1)
h_out = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out.fill_(2)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(kind, reps=4):
    def once():
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    once(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        return [float(x.item()) for x in allt]
    return [dt]


res = {"n_gpus": world, "GiB_per_rank_per_direction": 1, "affinity": aff_note}
for kind in ("h2d", "d2h", "both"):
    ts = run(kind)
    mult = 2 if kind == "both" else 1
    res[kind] = {"per_rank_GBps": [round(mult * n / 1e9 / t, 1) for t in ts], "aggregate_GBps": round(world * mult * n / 1e9 / max(ts), 1)}
if rank == 0:
    try:
        res["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000]
        res["numa"] = subprocess.run(["bash", "-c", "lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"], capture_output=True, text=True).stdout
    except Exception:
        pass
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
