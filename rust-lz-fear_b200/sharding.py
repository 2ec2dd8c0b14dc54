"""Frame-granular sharding over ranks.

Independent-block frames never communicate, so a batch of frames is split into contiguous ranges of
WHOLE frames, one per rank (each rank's content-checksum chains stay local), and the only exchange
step is moving the results: an all-gather of per-frame byte counts followed by a variable-length
gather of the frame bytes to a root.  `torch.distributed` is the plumbing (NCCL over NVLink on the
GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(nitems, world_size, rank):
    """Contiguous [lo, hi) of `nitems` for `rank`: the first (nitems % world_size) ranks get one extra."""
    base, extra = divmod(int(nitems), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_sizes(local_sizes, group=None):
    """local_sizes: int64 tensor (n_local,) on the communication device -> list of per-rank tensors."""
    world = dist.get_world_size(group)
    counts = [torch.zeros(1, dtype=torch.int64, device=local_sizes.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local_sizes.numel()], dtype=torch.int64, device=local_sizes.device), group=group)
    out = [torch.zeros(int(c.item()), dtype=torch.int64, device=local_sizes.device) for c in counts]
    # all_gather needs equal shapes: pad to the longest
    m = max(int(c.item()) for c in counts)
    padded = torch.zeros(m, dtype=torch.int64, device=local_sizes.device)
    padded[: local_sizes.numel()] = local_sizes
    bufs = [torch.zeros(m, dtype=torch.int64, device=local_sizes.device) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    for r in range(world):
        out[r] = bufs[r][: int(counts[r].item())].clone()
    return out


def gather_bytes(local, sizes_per_rank, dst=0, group=None):
    """Variable-length gather (NCCL has no gatherv: grouped send/recv).  `local` = this rank's packed
    uint8 payload; returns the concatenation in rank order on `dst`, None elsewhere."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    totals = [int(s.sum().item()) for s in sizes_per_rank]
    if rank == dst:
        out = torch.empty(sum(totals), dtype=torch.uint8, device=local.device)
        pos, reqs = 0, []
        for r in range(world):
            view = out[pos: pos + totals[r]]
            if r == dst:
                view.copy_(local[: totals[r]])
            elif totals[r]:
                reqs.append(dist.irecv(view, src=r, group=group))
            pos += totals[r]
        for q in reqs:
            q.wait()
        return out
    if totals[rank]:
        dist.send(local[: totals[rank]].contiguous(), dst=dst, group=group)
    return None
