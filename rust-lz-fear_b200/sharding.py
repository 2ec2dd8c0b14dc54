"""Frame-granular sharding over ranks.

Independent-block frames never communicate, so a batch of frames is split into contiguous ranges of
WHOLE frames, one per rank (each rank's content-checksum chains stay local), and the only exchange
steps move whole frames at the edges of the path:

  compress    all-gather of per-frame byte counts, then a variable-length GATHER of the frame bytes to a root
  decompress  the root walks the frame boundaries, SCATTERS contiguous ranges of whole frames, every rank
              decodes its own, and the plaintext comes back with a (fixed- or variable-length) gather

NCCL has no gatherv / scatterv: both are ONE grouped batch of point-to-point operations
(`batch_isend_irecv` = ncclGroupStart .. ncclGroupEnd), so the transfers of all peers run side by side over
NVLink instead of one after the other.  `torch.distributed` is the plumbing (NCCL on the GPUs, gloo in the
CPU tests).
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


def _run(ops):
    if ops:
        for q in dist.batch_isend_irecv(ops):
            q.wait()


def _peer(group, r):
    return r if group is None else dist.get_global_rank(group, r)


def gather_bytes(local, sizes_per_rank, dst=0, group=None, out=None):
    """Variable-length gather.  `local` = this rank's packed uint8 payload, `sizes_per_rank[r]` = byte counts
    of rank r's items; returns the concatenation in rank order on `dst` (into `out` when given), None elsewhere."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    totals = [int(s.sum().item()) if torch.is_tensor(s) else int(s) for s in sizes_per_rank]
    if rank == dst:
        if out is None:
            out = torch.empty(sum(totals), dtype=torch.uint8, device=local.device)
        pos, ops = 0, []
        for r in range(world):
            view = out[pos: pos + totals[r]]
            if r == dst:
                view.copy_(local[: totals[r]])
            elif totals[r]:
                ops.append(dist.P2POp(dist.irecv, view, _peer(group, r), group))
            pos += totals[r]
        _run(ops)
        return out[: sum(totals)]
    if totals[rank]:
        _run([dist.P2POp(dist.isend, local[: totals[rank]].contiguous(), _peer(group, dst), group)])
    return None


def scatter_bytes(packed, totals, src=0, group=None, device=None, out=None):
    """Variable-length scatter: rank `src` holds `packed` = the concatenation, in rank order, of every rank's byte
    range (`totals[r]` bytes for rank r, known on every rank); every rank gets its own range back."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    totals = [int(t) for t in totals]
    if rank == src:
        pos, ops, mine = 0, [], None
        for r in range(world):
            view = packed[pos: pos + totals[r]]
            if r == src:
                mine = view
            elif totals[r]:
                ops.append(dist.P2POp(dist.isend, view, _peer(group, r), group))
            pos += totals[r]
        _run(ops)
        if out is not None:
            out[: totals[rank]].copy_(mine)
            return out[: totals[rank]]
        return mine
    if out is None:
        out = torch.empty(totals[rank], dtype=torch.uint8, device=device if device is not None else (packed.device if packed is not None else None))
    if totals[rank]:
        _run([dist.P2POp(dist.irecv, out[: totals[rank]], _peer(group, src), group)])
    return out[: totals[rank]]


def payload_digest(t, chunk=32 << 20):
    """A cheap order-sensitive digest of a uint8 tensor computed where it lives (exchange verification: every rank
    digests what it sent, the receiver digests each slice it got): (sum of bytes, sum of byte * (index mod 65521 + 1)),
    both modulo 2^61 - 1.  Works in chunks so that the int64 temporaries stay small next to multi-GiB payloads."""
    if t is None or t.numel() == 0:
        return (0, 0)
    m = (1 << 61) - 1
    s0 = s1 = 0
    for lo in range(0, t.numel(), chunk):
        x = t[lo: lo + chunk].to(torch.int64)
        w = (torch.arange(lo, lo + x.numel(), device=t.device, dtype=torch.int64) % 65521) + 1
        s0 = (s0 + int(x.sum().item())) % m
        s1 = (s1 + int((x * w).sum().item())) % m
        del x, w
    return (s0, s1)


def all_ranks_ok(ok, device, group=None):
    """Collective AND: every rank calls it and every rank gets the same answer.  The exchange steps allocate their large
    buffers on the root only; a rank that fails locally must not leave the others waiting inside a collective it never
    joins, so every local step that can fail is followed by this agreement before the next collective."""
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(t.item())


class _WallTimer:
    def __init__(self):
        import time
        self._t, self._time = 0.0, time

    def start(self):
        self._t = self._time.perf_counter()

    def stop_ms(self):
        return (self._time.perf_counter() - self._t) * 1e3


def frames_exchange(packed, frame_sizes, device, reps=1, root=0, group=None, decode=None, timer=None, barrier=None,
                    alloc=torch.empty):
    """The full frame-granular exchange of SURVEY §8(e) with its verification, safe against rank-local failures:

        all-gather of the per-frame sizes -> GATHER of every rank's packed frames into one archive on `root`
        -> `root` cuts the archive at frame boundaries and SCATTERS every rank's range back -> decode(received) on every rank

    `packed` = this rank's frames back to back (uint8, on `device`), `frame_sizes` = their lengths (int64 tensor).
    `decode(received) -> bool` checks what came back (e.g. decompress and compare with the plaintext).  `alloc` is
    torch.empty (tests inject a failing one).  Returns a dict on every rank; {"skipped": ...} when some rank could not
    allocate — agreed on by all ranks, so nobody waits in a collective."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    timer = timer or _WallTimer()
    sync = barrier or (lambda: dist.barrier(group=group))
    per_rank = all_gather_sizes(frame_sizes, group=group)
    totals = [int(x.sum().item()) for x in per_rank]
    err, archive, recv = None, None, None
    try:
        if rank == root:
            archive = alloc(sum(totals), dtype=torch.uint8, device=device)
        recv = alloc(totals[rank], dtype=torch.uint8, device=device)
    except Exception as e:          # noqa: BLE001 — reported, and agreed on below
        err = "%s: %s" % (type(e).__name__, str(e)[:300])
    if not all_ranks_ok(err is None, device, group):
        return {"skipped": "a rank could not allocate its exchange buffers (root needs %d bytes)" % sum(totals), "this_rank_error": err}
    for _ in range(2):                                          # NCCL sets its point-to-point channels up on first use
        gather_bytes(packed, per_rank, dst=root, group=group, out=archive)
        scatter_bytes(archive, totals, src=root, group=group, out=recv)
    sync()
    timer.start()
    for _ in range(reps):
        per_rank = all_gather_sizes(frame_sizes, group=group)
        gather_bytes(packed, per_rank, dst=root, group=group, out=archive)
    g_ms = timer.stop_ms() / reps
    sync()
    timer.start()
    for _ in range(reps):
        scatter_bytes(archive, totals, src=root, group=group, out=recv)
    s_ms = timer.stop_ms() / reps
    sync()
    digs = [None] * world
    dist.all_gather_object(digs, payload_digest(packed), group=group)
    ok_archive = True
    if rank == root:
        pos = 0
        for r in range(world):
            ok_archive = ok_archive and payload_digest(archive[pos:pos + totals[r]]) == tuple(digs[r])
            pos += totals[r]
    ok_local = bool(torch.equal(recv, packed[: totals[rank]]))
    if decode is not None:
        try:
            ok_local = ok_local and bool(decode(recv))
        except Exception as e:      # noqa: BLE001
            ok_local, err = False, "%s: %s" % (type(e).__name__, str(e)[:200])
    ok_all = all_ranks_ok(ok_local, device, group)
    ok_archive = all_ranks_ok(ok_archive, device, group)       # the root's verdict, known everywhere
    return {"gather_ms": g_ms, "scatter_ms": s_ms, "totals": totals, "archive_slices_equal_senders_digests": ok_archive,
            "scattered_frames_equal_on_every_rank_and_decoded": ok_all, "this_rank_error": err}
