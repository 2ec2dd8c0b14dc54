"""Worker of tests/test_multi_rank.py: one rank of a world-size-N gloo job.  Compresses its shard of
the frames through the product binding (bound to the SIMT-emulated build on CPU boxes), gathers the
frames to rank 0 and compares with the oracle's single-process result."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import oracle  # noqa: E402
from lz_fear_b200 import _native as N  # noqa: E402
from lz_fear_b200 import sharding, workloads as W  # noqa: E402


def main():
    lib = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    N.load_library(lib)
    ctx = N.Context(0)
    nframes = 11
    datas = [(W.text(3000 + 700 * i, i).numpy().tobytes() if i % 2 else W.lowent(9000 + 11 * i, i).numpy().tobytes())
             for i in range(nframes)]
    lo, hi = sharding.shard_range(nframes, world, rank)
    mine = datas[lo:hi]
    frames = []
    for d in mine:
        st, fr = ctx.frame_compress(d, block_size=64 << 10)
        assert st == 0
        frames.append(fr)
    sizes = torch.tensor([len(f) for f in frames], dtype=torch.int64)
    per_rank = sharding.all_gather_sizes(sizes)
    assert [int(s.numel()) for s in per_rank] == [sharding.shard_range(nframes, world, r)[1] - sharding.shard_range(nframes, world, r)[0]
                                                  for r in range(world)]
    local = torch.from_numpy(np.frombuffer(b"".join(frames) or b"\0", dtype=np.uint8).copy())
    got = sharding.gather_bytes(local, per_rank, dst=0)
    if rank == 0:
        want = b"".join(oracle.frame_compress(d, block_size=64 << 10)[1] for d in datas)
        assert got.numpy().tobytes() == want, "gathered frames differ from the single-process reference"
        # and decode everything back on rank 0
        pos = 0
        for d, s in zip(datas, torch.cat(per_rank).tolist()):
            st, det, plain, _c = ctx.frame_decompress(got[pos: pos + s].numpy(), cap=len(d) + 16)
            assert (st, plain) == (0, d)
            pos += s
    # decompress direction (§8(e)): rank 0 walks the frame boundaries and scatters contiguous ranges of whole frames;
    # every rank decodes its own and the plaintext is gathered back, variable length
    all_sizes = torch.cat(per_rank).tolist()
    ranges = [sharding.shard_range(nframes, world, r) for r in range(world)]
    totals = [sum(all_sizes[a:b]) for a, b in ranges]
    mine_packed = sharding.scatter_bytes(got if rank == 0 else None, totals, src=0, device="cpu")
    assert sharding.payload_digest(mine_packed) == sharding.payload_digest(local[: totals[rank]])
    pos, plains = 0, []
    for i in range(lo, hi):
        st, det, plain, _c = ctx.frame_decompress(mine_packed[pos: pos + all_sizes[i]].numpy(), cap=len(datas[i]) + 16)
        assert (st, plain) == (0, datas[i])
        plains.append(plain)
        pos += all_sizes[i]
    pl = torch.from_numpy(np.frombuffer(b"".join(plains) or b"\0", dtype=np.uint8).copy())
    psz = [torch.tensor([len(datas[i]) for i in range(a, b)], dtype=torch.int64) for a, b in ranges]
    back = sharding.gather_bytes(pl, psz, dst=0)
    if rank == 0:
        assert back.numpy().tobytes() == b"".join(datas), "scattered -> decoded -> gathered plaintext differs"
        print("MULTI-RANK-OK")
    # the packaged exchange bench.py runs at N > 1 (gather -> scatter -> decode, verified), and its behaviour when ONE rank
    # cannot allocate its buffers: every rank must come back with the same "skipped" verdict instead of waiting forever
    def decode_ok(received):
        p, ok = 0, True
        for i in range(lo, hi):
            st, det, plain, _c = ctx.frame_decompress(received[p: p + all_sizes[i]].numpy(), cap=len(datas[i]) + 16)
            ok = ok and (st, plain) == (0, datas[i])
            p += all_sizes[i]
        return ok
    res = sharding.frames_exchange(local[: totals[rank]], sizes, "cpu", reps=2, decode=decode_ok)
    assert res.get("archive_slices_equal_senders_digests") and res.get("scattered_frames_equal_on_every_rank_and_decoded"), res

    def failing_alloc(n, **kw):
        if rank == 0 and n > totals[0]:
            raise MemoryError("injected: no room for the archive")
        return torch.empty(n, **kw)
    res = sharding.frames_exchange(local[: totals[rank]], sizes, "cpu", alloc=failing_alloc)
    assert "skipped" in res, res
    # ... and the ranks are still in step afterwards
    res = sharding.frames_exchange(local[: totals[rank]], sizes, "cpu", decode=lambda r: rank != 1)
    assert res["scattered_frames_equal_on_every_rank_and_decoded"] is False and res["archive_slices_equal_senders_digests"], res
    if rank == 0:
        print("EXCHANGE-OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
