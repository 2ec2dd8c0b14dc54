#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 LZ4 block codec (BASELINE.json / SURVEY.md §8(d)).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic input.  Per GPU (weak scaling):

  headline  config 2 — DEcompress 4 GiB of independent 64 KiB blocks ("seq50": ~50 % literal bytes,
            ~50 % short matches).  `value` = plaintext GiB/s of the batched block call
            (lzf_decompress_blocks, XXH32 fused) with inputs resident in HBM, CUDA events.
            `e2e` = the same blocks wrapped in LZ4 frames (16 blocks each, content checksum on),
            decoded through the host-buffer frame call (lzf_frames_decompress): pinned host frames
            -> H2D -> walk + decode + checksum kernels -> D2H plaintext, all inside the timed region.
  also      config 3 — compress 4 MiB text-like blocks with default CompressionSettings (frames of
            16 blocks), reported under "compress" with its own value / e2e / roofline.

`--impl reference` times the reference's algorithm on the host cores (the CPU oracle port — the
crate is Rust and cannot be built here) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GiB = float(1 << 30)
BLOCK2 = 65536                 # config 2 block size
BLOCKS_PER_FRAME2 = 16
BLOCK3 = 4 << 20               # config 3 block size (CompressionSettings::default)
BLOCKS_PER_FRAME3 = 16


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--decomp-gib", type=float, default=4.0, help="config-2 plaintext GiB per GPU")
    ap.add_argument("--comp-gib", type=float, default=16.0, help="config-3 plaintext GiB per GPU")
    ap.add_argument("--no-compress", action="store_true", help="skip the config-3 compress section")
    ap.add_argument("--no-decompress-e2e", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-xxh", action="store_true", help="experiment: skip the fused XXH32 epilogue")
    ap.add_argument("--gather", action="store_true", help="(default when N > 1) also time the frame-granular NCCL gather of compressed frames to rank 0")
    ap.add_argument("--no-gather", action="store_true", help="N > 1: skip the NCCL gather section")
    ap.add_argument("--extra", action="store_true", help="(default) also run BASELINE configs 4 (mixed-entropy frames) and 5 (large hash tables)")
    ap.add_argument("--no-extra", action="store_true", help="skip BASELINE configs 4 and 5")
    ap.add_argument("--mixed-gib", type=float, default=8.0, help="config-4 plaintext GiB per GPU (64 GiB over 8 GPUs)")
    ap.add_argument("--lowent-gib", type=float, default=1.0, help="config-5 plaintext GiB")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smax = float(p[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_traffic(kernel, nblocks):
    """(dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` on exactly this workload, where it comes
    from): read from the committed `ncu --set full` capture (profiles/ncu_traffic.json) — a number taken under the
    profiler on an earlier run of the same build, NOT measured in this run; (None, None) when no capture matches."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = t.get(kernel)
        if e and int(e.get("nblocks", -1)) == int(nblocks):
            return float(e["dram_bytes_per_launch"]), "ncu capture %s (build %s), not measured in this run" % (e.get("capture", "?"), e.get("build", "?"))
    except Exception:
        pass
    return None, None


def measured_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# CPU reference legs (oracle port; test/bench infrastructure)
# ---------------------------------------------------------------------------------------------
def cpu_decompress_sample(comp_h, off_h, len_h, nblocks, nthreads, reps=3, out=None):
    """C port of the lz-fear decode loop (oracle/, built -march=native on this host), one block per task."""
    import oracle
    if out is None:
        out = np.empty(nblocks * BLOCK2, dtype=np.uint8)
    out_off = np.arange(nblocks, dtype=np.uint64) * BLOCK2
    cap = np.full(nblocks, BLOCK2, dtype=np.uint32)
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        olen, st = oracle.decompress_blocks_mt(comp_h, off_h, len_h, out, out_off, cap, cap, nthreads=nthreads, native=True)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
        assert not st.any() and (olen == BLOCK2).all()
    return nblocks * BLOCK2 / GiB / best, out


def cpu_compress_sample(plain_h, nblocks, nthreads, reps=2, block=None):
    """C port of lz-fear compress2 (fresh table per block, cap = block length), one block per task."""
    import oracle
    block = block or BLOCK3
    off = np.arange(nblocks, dtype=np.uint64) * block
    ln = np.full(nblocks, block, dtype=np.uint32)
    out = np.empty(nblocks * block, dtype=np.uint8)
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        olen, st = oracle.compress_blocks_mt(plain_h, off, ln, out, off, nthreads=nthreads, native=True)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return nblocks * block / GiB / best, olen, out, st


def liblz4_sample(compress, inp, off, ln, out_block, nthreads, reps=3):
    """Second CPU bar: C lz4 1.9.x on the same blocks and thread pool; None when liblz4.so.1 is missing."""
    import oracle
    nb = len(ln)
    out = np.empty(nb * out_block, dtype=np.uint8)
    out_off = np.arange(nb, dtype=np.uint64) * out_block
    cap = np.full(nb, out_block, dtype=np.uint32)
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        r = oracle.liblz4_blocks_mt(compress, inp, off, ln, out, out_off, cap, nthreads=nthreads)
        dt = time.perf_counter() - t
        if r is None:
            return None
        best = dt if best is None else min(best, dt)
    plain = int(ln.astype(np.uint64).sum()) if compress else int(r[0].astype(np.uint64).sum())
    return {"value": plain / GiB / best, "unit": "GiB/s", "cores": nthreads, "kind": "liblz4 1.9.x (C lz4, not lz-fear)",
            "failed_blocks": int(r[1].sum())}


def run_reference(args):
    """--impl reference: the reference's algorithm (C port, -march=native, all host threads) on bounded samples of the
    same workloads: config 2 decompress is the line's value, config 3 compress rides along under "compress"."""
    import torch
    from lz_fear_b200 import workloads as W
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nb = 16384                                    # 1 GiB of config-2 plaintext per step
    comp, off, ln = W.seq50_blocks(nb, device="cpu")
    comp_h = comp.numpy(); off_h = off.numpy().astype(np.uint64); len_h = ln.numpy().astype(np.uint32)
    out_h = np.empty(nb * BLOCK2, dtype=np.uint8)      # one output buffer for every step: the warm-up takes its page faults
    for _ in range(max(args.warmup, 1)):
        cpu_decompress_sample(comp_h, off_h, len_h, nb, cores, reps=1, out=out_h)
    t = time.perf_counter()
    for _ in range(args.steps):
        v, _o = cpu_decompress_sample(comp_h, off_h, len_h, nb, cores, reps=1, out=out_h)
    dt = time.perf_counter() - t
    value = args.steps * nb * BLOCK2 / GiB / dt
    lz4d = liblz4_sample(False, comp_h, off_h, len_h, BLOCK2, cores)
    del comp, comp_h, out_h
    # config 3: compress, 1 GiB of text per step
    nb3 = 256
    src = W.TextSource(seed=0x4C5A0003, device="cpu")
    plain = src.make(nb3 * BLOCK3).numpy()
    for _ in range(max(min(args.warmup, 2), 1)):
        cpu_compress_sample(plain, nb3, cores, reps=1)
    ksteps = max(1, min(args.steps, 5))
    t = time.perf_counter()
    for _ in range(ksteps):
        cv, clen, _cout, _st = cpu_compress_sample(plain, nb3, cores, reps=1)
    cdt = time.perf_counter() - t
    cvalue = ksteps * nb3 * BLOCK3 / GiB / cdt
    lz4c = liblz4_sample(True, plain, np.arange(nb3, dtype=np.uint64) * BLOCK3, np.full(nb3, BLOCK3, np.uint32), BLOCK3, cores, reps=2)
    line = {
        "impl": "reference", "metric": "LZ4 block decompress throughput (config 2: 64 KiB independent blocks, seq50)",
        "value": value, "unit": "GiB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "config2: decompress independent 64 KiB seq50 blocks (bounded sample of %d blocks = %d MiB per step)"
                               % (nb, nb * BLOCK2 >> 20)},
        "cpu_baseline": {"value": value, "unit": "GiB/s", "cores": cores, "kind": "port",
                         "sample": "%d config-2 blocks (%d MiB plaintext) per step, C port of the lz-fear decode loop "
                                   "(gcc -O3 -march=native, the reference's memset/memcpy/16-byte fast paths), one block per task"
                                   % (nb, nb * BLOCK2 >> 20),
                         "liblz4": lz4d},
        "e2e": {"value": value, "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "compress": {"metric": "LZ4 block compress throughput (config 3: 4 MiB text-like blocks, default CompressionSettings)",
                     "value": cvalue, "unit": "GiB/s", "ms_per_step": cdt / ksteps * 1e3, "steps": ksteps,
                     "ratio": float(nb3 * BLOCK3) / float(clen.astype(np.uint64).sum()),
                     "cpu_baseline": {"value": cvalue, "unit": "GiB/s", "cores": cores, "kind": "port",
                                      "sample": "%d config-3 blocks (%d MiB of text) per step, C port of lz-fear compress2 "
                                                "(gcc -O3 -march=native), one block per task" % (nb3, nb3 * BLOCK3 >> 20),
                                      "liblz4": lz4c},
                     "e2e": {"value": cvalue, "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from lz_fear_b200 import _native as N
    from lz_fear_b200 import workloads as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the codec has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        # a desynchronised collective should end the run in minutes, not after the default 10-minute watchdog
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks_ok(ok):
        """Collective AND: every rank calls it, every rank gets the same answer.  The exchange sections allocate large
        buffers on rank 0 only; a rank that failed locally must not leave the others waiting inside a collective it never
        joins (that is a 10-minute NCCL timeout), so every local step that can fail is followed by this agreement."""
        if world == 1:
            return bool(ok)
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    class EventTimer:
        """CUDA events on the current stream (the exchange runs there): sharding.frames_exchange's timer on the GPU."""

        def __init__(self):
            self.a, self.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def start(self):
            self.a.record()

        def stop_ms(self):
            self.b.record()
            self.b.synchronize()
            return self.a.elapsed_time(self.b)

    mem_log = {}

    def log_mem(tag):
        free, total = torch.cuda.mem_get_info()
        mem_log[tag] = {"torch_allocated_gib": round(torch.cuda.memory_allocated() / GiB, 2),
                        "torch_reserved_gib": round(torch.cuda.memory_reserved() / GiB, 2), "device_free_gib": round(free / GiB, 2)}

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ctx = N.Context(local_rank)
    peak_gbs, peak_src = measured_peak()

    def copy_ceiling(host_in, dev_in, host_out, dev_out, h2d_bytes, d2h_bytes, reps=3):
        """What the box's host<->device copy path allows for one e2e step: EVERY rank moves the step's H2D bytes and its
        D2H bytes at the same time (two streams, pinned memory, nothing else running), released together by a barrier;
        -> seconds per step (max over ranks).  e2e can not be faster than this, whatever the kernels do."""
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def once():
            with torch.cuda.stream(s1):
                dev_in[:h2d_bytes].copy_(host_in[:h2d_bytes], non_blocking=True)
            with torch.cuda.stream(s2):
                host_out[:d2h_bytes].copy_(dev_out[:d2h_bytes], non_blocking=True)
        once(); torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        barrier()
        return max_over_ranks(dt)
    K, Wm = args.steps, max(args.warmup, 3)
    stream = torch.cuda.current_stream().cuda_stream

    # =========================================================================================
    # config 2: decompress
    # =========================================================================================
    nb = max(BLOCKS_PER_FRAME2, int(args.decomp_gib * GiB) // BLOCK2 // BLOCKS_PER_FRAME2 * BLOCKS_PER_FRAME2)
    chunk = 4096
    comp = torch.empty(nb * BLOCK2, dtype=torch.uint8, device=dev)
    in_len = torch.empty(nb, dtype=torch.int32, device=dev)
    for b0 in range(0, nb, chunk):
        n = min(chunk, nb - b0)
        c, _o, l = W.seq50_blocks(n, seed=0x4C5A0002 + 7919 * (rank * 1000003 + b0), device=dev)
        comp[b0 * BLOCK2:(b0 + n) * BLOCK2] = c
        in_len[b0:b0 + n] = l
        del c, l
    in_off = torch.arange(nb, device=dev, dtype=torch.int64) * BLOCK2
    plain = torch.empty(nb * BLOCK2, dtype=torch.uint8, device=dev)
    cap = torch.full((nb,), BLOCK2, dtype=torch.int32, device=dev)
    olen = torch.zeros(nb, dtype=torch.int32, device=dev)
    st = torch.zeros(nb, dtype=torch.int32, device=dev)
    xx = torch.zeros(nb, dtype=torch.int32, device=dev)
    comp_bytes = int(in_len.sum().item())
    plain_bytes = nb * BLOCK2

    def step_decompress():
        ctx.decompress_blocks(comp, in_off, in_len, nb, plain, in_off, cap, cap, olen, st, None if args.no_xxh else xx, stream=stream)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(Wm):
        step_decompress()
    torch.cuda.synchronize()
    assert int(st.abs().sum().item()) == 0 and bool((olen == BLOCK2).all()), "decode failed"
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 0.6:      # keep the GPU under the same load while nvidia-smi gets going
        step_decompress()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step_decompress()
    e1.record()
    barrier()
    dec_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    dec_launches = ctx.launch_count - launches0
    total_plain = sum_over_ranks(plain_bytes)
    dec_value = total_plain * K / GiB / (dec_ms / 1e3)
    local_ms = e0.elapsed_time(e1) / K
    dec_roof = {"bound": "hbm", "achieved": (comp_bytes + plain_bytes) / 1e9 / (local_ms / 1e3), "peak": peak_gbs,
                "unit": "GB/s", "kernel": "decode_blocks_kernel", "peak_source": peak_src,
                "traffic": ncu_traffic("decode_blocks_kernel", nb)[0], "traffic_source": ncu_traffic("decode_blocks_kernel", nb)[1],
                "algorithmic_bytes_per_launch": comp_bytes + plain_bytes}
    dec_roof["frac"] = dec_roof["achieved"] / peak_gbs

    # ---- e2e: frames in pinned host memory through lzf_frames_decompress
    dec_e2e = None
    nframes = nb // BLOCKS_PER_FRAME2
    frame_plain = BLOCKS_PER_FRAME2 * BLOCK2
    if not args.no_e2e:
        # content checksum of every frame's plaintext, computed on the device from the decoded blocks
        f_off = torch.arange(nframes, device=dev, dtype=torch.int64) * frame_plain
        f_len = torch.full((nframes,), frame_plain, dtype=torch.int64, device=dev)
        f_hash = torch.zeros(nframes, dtype=torch.int32, device=dev)
        ctx.xxh32_ranges(plain, f_off, f_len, nframes, f_hash, stream=stream)
        torch.cuda.synchronize()
        comp_h = comp.cpu().numpy()
        len_h = in_len.cpu().numpy().astype(np.int64)
        hash_h = f_hash.cpu().numpy().view(np.uint32)
        hdr = np.frombuffer(bytes([0x04, 0x22, 0x4D, 0x18, 0x64, 0x40, 0xA7]), dtype=np.uint8)   # independent, content checksum, 64 KiB
        fr_len = np.array([7 + int(len_h[f * BLOCKS_PER_FRAME2:(f + 1) * BLOCKS_PER_FRAME2].sum()) + 4 * BLOCKS_PER_FRAME2 + 8
                           for f in range(nframes)], dtype=np.uint64)
        fr_off = np.zeros(nframes, dtype=np.uint64)
        fr_off[1:] = np.cumsum(fr_len)[:-1]
        frames_t = torch.empty(int(fr_len.sum()), dtype=torch.uint8).pin_memory()
        frames_h = frames_t.numpy()
        for f in range(nframes):
            p = int(fr_off[f])
            frames_h[p:p + 7] = hdr
            p += 7
            for b in range(f * BLOCKS_PER_FRAME2, (f + 1) * BLOCKS_PER_FRAME2):
                l = int(len_h[b])
                frames_h[p:p + 4] = np.frombuffer(int(l).to_bytes(4, "little"), dtype=np.uint8)
                frames_h[p + 4:p + 4 + l] = comp_h[b * BLOCK2:b * BLOCK2 + l]
                p += 4 + l
            frames_h[p:p + 4] = 0
            frames_h[p + 4:p + 8] = np.frombuffer(int(hash_h[f]).to_bytes(4, "little"), dtype=np.uint8)
        out_t = torch.empty(nb * BLOCK2, dtype=torch.uint8).pin_memory()
        out_h = out_t.numpy()
        o_off = np.arange(nframes, dtype=np.uint64) * frame_plain
        o_cap = np.full(nframes, frame_plain, dtype=np.uint64)

        def step_e2e():
            return ctx.frames_decompress(frames_h, fr_off, fr_len, out_h, o_off, o_cap)

        for _ in range(Wm):
            ol, fs, det = step_e2e()
        assert not fs.any() and (ol == frame_plain).all(), "frame decode failed: %s" % fs[fs != 0][:4]
        check = plain[:frame_plain * 4].cpu().numpy()
        assert np.array_equal(out_h[:frame_plain * 4], check)
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            step_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        dt = max_over_ranks(dt)
        dec_e2e = {"value": total_plain * K / GiB / dt, "unit": "GiB/s", "h2d_bytes_per_step": int(fr_len.sum()),
                   "d2h_bytes_per_step": int(nb * BLOCK2), "ms_per_step": dt / K * 1e3,
                   "api": "lzf_frames_decompress (host buffers, %d frames of %d blocks, content checksum verified)"
                          % (nframes, BLOCKS_PER_FRAME2)}
        # the copy ceiling of this box for exactly these bytes (all ranks copying at once, no kernels)
        h2d_tmp = torch.empty(int(fr_len.sum()), dtype=torch.uint8, device=dev)
        cdt = copy_ceiling(frames_t, h2d_tmp, out_t, plain, int(fr_len.sum()), int(nb * BLOCK2))
        del h2d_tmp
        dec_e2e["ceiling_gbs"] = total_plain / GiB / cdt
        dec_e2e["ceiling_note"] = ("plaintext GiB/s if the step were ONLY its pinned H2D + D2H copies, both directions at once on "
                                   "all %d ranks (measured in this run); frac_of_ceiling = value / ceiling_gbs" % world)
        dec_e2e["frac_of_ceiling"] = dec_e2e["value"] / dec_e2e["ceiling_gbs"]
        del frames_t, out_t

    # ---- CPU baseline (rank 0, bounded sample)
    cpu_dec = None
    if rank == 0 and not args.no_cpu:
        cores = os.cpu_count() or 1
        ns = min(nb, 16384)
        comp_s = comp[:ns * BLOCK2].cpu().numpy()
        off_s = np.arange(ns, dtype=np.uint64) * BLOCK2
        len_s = in_len[:ns].cpu().numpy().astype(np.uint32)
        v, ref = cpu_decompress_sample(comp_s, off_s, len_s, ns, cores)
        assert np.array_equal(ref, plain[:ns * BLOCK2].cpu().numpy()), "GPU decode differs from the oracle"
        cpu_dec = {"value": v, "unit": "GiB/s", "cores": cores, "kind": "port",
                   "sample": "first %d of the config-2 blocks (%d MiB plaintext), best of 3, C port of the lz-fear decode loop "
                             "(gcc -O3 -march=native, the reference's memset/memcpy/16-byte fast paths), one block per task over all "
                             "host threads; GPU output verified equal on this sample" % (ns, ns * BLOCK2 >> 20),
                   "liblz4": liblz4_sample(False, comp_s, off_s, len_s, BLOCK2, cores)}
        del ref
    del comp, plain
    torch.cuda.empty_cache()

    # =========================================================================================
    # config 3: compress
    # =========================================================================================
    comp_section = None
    if not args.no_compress:
        nb3 = max(BLOCKS_PER_FRAME3, int(args.comp_gib * GiB) // BLOCK3 // BLOCKS_PER_FRAME3 * BLOCKS_PER_FRAME3)
        src = W.TextSource(seed=0x4C5A0003 + rank, device=dev)
        data = torch.empty(nb3 * BLOCK3, dtype=torch.uint8, device=dev)
        for b0 in range(0, nb3, 16):
            n = min(16, nb3 - b0)
            data[b0 * BLOCK3:(b0 + n) * BLOCK3] = src.make(n * BLOCK3)
        off3 = torch.arange(nb3, device=dev, dtype=torch.int64) * BLOCK3
        len3 = torch.full((nb3,), BLOCK3, dtype=torch.int32, device=dev)
        cbuf = torch.empty(nb3 * BLOCK3, dtype=torch.uint8, device=dev)
        clen = torch.zeros(nb3, dtype=torch.int32, device=dev)
        cst = torch.zeros(nb3, dtype=torch.int32, device=dev)
        cxx = torch.zeros(nb3, dtype=torch.int32, device=dev)

        def step_compress():
            ctx.compress_blocks(data, off3, len3, nb3, cbuf, off3, None, clen, cst, cxx, None, stream=stream, max_block_len=BLOCK3)

        for _ in range(Wm):
            step_compress()
        torch.cuda.synchronize()
        assert int(cst.abs().sum().item()) == 0
        c_bytes = int(clen.to(torch.int64).sum().item())
        l0 = ctx.launch_count
        barrier()
        e0.record()
        for _ in range(K):
            step_compress()
        e1.record()
        barrier()
        c_ms = max_over_ranks(e0.elapsed_time(e1))
        c_launches = ctx.launch_count - l0
        total3 = sum_over_ranks(nb3 * BLOCK3)
        c_local_ms = e0.elapsed_time(e1) / K
        c_roof = {"bound": "hbm", "achieved": (nb3 * BLOCK3 + c_bytes) / 1e9 / (c_local_ms / 1e3), "peak": peak_gbs,
                  "unit": "GB/s", "kernel": "encode_blocks_kernel", "peak_source": peak_src,
                  "traffic": ncu_traffic("encode_blocks_kernel", nb3)[0], "traffic_source": ncu_traffic("encode_blocks_kernel", nb3)[1],
                  "algorithmic_bytes_per_launch": nb3 * BLOCK3 + c_bytes}
        c_roof["frac"] = c_roof["achieved"] / peak_gbs
        comp_section = {"metric": "LZ4 block compress throughput (config 3: 4 MiB text-like blocks, default CompressionSettings)",
                        "value": total3 * K / GiB / (c_ms / 1e3), "unit": "GiB/s", "ms_per_step": c_ms / K,
                        "ratio": nb3 * BLOCK3 / max(c_bytes, 1), "plaintext_gib_per_gpu": nb3 * BLOCK3 / GiB,
                        "roofline": c_roof, "gpu_launches": c_launches}
        # round trip property at full size: decode what we just wrote and compare on the device
        back = torch.empty_like(data)
        cap3 = len3.clone()
        olen3 = torch.zeros_like(clen)
        ctx.decompress_blocks(cbuf, off3, clen, nb3, back, off3, cap3, cap3, olen3, cst, cxx, stream=stream)
        torch.cuda.synchronize()
        assert int(cst.abs().sum().item()) == 0 and torch.equal(back, data), "compress -> decompress round trip failed"
        # the same decode, timed: a realistic token mix (text, 4 MiB blocks) next to the synthetic config 2
        e0.record()
        for _ in range(K):
            ctx.decompress_blocks(cbuf, off3, clen, nb3, back, off3, cap3, cap3, olen3, cst, cxx, stream=stream)
        e1.record()
        torch.cuda.synchronize()
        comp_section["roundtrip_decompress"] = {"value": nb3 * BLOCK3 * K / GiB / (e0.elapsed_time(e1) / 1e3), "unit": "GiB/s",
                                                "note": "decode of the blocks just written (text, 4 MiB blocks, XXH32 fused), this rank"}
        del back
        # ---- the same 16 GiB as whole FRAMES, device-resident: encode + layout + assembly + content checksums (and walk +
        # decode + checksum verification on the way back) inside the timed region — what §8(d) calls the kernel pipeline
        torch.cuda.empty_cache()
        try:
            nf3d = nb3 // BLOCKS_PER_FRAME3
            fpd = BLOCKS_PER_FRAME3 * BLOCK3
            sd, _kd = N.make_settings()
            bnd = ctx.frame_bound(sd, fpd)
            frd = torch.empty(nf3d * bnd, dtype=torch.uint8, device=dev)
            fi_o = np.arange(nf3d, dtype=np.uint64) * fpd
            fi_l = np.full(nf3d, fpd, np.uint64)
            fo_o = np.arange(nf3d, dtype=np.uint64) * bnd
            fo_c = np.full(nf3d, bnd, np.uint64)
            for _ in range(2):
                fld, fsd = ctx.frames_compress_device(data, fi_o, fi_l, frd, fo_o, fo_c, sd)
            assert not fsd.any()
            l0 = ctx.launch_count
            barrier()
            e0.record()
            for _ in range(K):
                ctx.frames_compress_device(data, fi_o, fi_l, frd, fo_o, fo_c, sd)
            e1.record()
            barrier()
            fc_ms = max_over_ranks(e0.elapsed_time(e1))
            fc_launches = ctx.launch_count - l0
            backd = torch.empty(nb3 * BLOCK3, dtype=torch.uint8, device=dev)
            for _ in range(2):
                old_, dsd, _dd = ctx.frames_decompress_device(frd, fo_o, fld, backd, fi_o, fi_l)
            assert not dsd.any() and torch.equal(backd, data), "frame round trip failed"
            barrier()
            e0.record()
            for _ in range(K):
                ctx.frames_decompress_device(frd, fo_o, fld, backd, fi_o, fi_l)
            e1.record()
            barrier()
            fd_ms = max_over_ranks(e0.elapsed_time(e1))
            comp_section["frames_device"] = {
                "compress_GiB_per_s": total3 * K / GiB / (fc_ms / 1e3), "decompress_GiB_per_s": total3 * K / GiB / (fd_ms / 1e3),
                "gpu_launches_compress": fc_launches,
                "note": "lzf_frames_compress_device / lzf_frames_decompress_device on %d frames of %d x 4 MiB (default settings): all "
                        "kernels of the direction (encode, layout, assembly, content checksums / walk, decode, checksums) inside "
                        "the CUDA-event region; round trip bit-exact on the device" % (nf3d, BLOCKS_PER_FRAME3)}
            del frd, backd
        except torch.OutOfMemoryError:
            comp_section["frames_device"] = {"skipped": "not enough device memory beside the block buffers"}
        torch.cuda.empty_cache()

        if not args.no_e2e:
            import psutil
            avail = psutil.virtual_memory().available
            e2e_blocks = nb3
            while e2e_blocks > BLOCKS_PER_FRAME3 and e2e_blocks * BLOCK3 * 2.2 > avail * 0.5 / world:   # every rank pins its own
                e2e_blocks //= 2
            e2e_blocks = e2e_blocks // BLOCKS_PER_FRAME3 * BLOCKS_PER_FRAME3
            nf3 = e2e_blocks // BLOCKS_PER_FRAME3
            fp = BLOCKS_PER_FRAME3 * BLOCK3
            s, _keep = N.make_settings()          # CompressionSettings::default()
            bound = ctx.frame_bound(s, fp)
            in_t = torch.empty(e2e_blocks * BLOCK3, dtype=torch.uint8).pin_memory()
            in_t.copy_(data[:e2e_blocks * BLOCK3])
            out_t = torch.empty(nf3 * bound, dtype=torch.uint8).pin_memory()
            i_off = np.arange(nf3, dtype=np.uint64) * fp
            i_len = np.full(nf3, fp, dtype=np.uint64)
            o_off = np.arange(nf3, dtype=np.uint64) * bound
            o_cap = np.full(nf3, bound, dtype=np.uint64)
            in_h, out_h = in_t.numpy(), out_t.numpy()

            def step_e2e3():
                return ctx.frames_compress(in_h, i_off, i_len, out_h, o_off, o_cap, s)

            for _ in range(Wm):
                fl, fs = step_e2e3()
            assert not fs.any()
            barrier()
            t0 = time.perf_counter()
            for _ in range(K):
                fl, fs = step_e2e3()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            barrier()
            dt = max_over_ranks(dt)
            tot = sum_over_ranks(e2e_blocks * BLOCK3)
            comp_section["e2e"] = {"value": tot * K / GiB / dt, "unit": "GiB/s", "h2d_bytes_per_step": int(e2e_blocks * BLOCK3),
                                   "d2h_bytes_per_step": int(fl.sum()), "ms_per_step": dt / K * 1e3,
                                   "api": "lzf_frames_compress (host buffers, %d frames of %d x 4 MiB blocks, default settings)"
                                          % (nf3, BLOCKS_PER_FRAME3)}
            # the frames must decode back to the input (checked through the host frame API on 2 frames)
            st_, det_, pl_, _c = ctx.frame_decompress(out_h[: int(fl[0])], cap=fp + 16)
            assert st_ == 0 and np.array_equal(np.frombuffer(pl_, dtype=np.uint8), in_h[:fp])
            h2d_tmp = torch.empty(e2e_blocks * BLOCK3, dtype=torch.uint8, device=dev)
            cdt = copy_ceiling(in_t, h2d_tmp, out_t, cbuf, int(e2e_blocks * BLOCK3), min(int(fl.sum()), cbuf.numel(), out_t.numel()))
            del h2d_tmp
            comp_section["e2e"]["ceiling_gbs"] = tot / GiB / cdt
            comp_section["e2e"]["frac_of_ceiling"] = comp_section["e2e"]["value"] / comp_section["e2e"]["ceiling_gbs"]
            comp_section["e2e"]["ceiling_note"] = "as for decompress: the step's pinned H2D + D2H bytes only, all %d ranks at once" % world
            del in_t, out_t
        if world > 1 and not args.no_gather:
            # config-4 flavour: the one real exchange step — compressed frames of every rank travel to rank 0 over NCCL
            # (all-gather of per-frame sizes, then ONE grouped batch of send/recv of the variable-length payloads).
            # Every local step that can fail is agreed on by all ranks before the next collective (all_ranks_ok).
            from lz_fear_b200 import sharding
            log_mem("before_gather")
            g_err, fr, packed, sizes = None, None, None, None
            try:
                gb = min(nb3, 1024)                                     # up to 4 GiB of plaintext per rank
                nf = gb // BLOCKS_PER_FRAME3
                fp = BLOCKS_PER_FRAME3 * BLOCK3
                sset, _k2 = N.make_settings()
                bound = ctx.frame_bound(sset, fp)
                fr = torch.empty(nf * bound, dtype=torch.uint8, device=dev)
                g_off = np.arange(nf, dtype=np.uint64) * bound
                fl, fs = ctx.frames_compress_device(data, np.arange(nf, dtype=np.uint64) * fp, np.full(nf, fp, np.uint64), fr, g_off,
                                                    np.full(nf, bound, np.uint64), sset)
                assert not fs.any()
                packed = torch.cat([fr[int(o):int(o) + int(l)] for o, l in zip(g_off, fl)])
                fr = None
                sizes = torch.from_numpy(fl.astype(np.int64)).to(dev)
            except Exception as e:
                g_err = "%s: %s" % (type(e).__name__, str(e)[:200])
            if not all_ranks_ok(g_err is None):
                comp_section["gather"] = {"skipped": "a rank could not prepare its frames", "this_rank": g_err}
            else:
                res = sharding.frames_exchange(packed, sizes, dev, reps=K, timer=EventTimer(), barrier=barrier)
                if "skipped" in res:
                    comp_section["gather"] = res
                else:
                    g_ms, s_ms = max_over_ranks(res["gather_ms"]), max_over_ranks(res["scatter_ms"])
                    total_bytes = sum(res["totals"])
                    remote = total_bytes - res["totals"][0]
                    comp_section["gather"] = {"ms": g_ms, "scatter_ms": s_ms, "compressed_bytes_all_ranks": total_bytes,
                                              "GiB_per_s_into_rank0": remote / GiB / (g_ms / 1e3),
                                              "scatter_GiB_per_s_out_of_rank0": remote / GiB / (s_ms / 1e3),
                                              "plaintext_GiB_per_rank": nf * fp / GiB,
                                              "verified": {"archive_slices_equal_senders_digests": res["archive_slices_equal_senders_digests"],
                                                           "scattered_frames_equal_on_every_rank": res["scattered_frames_equal_on_every_rank_and_decoded"]},
                                              "note": "sharding.frames_exchange: NCCL all_gather(sizes) + one grouped batch of send/recv "
                                                      "(ncclGroupStart/End) of whole frames per direction; not part of `value`"}
            del fr, packed, sizes
            torch.cuda.empty_cache()
        if rank == 0 and not args.no_cpu:
            cores = os.cpu_count() or 1
            ns = min(nb3, 256)
            h = data[:ns * BLOCK3].cpu().numpy()
            v, rl, rout, rst = cpu_compress_sample(h, ns, cores)
            g = cbuf[:ns * BLOCK3].cpu().numpy()
            gl = clen[:ns].cpu().numpy().view(np.uint32)
            assert not rst.any() and np.array_equal(gl, rl), "compressed sizes differ from the oracle"
            for b in range(ns):
                assert np.array_equal(g[b * BLOCK3:b * BLOCK3 + rl[b]], rout[b * BLOCK3:b * BLOCK3 + rl[b]])
            comp_section["cpu_baseline"] = {"value": v, "unit": "GiB/s", "cores": cores, "kind": "port",
                                            "sample": "first %d config-3 blocks (%d MiB), best of 2, C port of lz-fear compress2 (gcc -O3 "
                                                      "-march=native), one block per task; GPU output byte-identical on this sample" % (ns, ns * BLOCK3 >> 20),
                                            "liblz4": liblz4_sample(True, h, np.arange(ns, dtype=np.uint64) * BLOCK3,
                                                                    np.full(ns, BLOCK3, np.uint32), BLOCK3, cores, reps=2)}
            del h, rout, g

    # =========================================================================================
    # config 4 (mixed-entropy frames, this rank's shard) and config 5 (large hash tables) — opt-in
    # =========================================================================================
    extra = None
    if not args.no_compress:
        del data, cbuf, off3, len3, clen, cst, cxx
    ctx.trim()                              # the e2e pipelines left tens of GiB of staging scratch in the context
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    log_mem("before_extra_configs")
    if not args.no_extra:
        extra = {}
        torch.cuda.empty_cache()
        # ---- config 4: block b's class = b mod 3 -> random (stored-block fallback) / text / lowent; frames of 16 x 4 MiB
        nb4 = max(BLOCKS_PER_FRAME3, int(args.mixed_gib * GiB) // BLOCK3 // BLOCKS_PER_FRAME3 * BLOCKS_PER_FRAME3)
        mixed = W.mixed_blocks(nb4, BLOCK3, seed=0x4C5A0004 + 1000003 * rank, device=dev)
        nf4 = nb4 // BLOCKS_PER_FRAME3
        fp4 = BLOCKS_PER_FRAME3 * BLOCK3
        s4, _k4 = N.make_settings()
        bound4 = ctx.frame_bound(s4, fp4)
        fr4 = torch.empty(nf4 * bound4, dtype=torch.uint8, device=dev)
        fi_off = np.arange(nf4, dtype=np.uint64) * fp4
        fi_len = np.full(nf4, fp4, np.uint64)
        fo_off = np.arange(nf4, dtype=np.uint64) * bound4
        fo_cap = np.full(nf4, bound4, np.uint64)
        back4 = torch.empty_like(mixed)

        def c4_compress():
            return ctx.frames_compress_device(mixed, fi_off, fi_len, fr4, fo_off, fo_cap, s4)

        def c4_decompress(fl):
            return ctx.frames_decompress_device(fr4, fo_off, fl, back4, fi_off, fi_len)

        for _ in range(2):
            fl4, fs4 = c4_compress()
            ol4, ds4, _d = c4_decompress(fl4)
        assert not fs4.any() and not ds4.any() and (ol4 == fp4).all() and torch.equal(back4, mixed), "config 4 round trip failed"
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            c4_compress()                    # synchronous: returns once the frames are assembled
        torch.cuda.synchronize()
        tc = time.perf_counter() - t0
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            c4_decompress(fl4)
        torch.cuda.synchronize()
        td = time.perf_counter() - t0
        barrier()
        tc, td = max_over_ranks(tc), max_over_ranks(td)
        tot4 = sum_over_ranks(nb4 * BLOCK3)
        extra["config4"] = {
            "workload": "mixed-entropy frames (block class = b mod 3: random / text / lowent), %d frames of 16 x 4 MiB per GPU, "
                        "default CompressionSettings, device-resident, whole frame calls (walk/encode/layout/assemble, content checksums)" % nf4,
            "compress_GiB_per_s": tot4 * K / GiB / tc, "decompress_GiB_per_s": tot4 * K / GiB / td,
            "ratio": float(nb4 * BLOCK3) / float(fl4.sum()), "n_gpus": world,
            "timing": "host clock around K synchronous frame calls, max over ranks", "roundtrip": "bit-exact on the device"}
        if world > 1 and not args.no_gather:
            # §8(e), the full exchange: every rank compresses its frames -> all-gather of sizes + grouped send/recv GATHER of
            # whole frames to rank 0 (the archive) -> rank 0 walks the frame boundaries and SCATTERS contiguous ranges of whole
            # frames back -> every rank decompresses what it received -> bit-exact against its own plaintext.
            # Collective-safe: all ranks agree (all_ranks_ok) after every local step that can fail, and nothing raises
            # between two collectives.
            from lz_fear_b200 import sharding
            x_err, packed4, sizes4 = None, None, None
            try:
                packed4 = torch.cat([fr4[int(o):int(o) + int(l)] for o, l in zip(fo_off, fl4)])
                sizes4 = torch.from_numpy(fl4.astype(np.int64)).to(dev)
            except Exception as e:
                x_err = "%s: %s" % (type(e).__name__, str(e)[:200])
            del fr4                                               # the frames live on in packed4
            torch.cuda.empty_cache()
            log_mem("before_config4_exchange")
            if not all_ranks_ok(x_err is None):
                extra["config4"]["exchange"] = {"skipped": "a rank could not pack its frames", "this_rank": x_err}
            else:
                def decode_back(received):
                    # decode what came back over the wire (dense layout: frame f at the running sum of the frame lengths)
                    r_off = np.zeros(nf4, dtype=np.uint64); r_off[1:] = np.cumsum(fl4)[:-1]
                    back4.zero_()
                    ol4b, ds4b, _d = ctx.frames_decompress_device(received, r_off, fl4, back4, fi_off, fi_len)
                    return bool(not ds4b.any() and (ol4b == fp4).all() and torch.equal(back4, mixed))
                res = sharding.frames_exchange(packed4, sizes4, dev, reps=K, decode=decode_back, timer=EventTimer(), barrier=barrier)
                if "skipped" in res:
                    res["mem"] = mem_log.get("before_config4_exchange")
                    extra["config4"]["exchange"] = res
                else:
                    g4_ms, s4_ms = max_over_ranks(res["gather_ms"]), max_over_ranks(res["scatter_ms"])
                    totals4 = res["totals"]
                    remote = sum(totals4) - totals4[0]
                    comp_s4, dec_s4 = tc / K, td / K
                    extra["config4"]["exchange"] = {
                        "gather_ms": g4_ms, "scatter_ms": s4_ms, "frame_bytes_all_ranks": sum(totals4),
                        "gather_GiB_per_s_into_rank0": remote / GiB / (g4_ms / 1e3), "scatter_GiB_per_s_out_of_rank0": remote / GiB / (s4_ms / 1e3),
                        "compress_GiB_per_s_with_gather": tot4 / GiB / (comp_s4 + g4_ms / 1e3),
                        "decompress_GiB_per_s_with_scatter": tot4 / GiB / (dec_s4 + s4_ms / 1e3),
                        "verified": {"archive_slices_equal_senders_digests": res["archive_slices_equal_senders_digests"],
                                     "scattered_frames_equal_on_every_rank_and_decode_bit_exact": res["scattered_frames_equal_on_every_rank_and_decoded"],
                                     "this_rank_error": res["this_rank_error"]},
                        "how": "sharding.frames_exchange: NCCL all_gather(sizes) + one grouped batch of isend/irecv per direction "
                               "(batch_isend_irecv), whole frames only"}
            del packed4, sizes4
            torch.cuda.empty_cache()
        if rank == 0 and not args.no_cpu:
            # CPU bar for config 4: the same block mix through the C port (blocks of the first frames; stored blocks included)
            cores = os.cpu_count() or 1
            ns4 = min(nb4, 96)
            h4 = mixed[:ns4 * BLOCK3].cpu().numpy()
            cv4, cl4, cout4, cst4 = cpu_compress_sample(h4, ns4, cores, reps=2)
            # decode what compressed; stored (refused) blocks are a memcpy in the frame reader and are skipped here
            ok4 = np.nonzero(cst4 == 0)[0]
            import oracle
            d_out = np.empty(len(ok4) * BLOCK3, dtype=np.uint8)
            d_off = np.arange(len(ok4), dtype=np.uint64) * BLOCK3
            d_cap = np.full(len(ok4), BLOCK3, dtype=np.uint32)
            t0 = time.perf_counter()
            dl4, ds4_ = oracle.decompress_blocks_mt(cout4, ok4.astype(np.uint64) * BLOCK3, cl4[ok4].astype(np.uint32), d_out, d_off, d_cap, d_cap,
                                                    nthreads=cores, native=True)
            td4 = time.perf_counter() - t0
            assert not ds4_.any()
            extra["config4"]["cpu_baseline"] = {"compress_GiB_per_s": cv4, "decompress_GiB_per_s": len(ok4) * BLOCK3 / GiB / td4, "cores": cores,
                                                "kind": "port", "sample": "first %d blocks (%d MiB): block loops only, no frame assembly or checksums; "
                                                "decompress over the %d blocks that compressed" % (ns4, ns4 * BLOCK3 >> 20, len(ok4))}
            del h4, cout4, d_out
        fr4 = None
        del mixed, fr4, back4
        torch.cuda.empty_cache()
        # ---- config 5: low-entropy blocks, HASHLOG 12 (reference) / 14 / 16 (extension): sizes vs the oracle at the same HASHLOG
        if rank == 0:
            import oracle
            nb5 = max(1, int(args.lowent_gib * GiB) // BLOCK3)
            low = torch.cat([W.lowent(BLOCK3, 0x4C5A0005 + b, device=dev) for b in range(nb5)])
            off5 = torch.arange(nb5, device=dev, dtype=torch.int64) * BLOCK3
            len5 = torch.full((nb5,), BLOCK3, dtype=torch.int32, device=dev)
            c5 = torch.empty(nb5 * BLOCK3, dtype=torch.uint8, device=dev)
            cl5 = torch.zeros(nb5, dtype=torch.int32, device=dev)
            cs5 = torch.zeros(nb5, dtype=torch.int32, device=dev)
            back5 = torch.empty_like(low)
            res5 = {}
            ns5 = min(nb5, 64)
            sample = low[:ns5 * BLOCK3].cpu().numpy()
            cores = os.cpu_count() or 1
            for hl in (12, 14, 16):
                def run5():
                    ctx.compress_blocks(low, off5, len5, nb5, c5, off5, None, cl5, cs5, None, None, hashlog=hl, stream=stream, max_block_len=BLOCK3)
                for _ in range(2):
                    run5()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(K):
                    run5()
                e1.record()
                torch.cuda.synchronize()
                assert int(cs5.abs().sum().item()) == 0
                ctx.decompress_blocks(c5, off5, cl5, nb5, back5, off5, len5, len5, torch.zeros_like(cl5), cs5, None, stream=stream)
                torch.cuda.synchronize()
                assert int(cs5.abs().sum().item()) == 0 and torch.equal(back5, low), "config 5 round trip failed"
                got = cl5[:ns5].cpu().numpy().view(np.uint32)
                res5["hashlog%d" % hl] = {"compress_GiB_per_s": nb5 * BLOCK3 * K / GiB / (e0.elapsed_time(e1) / 1e3),
                                          "ratio": float(nb5 * BLOCK3) / float(cl5.to(torch.int64).sum().item())}
                if not args.no_cpu:
                    o5 = np.arange(ns5, dtype=np.uint64) * BLOCK3
                    l5 = np.full(ns5, BLOCK3, dtype=np.uint32)
                    out5 = np.empty(ns5 * BLOCK3, dtype=np.uint8)
                    t0 = time.perf_counter()
                    want, wst = oracle.compress_blocks_mt(sample, o5, l5, out5, o5, hashlog=hl, nthreads=cores, native=True)
                    t5 = time.perf_counter() - t0
                    gbytes = c5[:ns5 * BLOCK3].cpu().numpy()
                    same = all(np.array_equal(gbytes[b * BLOCK3:b * BLOCK3 + int(want[b])], out5[b * BLOCK3:b * BLOCK3 + int(want[b])]) for b in range(ns5))
                    res5["hashlog%d" % hl].update({
                        "size_vs_oracle_same_hashlog": {"blocks": ns5, "max_abs_diff_bytes": int(np.abs(got.astype(np.int64) - want.astype(np.int64)).max()),
                                                        "bytes_identical": bool(same)},
                        "cpu_baseline": {"value": ns5 * BLOCK3 / GiB / t5, "unit": "GiB/s", "cores": cores, "kind": "port",
                                         "sample": "first %d blocks, one run, C port of compress2 with the same HASHLOG" % ns5}})
                    del out5, gbytes
            # the same 256 blocks with the segmented parse (LZF_OPT_SEGMENT_BYTES): 256 blocks fill 6 % of the warp slots, so every
            # block is cut into segments parsed side by side and stitched into one LZ4 block — valid LZ4 of the reference's
            # size (north_star: within 1 %), not its bytes
            exact_total = None
            try:
                run_exact = lambda: ctx.compress_blocks(low, off5, len5, nb5, c5, off5, None, cl5, cs5, None, None, hashlog=12, stream=stream, max_block_len=BLOCK3)
                run_exact(); torch.cuda.synchronize()
                exact_total = int(cl5.to(torch.int64).sum().item())
                ctx.set_option(N.OPT_SEGMENT_BYTES, 65536)
                for _ in range(2):
                    run_exact()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(K):
                    run_exact()
                e1.record()
                torch.cuda.synchronize()
                seg_ms = e0.elapsed_time(e1)
                assert int(cs5.abs().sum().item()) == 0
                seg_total = int(cl5.to(torch.int64).sum().item())
                ctx.decompress_blocks(c5, off5, cl5, nb5, back5, off5, len5, len5, torch.zeros_like(cl5), cs5, None, stream=stream)
                torch.cuda.synchronize()
                assert int(cs5.abs().sum().item()) == 0 and torch.equal(back5, low), "segmented parse round trip failed"
                res5["hashlog12_segmented"] = {"compress_GiB_per_s": nb5 * BLOCK3 * K / GiB / (seg_ms / 1e3),
                                               "ratio": float(nb5 * BLOCK3) / float(seg_total),
                                               "size_vs_exact_parse": seg_total / float(exact_total) - 1.0,
                                               "note": "LZF_OPT_SEGMENT_BYTES=65536: segments parsed side by side + stitch; round trip bit-exact on the "
                                                       "device; size relative to the byte-identical parse above"}
            except Exception as e:
                res5["hashlog12_segmented"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
            finally:
                ctx.set_option(N.OPT_SEGMENT_BYTES, 0)
            extra["config5"] = {"workload": "%d x 4 MiB low-entropy blocks (4-symbol alphabet, runs U{1..64}); HASHLOG 12 is the reference's table, "
                                            "14 / 16 are the large-table extension (parity against the oracle run with the same HASHLOG)" % nb5,
                                "results": res5}
            del low, c5, back5

    # =========================================================================================
    # the drop-in case: ONE file through CompressionSettings::compress / decompress_frame (host buffers, 64 MiB of text)
    # =========================================================================================
    single = None
    if rank == 0 and not args.no_extra and not args.no_e2e:
        try:
            sf = W.TextSource(seed=0x4C5A0006, device="cpu").make(64 << 20).numpy()
            sset, _ks = N.make_settings()
            cap_sf = ctx.frame_bound(sset, sf.size)

            def timed(fn, reps=3):
                fn()
                t0 = time.perf_counter()
                for _ in range(reps):
                    r = fn()
                return (time.perf_counter() - t0) / reps, r
            t_exact, (st_e, fr_e) = timed(lambda: ctx.frame_compress(sf))
            ctx.set_option(N.OPT_SEGMENT_BYTES, 65536)
            try:
                t_seg, (st_s, fr_s) = timed(lambda: ctx.frame_compress(sf))
            finally:
                ctx.set_option(N.OPT_SEGMENT_BYTES, 0)
            assert st_e == 0 and st_s == 0
            t_dec, (dst_, ddet_, dplain_, _dc) = timed(lambda: ctx.frame_decompress(fr_s, cap=sf.size + 16))
            assert dst_ == 0 and np.array_equal(np.frombuffer(dplain_, dtype=np.uint8), sf), "single-file round trip failed"
            # the streaming host mirror on the same bytes (examples/delz4.rs: LZ4FrameReader::new(f)?.into_read(), fill_buf /
            # consume): read-ahead batches of whole blocks, one launch each, decoded while the caller consumes the previous one
            import io
            import lz_fear_b200 as L
            L.raw.set_default_context(ctx)
            try:
                def stream_read():
                    rd = L.LZ4FrameReader(io.BytesIO(fr_e)).into_read()
                    n = 0
                    while True:
                        buf = rd.fill_buf()
                        if not buf:
                            return n
                        n += len(buf)
                        rd.consume(len(buf))
                t_stream, n_stream = timed(stream_read)
                assert n_stream == sf.size
                t_oneshot, (ost_, odet_, opl_, _oc) = timed(lambda: ctx.frame_decompress(fr_e, cap=sf.size + 16))
                assert ost_ == 0 and len(opl_) == sf.size
            finally:
                L.raw.set_default_context(None)
            import oracle
            t0 = time.perf_counter()
            orc, ofr = oracle.frame_compress(sf.tobytes())
            t_cpu = time.perf_counter() - t0
            assert (orc, ofr) == (0, fr_e), "single-file frame differs from the oracle"
            single = {"workload": "one 64 MiB text file, default CompressionSettings (16 blocks of 4 MiB), host buffers in and out",
                      "compress_exact_GiB_per_s": sf.size / GiB / t_exact, "compress_segmented_GiB_per_s": sf.size / GiB / t_seg,
                      "segmented_size_vs_exact": len(fr_s) / float(len(fr_e)) - 1.0,
                      "decompress_GiB_per_s": sf.size / GiB / t_dec,
                      "stream_reader_GiB_per_s": sf.size / GiB / t_stream, "stream_reader_vs_one_shot": t_stream / t_oneshot,
                      "cpu_baseline": {"compress_GiB_per_s": sf.size / GiB / t_cpu, "cores": 1, "kind": "port",
                                       "sample": "the same file through the C port of compress_internal, one thread (the reference is single-threaded)"},
                      "note": "exact = byte-identical to the reference (16 warps busy); segmented = LZF_OPT_SEGMENT_BYTES=65536"}
        except Exception as e:
            single = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

    if rank == 0:
        line = {
            "metric": "LZ4 block decompress throughput (config 2: 64 KiB independent blocks, seq50)",
            "value": dec_value, "unit": "GiB/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": dec_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "config2: decompress %d independent 64 KiB seq50 blocks per GPU (%.2f GiB plaintext, %.2f GiB compressed)"
                                   % (nb, plain_bytes / GiB, comp_bytes / GiB),
                       "l2": "inputs (%.1f GiB read + %.1f GiB written per step) exceed the 126 MB L2; no flush needed"
                             % (comp_bytes / GiB, plain_bytes / GiB),
                       "sharding": "independent blocks split evenly over ranks, no data-path collective"},
            "roofline": dec_roof, "cpu_baseline": cpu_dec, "e2e": dec_e2e, "gpu_launches": dec_launches,
            "clocks": clocks, "compress": comp_section,
        }
        if extra is not None:
            line["extra_configs"] = extra
        if single is not None:
            line["single_file"] = single
        line["device_memory_log"] = mem_log
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
