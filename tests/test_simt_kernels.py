"""Runs the REAL kernel + C-ABI sources (rust-lz-fear_b200/csrc) on the CPU SIMT emulator
(tests/simt) and checks them bit-for-bit against the oracle.  Same checks as the `-m gpu` tier
(tests/test_gpu_parity.py) at sizes the emulator finishes in seconds."""
import numpy as np
import pytest

import parity
from lz_fear_b200 import _native as N


def _strings(vectors):
    return [s.encode("latin-1") for v in vectors["roundtrip_strings"].values() for s in v]


def test_raw_compress_matches_oracle(emu, oracle, vectors):
    parity.check_raw_compress(emu, oracle, parity.sample_inputs() + _strings(vectors))


def test_raw_compress_u16_table(emu, oracle, vectors):          # src/lib.rs:26-27 helper path
    inputs = [b for b in parity.sample_inputs() + _strings(vectors) if len(b) <= 0xFFFF]
    parity.check_raw_compress(emu, oracle, inputs, table=N.TABLE_U16)
    st, _ = emu.ctx.raw_compress_into(bytes(70000), table=N.TABLE_U16)
    assert st == N.PANIC                                        # assert at src/raw/compress/mod.rs:167


def test_raw_compress_large_hashlog(emu, oracle):               # BASELINE config 5 extension
    inputs = [b for b in parity.sample_inputs() if 1000 <= len(b) <= 100000][:6]
    for hashlog in (13, 14, 16):
        parity.check_raw_compress(emu, oracle, inputs, hashlog=hashlog, caps=False)


def test_decode_kats(emu, oracle, vectors):                     # src/raw/decompress.rs:153-175
    blocks = [(bytes(k["input"]), None) for k in vectors["decode_kats"]]
    parity.check_raw_decompress(emu, oracle, blocks)
    st, out, n = emu.ctx.raw_decompress(bytes([0x11, 97, 1, 0, 0x22, 98, 99, 2, 0]))
    assert (st, out) == (0, b"aaaaaabcbcbcbc")


def test_raw_decompress_matches_oracle(emu, oracle):
    blocks = []
    for data in parity.sample_inputs():
        st, comp = oracle.compress_block(data)
        blocks.append((comp, len(data)))
    parity.check_raw_decompress(emu, oracle, blocks)


def test_raw_decompress_seq50(emu, oracle):
    from lz_fear_b200 import workloads as W
    comp, off, ln = W.seq50_blocks(4, seed=11)
    c = comp.numpy()
    blocks = [(c[int(o): int(o) + int(l)].tobytes(), 65536) for o, l in zip(off, ln)]
    parity.check_raw_decompress(emu, oracle, blocks)


def test_raw_decompress_malformed(emu, oracle):
    """error precedence of SURVEY §8 row D1 on mutated blocks"""
    base = [oracle.compress_block(d)[1] for d in parity.sample_inputs() if 20 <= len(d) <= 70000][:12]
    blocks = []
    for i, comp in enumerate(base):
        for k in range(6):
            blocks.append((parity.mutate(comp, 100 * i + k, k=1 + k % 3), None))
    blocks += [(bytes([0x0F]), None), (bytes([0xF0]), None), (bytes([0xF0, 0xFF]), None), (bytes([0x00, 0x00, 0x00]), None),
               (bytes([0x10, 65, 0x01]), None), (bytes([0x1F, 65, 1, 0, 0xFF]), None), (bytes([0x00, 1, 0]), None)]
    parity.check_raw_decompress(emu, oracle, blocks)


def test_raw_decompress_with_prefix(emu, oracle):
    prefix = b"0123456789abcdef" * 8
    comp = bytes([0x42, 120, 121, 122, 119, 20, 0, 0x00, 3, 0])      # lits "xyzw", match 6 @ off 20 -> reaches prefix
    for lim in (6, 10, 1 << 20):
        st, out, n = emu.ctx.raw_decompress(comp, prefix=prefix, out_limit=lim, cap=256)
        assert (st, out, n) == oracle.decompress_raw(comp, prefix=prefix, out_limit=lim, cap=256)
    st, out, n = emu.ctx.raw_decompress(comp, prefix=prefix[:10], out_limit=1 << 20, cap=256)
    assert st == N.INVALID_DEDUP_OFFSET


def test_batched_blocks(emu, oracle):
    inputs = [b for b in parity.sample_inputs() if len(b) > 0]
    parity.check_batched_blocks(emu, oracle, inputs)


def test_batched_blocks_tables_in_global_scratch(emu, oracle, monkeypatch):
    """max_block_len promise -> u16 / packed 17-bit slot tables in one big CTA per SM; with fewer shared-memory
    tables than warps the remaining warps keep theirs in the global scratch (same bytes either way)."""
    inputs = [b for b in parity.sample_inputs() if len(b) > 0]
    small = [b for b in inputs if len(b) <= 65536]
    for smem_warps in ("1", "0", None):
        if smem_warps is None:
            monkeypatch.delenv("LZF_B200_ENC_SMEM_WARPS", raising=False)
        else:
            monkeypatch.setenv("LZF_B200_ENC_SMEM_WARPS", smem_warps)
        with emu.fresh() as b:
            parity.check_batched_blocks(b, oracle, inputs, max_block_len=max(len(x) for x in inputs))     # packed slots
            parity.check_batched_blocks(b, oracle, small, max_block_len=65536)                            # u16 slots


def test_frame_compress_with_sliced_input_feed(emu, oracle, monkeypatch, scale=1):
    """Host-buffer compress feeds large independent blocks slice by slice while the kernel runs (the warps poll a
    progress word); same frames as the oracle, stored (incompressible) blocks included."""
    from lz_fear_b200 import workloads as W
    monkeypatch.setenv("LZF_B200_FEED_MIN_BLOCKS", "2")
    monkeypatch.setenv("LZF_B200_FEED_SLICE", "65536")
    bs = 256 << 10
    data = (W.text(bs + bs // 2, 21).numpy().tobytes() + W.random_bytes(bs, 22).numpy().tobytes() +
            W.lowent(bs // 2, 23).numpy().tobytes() + bytes(bs)) * scale
    assert len(data) % bs == 0
    with emu.fresh() as b:
        for kw in (dict(block_size=bs), dict(block_size=bs, block_checksums=True, content_checksum=False)):
            st, frame = b.ctx.frame_compress(data, **kw)
            assert (st, frame) == oracle.frame_compress(data, **kw)
        assert b.ctx.frame_decompress(frame, cap=len(data) + 16)[:3] == (0, 0, data)


def test_batched_host_frames_with_a_refused_frame(emu, oracle):
    """lzf_frames_compress assembles the frames before their content checksums are known and patches the 4-byte
    trailers afterwards: a frame whose capacity is too small must stay a WriteError and must not be touched."""
    from lz_fear_b200 import workloads as W
    nf, fp = 4, 96 << 10
    src = np.concatenate([W.text(fp, 31 + f).numpy() for f in range(nf)])
    s, keep = N.make_settings(block_size=64 << 10)
    bound = emu.ctx.frame_bound(s, fp)
    out = np.full(nf * bound, 0xEE, dtype=np.uint8)
    caps = np.full(nf, bound, np.uint64)
    caps[2] = 100                                                   # far too small for frame 2
    fl, fs = emu.ctx.frames_compress(src, np.arange(nf, dtype=np.uint64) * fp, np.full(nf, fp, np.uint64), out,
                                     np.arange(nf, dtype=np.uint64) * bound, caps, s)
    for f in range(nf):
        if f == 2:
            assert fs[f] == N.F_WRITE_ERROR == oracle.F_WRITE_ERROR and (out[f * bound + 100:(f + 1) * bound] == 0xEE).all()
        else:
            want = oracle.frame_compress(src[f * fp:(f + 1) * fp].tobytes(), block_size=64 << 10)
            assert (int(fs[f]), out[f * bound:f * bound + int(fl[f])].tobytes()) == want, f


def test_frames_roundtrip_and_bytes(emu, oracle):
    inputs = [b"", b"a", bytes(65536), parity.sample_inputs()[6], parity.sample_inputs()[7][:70001],
              parity.sample_inputs()[5] * 30]
    parity.check_frames(emu, oracle, inputs)
    z = bytes(65536)                                               # BASELINE config 1 KAT
    st, frame = emu.ctx.frame_compress(z)
    assert st == 0 and len(frame) == 286 and frame[-4:] == bytes([0x1C, 0xE8, 0x64, 0x0F])


def test_frame_settings_errors(emu, oracle):
    data = b"x" * 1000
    assert emu.ctx.frame_compress(data, block_size=12345)[0] == N.F_INVALID_BLOCK_SIZE
    assert emu.ctx.frame_compress(data, block_size=16 << 20)[0] == N.F_PANIC
    st, frame = emu.ctx.frame_compress(data, cap=10)
    assert st == N.F_WRITE_ERROR == oracle.F_WRITE_ERROR
    assert emu.ctx.frame_compress(data, independent_blocks=False) == oracle.frame_compress(data, independent_blocks=False)


def test_frame_decode_corpus(emu, oracle, corpora):              # fuzz/corpus/decode replay
    blobs = [b for _, b in corpora["decode"]]
    n_ok = parity.check_frame_decode_errors(emu, oracle, blobs[::3])
    assert n_ok >= 0


def test_frame_decode_mutations(emu, oracle):
    data = parity.sample_inputs()[6] + parity.sample_inputs()[5]
    frames = []
    for kw in parity.FRAME_SETTINGS[1:5]:
        rc, frame = oracle.frame_compress(data, **kw)
        for k in range(12):
            frames.append(parity.mutate(frame, hash(str(kw)) % 1000 + k, k=1 + k % 2))
        frames += [frame[:n] for n in (0, 3, 6, 7, 10, 11, len(frame) - 5, len(frame) - 1)]
    parity.check_frame_decode_errors(emu, oracle, frames)


def test_short_and_empty_blocks_in_frame(emu, oracle):
    import struct
    hdr = bytes([0x04, 0x22, 0x4D, 0x18, 0x60, 0x40, 0x82])
    a = bytes([0x30]) + b"abc"
    big = oracle.compress_block(bytes(65536))[1]
    frames = [
        hdr + struct.pack("<I", 4) + a + struct.pack("<I", 1) + b"\x00" + struct.pack("<I", 4) + a + struct.pack("<I", 0),
        hdr + struct.pack("<I", 4) + a + struct.pack("<I", len(big)) + big + struct.pack("<I", 4) + a + struct.pack("<I", 0),
        hdr + struct.pack("<I", 3 | 0x80000000) + b"xyz" + struct.pack("<I", 4) + a + struct.pack("<I", 0),
        hdr + struct.pack("<I", 70000) + bytes(70000),
    ]
    parity.check_frame_decode_errors(emu, oracle, frames)


def test_streaming_xxh32(emu, oracle):
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256, 100000, dtype=np.uint8).tobytes()
    st = emu.ctx.xxh32_new()
    pos = 0
    for step in [0, 1, 3, 15, 16, 17, 31, 32, 1000, 4096, 5, 50000]:
        emu.ctx.xxh32_update(st, data[pos: pos + step])
        pos += step
        assert emu.ctx.xxh32_finish(st) == oracle.xxh32(data[:pos])


def test_batched_frames_pipeline_and_hash_queue(simt_lib_path, oracle, monkeypatch):
    """Many frames through the host-buffer batch calls with tiny pipeline chunks (several chunks per
    call, all three slots in rotation) on a 1-SM device (one CTA: the XXH32 queue wraps its ring)."""
    from lz_fear_b200 import _native
    from lz_fear_b200 import workloads as W
    monkeypatch.setenv("LZF_B200_CHUNK_BYTES", "20000")
    monkeypatch.setenv("SIMT_NUM_SMS", "1")
    saved = (_native._lib, _native._lib_path)
    _native.load_library(simt_lib_path)
    ctx = _native.Context(0)
    try:
        rng = np.random.default_rng(3)
        datas = []
        for i in range(90):
            n = int(rng.integers(0, 9000))
            kind = i % 3
            datas.append(W.text(n, i).numpy().tobytes() if kind == 0 else
                         W.lowent(n, i).numpy().tobytes() if kind == 1 else W.random_bytes(n, i).numpy().tobytes())
        s, _k = _native.make_settings(block_size=64 << 10, block_checksums=True)
        in_len = np.array([len(d) for d in datas], dtype=np.uint64)
        in_off = np.zeros(len(datas), dtype=np.uint64); in_off[1:] = np.cumsum(in_len)[:-1]
        inp = np.frombuffer(b"".join(datas), dtype=np.uint8)
        caps = np.array([ctx.frame_bound(s, int(n)) for n in in_len], dtype=np.uint64)
        out_off = np.zeros(len(datas), dtype=np.uint64); out_off[1:] = np.cumsum(caps)[:-1]
        out = np.zeros(int(caps.sum()), dtype=np.uint8)
        flen, fst = ctx.frames_compress(inp, in_off, in_len, out, out_off, caps, s)
        assert not fst.any()
        frames = []
        for i, d in enumerate(datas):
            got = out[int(out_off[i]): int(out_off[i]) + int(flen[i])].tobytes()
            assert got == oracle.frame_compress(d, block_size=64 << 10, block_checksums=True)[1], i
            frames.append(got)
        # decode them back in one batch (dense output layout), corrupting two frames on the way
        frames[5] = frames[5][:-1] + bytes([frames[5][-1] ^ 1])
        frames[50] = frames[50][:len(frames[50]) // 2]
        fl = np.array([len(f) for f in frames], dtype=np.uint64)
        fo = np.zeros(len(frames), dtype=np.uint64); fo[1:] = np.cumsum(fl)[:-1]
        fin = np.frombuffer(b"".join(frames), dtype=np.uint8)
        back = np.zeros(int(in_len.sum()) + 1, dtype=np.uint8)
        olen, st, det = ctx.frames_decompress(fin, fo, fl, back, in_off, in_len)
        for i, d in enumerate(datas):
            orc, odet, oplain, _c = oracle.frame_decompress(frames[i], cap=len(d))
            assert (st[i], det[i], olen[i]) == (orc, odet, len(oplain)), i
            assert back[int(in_off[i]): int(in_off[i]) + len(oplain)].tobytes() == oplain
        assert st[5] == _native.F_FRAME_CHECKSUM_FAIL and st[50] != 0
        # 150 blocks through ONE CTA: the per-CTA XXH32 queue (ring of 64) wraps twice
        class _B:
            pass
        b = _B()
        b.ctx = ctx
        parity.check_batched_blocks(b, oracle, [d for d in datas if d] + [d[:777] for d in datas if len(d) > 800])
    finally:
        ctx.close()
        _native._lib, _native._lib_path = saved


def test_packed17_table_window_edges(emu, oracle):
    """Blocks > 64 KiB use the packed 17-bit table: repeats at distances around 64 KiB and 128 KiB
    (where a modulo-2^17 position would alias) must be accepted / rejected exactly like the reference."""
    from lz_fear_b200 import workloads as W
    rng = np.random.default_rng(17)
    a = rng.integers(0, 256, 3000, dtype=np.uint8).tobytes()
    inputs = []
    for dist in (65535, 65536, 65537, 131071, 131072, 131073, 196608, 70000):
        filler = W.text(dist - len(a), dist).numpy().tobytes()
        inputs.append(a + filler + a + W.lowent(5000, dist).numpy().tobytes())
    inputs.append(W.text(400000, 99).numpy().tobytes())
    inputs.append(bytes(300000))
    parity.check_raw_compress(emu, oracle, inputs, caps=False)


def _dependent_frames(oracle):
    from lz_fear_b200 import workloads as W
    data = W.text(90000, 41).numpy().tobytes() + W.lowent(50000, 42).numpy().tobytes() + W.text(30000, 41).numpy().tobytes()
    dic = W.text(70000, 43).numpy().tobytes()
    frames = []
    for kw in (dict(independent_blocks=False, block_size=64 << 10),
               dict(independent_blocks=False, block_size=64 << 10, block_checksums=True),
               dict(independent_blocks=False, block_size=64 << 10, content_checksum=False)):
        rc, fr = oracle.frame_compress(data, **kw)
        assert rc == 0
        frames.append(fr)
    return data, dic, frames


def test_dependent_block_frames_decode(emu, oracle, issue15_input):     # tests/issue-15.rs shape
    data, dic, frames = _dependent_frames(oracle)
    for fr in frames:
        st, det, plain, cons = emu.ctx.frame_decompress(fr, cap=len(data) + 16)
        assert (st, plain, cons) == (0, data, len(fr))
    rc, fr = oracle.frame_compress(issue15_input, independent_blocks=False, block_size=64 << 10)
    assert emu.ctx.frame_decompress(fr, cap=len(issue15_input) + 16)[:3] == (0, 0, issue15_input)
    # mutated dependent frames: same status / detail / delivered plaintext as the oracle
    muts = [parity.mutate(frames[1], 900 + k, k=1 + k % 2) for k in range(8)] + [frames[0][:n] for n in (70000, len(frames[0]) - 3)]
    parity.check_frame_decode_errors(emu, oracle, muts)


def test_dictionary_frames_decode(emu, oracle):
    data, dic, _ = _dependent_frames(oracle)
    for kw in (dict(block_size=64 << 10), dict(independent_blocks=False, block_size=64 << 10)):
        rc, fr = oracle.frame_compress(data, dictionary=dic, dictionary_id=7, **kw)
        assert rc == 0
        want = oracle.frame_decompress(fr, dictionary=dic, cap=len(data) + 16)
        assert want[0] == 0 and want[2] == data
        got = emu.ctx.frame_decompress(fr, dictionary=dic, cap=len(data) + 16)
        assert got == want
        # the wrong (short) dictionary fails exactly like the reference
        assert emu.ctx.frame_decompress(fr, dictionary=dic[:100], cap=len(data) + 16)[:3] == \
            oracle.frame_decompress(fr, dictionary=dic[:100], cap=len(data) + 16)[:3]


def test_dependent_frames_with_short_blocks(emu, oracle):
    """hand-crafted dependent frames whose non-final blocks are short: the exact, block-at-a-time path"""
    import struct
    hdr = bytes([0x04, 0x22, 0x4D, 0x18, 0x40, 0x40, 0xC0])         # version 1, dependent, no checksums, 64 KiB
    assert oracle.parse_header(hdr)[0] == 0
    a = bytes([0x80]) + b"abcdefgh"                                  # 8 literals
    b = bytes([0x00, 8, 0, 0x40]) + b"wxyz"                          # match 4 @ offset 8 (reaches block 1), 4 literals
    c = bytes([0x04, 12, 0, 0x10, 0x21])                             # match 8 @ offset 12, 1 literal
    bad = bytes([0x00, 40, 0, 0x10, 0x21])                           # offset beyond the window
    big = oracle.compress_block(bytes(65536))[1]
    W = lambda blk: struct.pack("<I", len(blk)) + blk
    frames = [hdr + W(a) + W(b) + W(c) + struct.pack("<I", 0),
              hdr + W(a) + W(b) + W(bad) + struct.pack("<I", 0),
              hdr + W(a) + W(big) + W(b) + W(c) + struct.pack("<I", 0),
              hdr + W(a) + W(b) + struct.pack("<I", 3 | 0x80000000) + b"RAW" + W(c) + struct.pack("<I", 0)]
    parity.check_frame_decode_errors(emu, oracle, frames)
    dic = b"0123456789" * 10
    far = bytes([0x00, 30, 0, 0x10, 0x21])                           # offset 30 at o=0: into the dictionary
    parity.check_frame_decode_errors(emu, oracle, [hdr + W(far) + W(b) + W(c) + struct.pack("<I", 0)], dictionary=dic)


def _dep_inputs():
    from lz_fear_b200 import workloads as W
    t = W.text(140000, 41).numpy().tobytes()
    return [b"", b"abc", t[:70000], t + W.lowent(40000, 42).numpy().tobytes() + t[:30000], bytes(150000),
            W.random_bytes(140000, 5).numpy().tobytes(), t[:65536], t[:65537], t[:131072]]


def test_dependent_block_frames_compress(emu, oracle, issue15_input):   # compress.rs:220,271-275; tests/issue-15.rs
    for kw in (dict(independent_blocks=False, block_size=64 << 10),
               dict(independent_blocks=False, block_size=64 << 10, block_checksums=True, content_checksum=False)):
        for data in _dep_inputs() + [issue15_input]:
            st, frame = emu.ctx.frame_compress(data, **kw)
            orc, oframe = oracle.frame_compress(data, **kw)
            assert (st, frame) == (orc, oframe), (kw, len(data))
            assert emu.ctx.frame_decompress(frame, cap=len(data) + 16)[:3] == (0, 0, data)


def test_dictionary_frames_compress(emu, oracle):                       # compress.rs:202-214
    from lz_fear_b200 import workloads as W
    dic_small = [1, 3, 3, 7]                                             # tests/output_equivalence.rs:44
    dic = W.text(70000, 43).numpy().tobytes()
    for d in (bytes(dic_small), dic[:1000], dic):
        for kw in (dict(block_size=64 << 10), dict(independent_blocks=False, block_size=64 << 10, block_checksums=True)):
            for data in _dep_inputs()[2:5]:
                st, frame = emu.ctx.frame_compress(data, dictionary=d, dictionary_id=9, **kw)
                orc, oframe = oracle.frame_compress(data, dictionary=d, dictionary_id=9, **kw)
                assert (st, frame) == (orc, oframe), (len(d), kw, len(data))
                assert emu.ctx.frame_decompress(frame, dictionary=d, cap=len(data) + 16)[:3] == (0, 0, data)


def test_packed17_long_matches_never_alias(emu, oracle):
    parity.check_raw_compress(emu, oracle, parity.long_match_inputs(), caps=False)
    # the same through the batched call with the max_block_len promise (packed 17-bit slots in the big CTA)
    inputs = parity.long_match_inputs()
    parity.check_batched_blocks(emu, oracle, inputs, max_block_len=max(len(b) for b in inputs))


def test_short_nonfinal_blocks_with_exact_capacity(emu, oracle):
    parity.check_short_block_frames(emu, oracle)


def test_raw_compress2_with_history_and_carried_table(emu, oracle):  # src/raw/compress/mod.rs:165-170
    parity.check_raw_compress2_with_history(emu, oracle)
    parity.check_raw_compress2_with_history(emu, oracle, table_kind=N.TABLE_U16)


def test_segmented_parse_is_valid_lz4_of_reference_size(emu, oracle):
    worst = parity.check_segmented_parse(emu, oracle)
    assert worst < 0.01


def test_streaming_reader_and_writer(emu, oracle):               # src/framed/decompress.rs:46-77, examples/delz4.rs
    parity.check_streaming_host_mirror(emu, oracle)


def test_raw_mirror_compress2_with_history(emu, oracle):
    parity.check_raw_mirror_with_history(emu, oracle)


def test_seeded_structural_fuzz(emu, oracle):                   # 3900 more cases of the same generator were run once by hand
    parity.check_fuzz_blocks(emu, oracle, seed=11, count=60, max_len=60000)
    parity.check_fuzz_frames(emu, oracle, seed=11, count=35, max_len=60000)
    parity.check_fuzz_frame_batches(emu, oracle, seed=11, count=12, max_len=50000)
    parity.check_fuzz_frame_batches(emu, oracle, seed=12, count=8, max_len=50000, device_api="cpu")
    parity.check_fuzz_block_batches(emu, oracle, seed=11, count=4, max_len=40000)


def test_examples_dolz4_delz4(emu, oracle, simt_lib_path, tmp_path):        # examples/dolz4.rs, examples/delz4.rs
    """The two example programs as a user runs them (separate processes, files in and out): dolz4 writes the frame the
    reference writes (compress_with_size: content size in the header), delz4 restores the file."""
    parity.check_examples(oracle, tmp_path, simt_lib_path)


def test_dependent_frame_beyond_the_position_limit_panics_alone(emu, oracle, monkeypatch):   # src/raw/compress/mod.rs:67
    monkeypatch.setenv("LZF_B200_TEST_POS_LIMIT", "200000")
    with emu.fresh() as b:
        parity.check_dependent_frame_position_limit(b, oracle)
