"""`-m gpu` tier: the shipped liblzfear_b200.so on a real B200, through the C ABI, against the
oracle — bit-exact on the same seeded inputs, the committed golden fixtures, and size-independent
properties at BASELINE.json sizes."""
import numpy as np
import pytest

import parity
from lz_fear_b200 import _native as N
from lz_fear_b200 import workloads as W

pytestmark = pytest.mark.gpu


def _strings(vectors):
    return [s.encode("latin-1") for v in vectors["roundtrip_strings"].values() for s in v]


def test_raw_compress_matches_oracle(gpu, oracle, vectors):
    parity.check_raw_compress(gpu, oracle, parity.sample_inputs(scale=16) + _strings(vectors))


def test_raw_compress_u16_table(gpu, oracle, vectors):
    inputs = [b for b in parity.sample_inputs() + _strings(vectors) if len(b) <= 0xFFFF]
    parity.check_raw_compress(gpu, oracle, inputs, table=N.TABLE_U16)
    assert gpu.ctx.raw_compress_into(bytes(70000), table=N.TABLE_U16)[0] == N.PANIC


def test_raw_compress_large_hashlog(gpu, oracle):
    inputs = [b for b in parity.sample_inputs(scale=16) if len(b) >= 1000][:12]
    for hashlog in (13, 14, 16):
        parity.check_raw_compress(gpu, oracle, inputs, hashlog=hashlog)


def test_packed17_table_window_edges(gpu, oracle):
    """Repeats at distances around 64 KiB / 128 KiB / multiples inside 4 MiB blocks (packed 17-bit table + sweeps)."""
    rng = np.random.default_rng(17)
    a = rng.integers(0, 256, 3000, dtype=np.uint8).tobytes()
    inputs = []
    for dist in (65535, 65536, 65537, 131071, 131072, 131073, 196608, 262143, 262144, 1 << 20, (1 << 21) + 1):
        filler = W.text(dist - len(a), dist).numpy().tobytes()
        inputs.append(a + filler + a + W.lowent(5000, dist).numpy().tobytes())
    inputs += [W.text(4 << 20, 99).numpy().tobytes(), bytes(4 << 20), W.random_bytes(1 << 20, 5).numpy().tobytes() * 4,
               W.text(17 << 20, 98).numpy().tobytes()]           # the last one is beyond 16 MiB: plain u32 table
    parity.check_raw_compress(gpu, oracle, inputs, caps=False)


def test_decode_kats(gpu, oracle, vectors):
    parity.check_raw_decompress(gpu, oracle, [(bytes(k["input"]), None) for k in vectors["decode_kats"]])
    assert gpu.ctx.raw_decompress(bytes([0x11, 97, 1, 0, 0x22, 98, 99, 2, 0]))[:2] == (0, b"aaaaaabcbcbcbc")


def test_raw_decompress_matches_oracle(gpu, oracle):
    blocks = [(oracle.compress_block(d)[1], len(d)) for d in parity.sample_inputs(scale=16)]
    parity.check_raw_decompress(gpu, oracle, blocks)


def test_raw_decompress_malformed(gpu, oracle):
    base = [oracle.compress_block(d)[1] for d in parity.sample_inputs() if 20 <= len(d) <= 70000]
    blocks = [(parity.mutate(comp, 100 * i + k, k=1 + k % 3), None) for i, comp in enumerate(base) for k in range(8)]
    blocks += [(bytes([0x0F]), None), (bytes([0xF0]), None), (bytes([0xF0, 0xFF]), None), (bytes([0, 0, 0]), None),
               (bytes([0x10, 65, 0x01]), None), (bytes([0x1F, 65, 1, 0, 0xFF]), None), (bytes([0x00, 1, 0]), None)]
    parity.check_raw_decompress(gpu, oracle, blocks)


def test_raw_decompress_with_prefix(gpu, oracle):
    prefix = b"0123456789abcdef" * 8
    comp = bytes([0x42, 120, 121, 122, 119, 20, 0, 0x00, 3, 0])
    for lim in (6, 10, 1 << 20):
        assert gpu.ctx.raw_decompress(comp, prefix=prefix, out_limit=lim, cap=256) == \
            oracle.decompress_raw(comp, prefix=prefix, out_limit=lim, cap=256)
    assert gpu.ctx.raw_decompress(comp, prefix=prefix[:10], out_limit=1 << 20, cap=256)[0] == N.INVALID_DEDUP_OFFSET


def test_big_compression_roundtrip(gpu, oracle):               # src/lib.rs:97-106, 80 000 000 bytes
    i = np.arange(80_000_000, dtype=np.uint64).astype(np.uint8)
    data = ((i * np.uint8(0xA) + np.uint8(33)) ^ np.uint8(0xA2)).astype(np.uint8)
    st, comp = gpu.ctx.raw_compress_into(data)
    assert (st, comp) == oracle.compress_block(data)
    st, out, n = gpu.ctx.raw_decompress(comp, cap=len(data) + 16)
    assert st == 0 and n == len(data) and np.array_equal(np.frombuffer(out, dtype=np.uint8), data)


def test_batched_blocks_device(gpu, oracle):
    inputs = [b for b in parity.sample_inputs(scale=16) if len(b) > 0]
    parity.check_batched_blocks(gpu, oracle, inputs, use_torch_device="cuda")


def test_frames_roundtrip_and_bytes(gpu, oracle):
    s = parity.sample_inputs(scale=16)
    inputs = [b"", b"a", bytes(65536), s[6], s[7][:70001], s[5] * 30, s[-1] + s[-3] + s[-5][:1234567]]
    parity.check_frames(gpu, oracle, inputs)
    st, frame = gpu.ctx.frame_compress(bytes(65536))              # BASELINE config 1 KAT
    assert st == 0 and len(frame) == 286 and frame[-4:] == bytes([0x1C, 0xE8, 0x64, 0x0F])


def test_frame_decode_corpus(gpu, oracle, corpora):
    parity.check_frame_decode_errors(gpu, oracle, [b for _, b in corpora["decode"]])


def test_roundtrip_and_interop_corpora(gpu, oracle, corpora):
    for name, data in corpora["roundtrip_fuzz"] + corpora["interop_decode"]:
        st, frame = gpu.ctx.frame_compress(data)
        assert (st, frame) == oracle.frame_compress(data), name
        st, det, plain, cons = gpu.ctx.frame_decompress(frame, cap=len(data) + 16)
        assert (st, plain) == (0, data), name


def test_frame_decode_mutations(gpu, oracle):
    s = parity.sample_inputs()
    data = s[6] + s[5]
    frames = []
    for kw in parity.FRAME_SETTINGS[1:5]:
        rc, frame = oracle.frame_compress(data, **kw)
        for k in range(40):
            frames.append(parity.mutate(frame, hash(str(sorted(kw.items()))) % 1000 + k, k=1 + k % 3))
        frames += [frame[:n] for n in (0, 3, 6, 7, 10, 11, len(frame) - 5, len(frame) - 1)]
    parity.check_frame_decode_errors(gpu, oracle, frames)


def test_config2_seq50_decompress_vs_oracle(gpu, oracle):
    """4096 of the config-2 blocks bit-exact against the oracle (the full 4 GiB runs in bench.py)."""
    import torch
    nb = 4096
    comp, off, ln = W.seq50_blocks(nb, device="cuda")
    out = torch.empty(nb * 65536, dtype=torch.uint8, device="cuda")
    out_off = torch.arange(nb, device="cuda", dtype=torch.int64) * 65536
    cap = torch.full((nb,), 65536, dtype=torch.int32, device="cuda")
    olen = torch.zeros(nb, dtype=torch.int32, device="cuda")
    st = torch.zeros(nb, dtype=torch.int32, device="cuda")
    xx = torch.zeros(nb, dtype=torch.int32, device="cuda")
    gpu.ctx.decompress_blocks(comp, off, ln, nb, out, out_off, cap, cap, olen, st, xx)
    torch.cuda.synchronize()
    assert int(st.abs().sum()) == 0 and bool((olen == 65536).all())
    c, o = comp.cpu().numpy(), out.cpu().numpy()
    offs, lens = off.cpu().numpy(), ln.cpu().numpy()
    ref = np.empty(nb * 65536, dtype=np.uint8)
    ool, ost = oracle.decompress_blocks_mt(c, offs.astype(np.uint64), lens.astype(np.uint32), ref,
                                           out_off.cpu().numpy().astype(np.uint64), np.full(nb, 65536, np.uint32),
                                           np.full(nb, 65536, np.uint32), nthreads=8)
    assert not ost.any() and np.array_equal(ref, o)
    xs = xx.cpu().numpy().view(np.uint32)
    for b in range(0, nb, 97):
        assert xs[b] == oracle.xxh32(ref[b * 65536:(b + 1) * 65536])


@pytest.mark.parametrize("variant", ["u32-small-ctas", "packed-28-warps", "packed-all-global", "u32-all-global"])
def test_config3_text_compress_vs_oracle(gpu, oracle, monkeypatch, variant):
    """64 of the config-3 blocks (4 MiB text) byte-identical to the oracle, then decoded back — on every table
    layout of the encode kernel: u32 slots in 4-warp CTAs (no max_block_len promise), packed 17-bit slots in the
    28-warp CTA (13 tables in shared memory, the rest in the L2 scratch), all tables in the scratch, plain u32 there."""
    import torch
    nb, B = 64, 4 << 20
    mbl = 0 if variant == "u32-small-ctas" else B
    if variant in ("packed-all-global", "u32-all-global"):
        monkeypatch.setenv("LZF_B200_ENC_SMEM_WARPS", "0")
    if variant == "u32-all-global":
        monkeypatch.setenv("LZF_B200_ENC_U32", "1")
    data = W.TextSource(device="cuda").make(nb * B)
    off = torch.arange(nb, device="cuda", dtype=torch.int64) * B
    ln = torch.full((nb,), B, dtype=torch.int32, device="cuda")
    comp = torch.empty(nb * B, dtype=torch.uint8, device="cuda")
    olen = torch.zeros(nb, dtype=torch.int32, device="cuda")
    st = torch.zeros(nb, dtype=torch.int32, device="cuda")
    with gpu.fresh() as g:                                   # the knobs are read when the context is created
        g.ctx.compress_blocks(data, off, ln, nb, comp, off, None, olen, st, max_block_len=mbl)
        torch.cuda.synchronize()
    assert int(st.abs().sum()) == 0
    h = data.cpu().numpy()
    ref = np.empty(nb * B, dtype=np.uint8)
    rlen, rst = oracle.compress_blocks_mt(h, off.cpu().numpy().astype(np.uint64), np.full(nb, B, np.uint32), ref,
                                          off.cpu().numpy().astype(np.uint64), nthreads=8)
    got_len = olen.cpu().numpy().view(np.uint32)
    assert np.array_equal(got_len, rlen)
    g = comp.cpu().numpy()
    for b in range(nb):
        assert np.array_equal(g[b * B: b * B + rlen[b]], ref[b * B: b * B + rlen[b]]), b
    # and back
    plain = torch.empty(nb * B, dtype=torch.uint8, device="cuda")
    cap = torch.full((nb,), B, dtype=torch.int32, device="cuda")
    gpu.ctx.decompress_blocks(comp, off, olen, nb, plain, off, cap, cap, olen.clone(), st, None)
    torch.cuda.synchronize()
    assert int(st.abs().sum()) == 0 and torch.equal(plain, data)


def test_mixed_entropy_frames_roundtrip_property(gpu, oracle):
    """config-4 style: random / text / lowent blocks; every frame byte-identical to the oracle's, encode -> decode is
    the identity, stored blocks appear exactly for the random class."""
    import torch
    B, nb = 1 << 20, 48
    data = W.mixed_blocks(nb, B, device="cuda")
    s, _k = N.make_settings(block_size=B)
    nframes = 6
    per = nb // nframes * B
    in_off = np.arange(nframes, dtype=np.uint64) * per
    in_len = np.full(nframes, per, dtype=np.uint64)
    bound = gpu.ctx.frame_bound(s, per)
    out_off = np.arange(nframes, dtype=np.uint64) * ((bound + 255) // 256 * 256)
    frames = torch.empty(int(out_off[-1]) + bound, dtype=torch.uint8, device="cuda")
    flen, fst = gpu.ctx.frames_compress_device(data, in_off, in_len, frames, out_off, np.full(nframes, bound, np.uint64), s)
    assert not fst.any()
    back = torch.empty_like(data)
    olen, st, det = gpu.ctx.frames_decompress_device(frames, out_off, flen, back, in_off, in_len)
    assert not st.any() and (olen == in_len).all() and torch.equal(back, data)
    f0 = frames[: int(flen[0])].cpu().numpy()
    first_word = int.from_bytes(f0[7:11].tobytes(), "little")
    assert first_word == (B | 0x80000000)                         # block 0 is random -> stored
    h, fr = data.cpu().numpy(), frames.cpu().numpy()
    for f in range(nframes):                                      # every frame against the oracle, byte for byte
        want = oracle.frame_compress(h[f * per:(f + 1) * per].tobytes(), block_size=B)
        assert (0, fr[int(out_off[f]): int(out_off[f]) + int(flen[f])].tobytes()) == want, f
        assert oracle.frame_decompress(want[1], cap=per)[:3] == (0, 0, h[f * per:(f + 1) * per].tobytes())


def test_streaming_xxh32_and_host_mirror(gpu, oracle):
    import io
    import lz_fear_b200 as L
    L.raw.set_default_context(gpu.ctx)
    data = W.text(300000, 77).numpy().tobytes()
    assert gpu.ctx.xxh32(data) == oracle.xxh32(data)
    out = io.BytesIO()
    L.CompressionSettings.default().block_size(64 << 10).block_checksums(True).compress(io.BytesIO(data), out)
    assert out.getvalue() == oracle.frame_compress(data, block_size=64 << 10, block_checksums=True)[1]
    rd = L.LZ4FrameReader(io.BytesIO(out.getvalue()))
    assert rd.block_size() == 64 << 10 and rd.frame_size() is None
    assert rd.into_read().read_to_end() == data
    assert L.decompress_frame(io.BytesIO(out.getvalue())) == data
    buf = bytearray()
    L.raw.decompress_raw(oracle.compress_block(data)[1], b"", buf, 1 << 30)
    assert bytes(buf) == data
    L.raw.set_default_context(None)


def test_dependent_and_dictionary_frames(gpu, oracle, issue15_input, corpora, liblz4):
    """Decode side of SURVEY §8(f) rank 2: dependent-block frames (block i waits for block i-1 inside the
    kernel), dictionaries, C-written linked frames, and the block-at-a-time exact path."""
    import test_simt_kernels as T
    from test_oracle import lz4f_compress
    T.test_dependent_block_frames_decode(gpu, oracle, issue15_input)
    T.test_dictionary_frames_decode(gpu, oracle)
    T.test_dependent_frames_with_short_blocks(gpu, oracle)
    data = (W.text(3 << 20, 5).numpy().tobytes() + W.lowent(1 << 20, 6).numpy().tobytes()) * 2
    for kw in (dict(blockSizeID=4, blockMode=0, contentChecksumFlag=1), dict(blockSizeID=7, blockMode=0, blockChecksumFlag=1),
               dict(blockSizeID=5, blockMode=0, compressionLevel=4, contentChecksumFlag=1)):
        frame = lz4f_compress(liblz4, data, **kw)               # blockMode 0 = linked (dependent) blocks
        st, det, plain, cons = gpu.ctx.frame_decompress(frame, cap=len(data) + 16)
        assert (st, plain) == (0, data), kw
    rc, fr = oracle.frame_compress(data, independent_blocks=False, block_size=64 << 10)
    assert gpu.ctx.frame_decompress(fr, cap=len(data) + 16)[:3] == (0, 0, data)


def test_sliced_input_feed(gpu, oracle, monkeypatch):
    """lzf_frames_compress over host buffers: the plaintext arrives slice by slice while the block kernel runs."""
    import test_simt_kernels as T
    T.test_frame_compress_with_sliced_input_feed(gpu, oracle, monkeypatch, scale=8)
    # many frames in one call, default slice size, 1 MiB blocks
    monkeypatch.setenv("LZF_B200_FEED_MIN_BLOCKS", "2")
    monkeypatch.delenv("LZF_B200_FEED_SLICE")
    nf, fp = 6, 8 << 20
    src = np.concatenate([W.mixed_blocks(8, 1 << 20, seed=77 + f).numpy() for f in range(nf)])
    s, keep = N.make_settings(block_size=1 << 20, block_checksums=True)
    bound = gpu.ctx.frame_bound(s, fp)
    out = np.zeros(nf * bound, dtype=np.uint8)
    with gpu.fresh() as g:
        fl, fs = g.ctx.frames_compress(src, np.arange(nf, dtype=np.uint64) * fp, np.full(nf, fp, np.uint64), out,
                                       np.arange(nf, dtype=np.uint64) * bound, np.full(nf, bound, np.uint64), s)
    assert not fs.any()
    for f in range(nf):
        want = oracle.frame_compress(src[f * fp:(f + 1) * fp].tobytes(), block_size=1 << 20, block_checksums=True)
        assert (0, out[f * bound:f * bound + int(fl[f])].tobytes()) == want, f
    # 4 KiB slices (256 per block): the warps outrun the feed all the time and wait on the progress word
    monkeypatch.setenv("LZF_B200_FEED_SLICE", "4096")
    out2 = np.zeros_like(out)
    with gpu.fresh() as g:
        fl2, fs2 = g.ctx.frames_compress(src, np.arange(nf, dtype=np.uint64) * fp, np.full(nf, fp, np.uint64), out2,
                                         np.arange(nf, dtype=np.uint64) * bound, np.full(nf, bound, np.uint64), s)
    assert not fs2.any() and np.array_equal(fl2, fl)
    for f in range(nf):
        assert np.array_equal(out2[f * bound:f * bound + int(fl[f])], out[f * bound:f * bound + int(fl[f])]), f


def test_batched_host_frames_with_a_refused_frame(gpu, oracle):
    import test_simt_kernels as T
    T.test_batched_host_frames_with_a_refused_frame(gpu, oracle)


def test_dependent_and_dictionary_frames_compress(gpu, oracle, issue15_input):
    """Compress side of SURVEY §8(f) rank 2: one warp carries the table through all blocks of a dependent
    frame; dictionaries prime the table.  Frames are byte-identical to the oracle's."""
    import test_simt_kernels as T
    T.test_dependent_block_frames_compress(gpu, oracle, issue15_input)
    T.test_dictionary_frames_compress(gpu, oracle)
    for kw in (dict(independent_blocks=False, block_size=256 << 10), dict(block_size=256 << 10, block_checksums=True, dictionary=bytes(range(200)) * 400)):
        for d in T._dep_inputs():
            assert gpu.ctx.frame_compress(d, **kw) == oracle.frame_compress(d, **kw)
    data = W.text(9 << 20, 7).numpy().tobytes() + W.random_bytes(5 << 20, 8).numpy().tobytes() + W.lowent(3 << 20, 9).numpy().tobytes()
    dic = W.text(100000, 10).numpy().tobytes()
    for kw in (dict(independent_blocks=False), dict(independent_blocks=False, block_size=1 << 20, block_checksums=True),
               dict(dictionary=dic, dictionary_id=1), dict(dictionary=dic, independent_blocks=False, block_size=256 << 10)):
        st, frame = gpu.ctx.frame_compress(data, **kw)
        assert (st, frame) == oracle.frame_compress(data, **kw), sorted(kw)
        d = kw.get("dictionary", b"")
        assert gpu.ctx.frame_decompress(frame, dictionary=d, cap=len(data) + 16)[:3] == (0, 0, data)


def test_packed17_long_matches_never_alias(gpu, oracle):
    inputs = parity.long_match_inputs()
    rng = np.random.default_rng(4)
    x = rng.integers(0, 256, 100000, dtype=np.uint8).tobytes()
    inputs += [x * 30 + b"!" + x * 11, bytes(3 << 20) + x + bytes(1 << 20) + x]      # 4 MiB-class blocks
    parity.check_raw_compress(gpu, oracle, inputs, caps=False)
    parity.check_batched_blocks(gpu, oracle, inputs, use_torch_device="cuda", max_block_len=max(len(b) for b in inputs))


def test_short_nonfinal_blocks_with_exact_capacity(gpu, oracle):
    parity.check_short_block_frames(gpu, oracle)


def test_raw_compress2_with_history_and_carried_table(gpu, oracle):  # src/raw/compress/mod.rs:165-170
    parity.check_raw_compress2_with_history(gpu, oracle)
    parity.check_raw_compress2_with_history(gpu, oracle, table_kind=N.TABLE_U16)


def test_segmented_parse_is_valid_lz4_of_reference_size(gpu, oracle):
    worst = parity.check_segmented_parse(gpu, oracle, sizes=(70001, 300000, 4 << 20, (4 << 20) - 77), scale=1)
    assert worst < 0.01


def test_streaming_reader_and_writer(gpu, oracle):               # src/framed/decompress.rs:46-77, examples/delz4.rs
    parity.check_streaming_host_mirror(gpu, oracle, scale=8)


def test_batched_calls_on_two_streams_do_not_race(gpu, oracle):       # shared work counter / table scratch
    """Two lzf_compress_blocks and two lzf_decompress_blocks calls queued back to back on DIFFERENT streams of one ctx:
    the second launch of each pair must wait for the first (they share the ctx's work queue), results as for one stream."""
    import torch
    B, nb = 256 << 10, 96
    datas = [W.TextSource(seed=900 + k, device="cuda").make(nb * B) for k in range(2)]
    off = torch.arange(nb, device="cuda", dtype=torch.int64) * B
    ln = torch.full((nb,), B, dtype=torch.int32, device="cuda")
    comps = [torch.zeros(nb * B, dtype=torch.uint8, device="cuda") for _ in range(2)]
    olens = [torch.zeros(nb, dtype=torch.int32, device="cuda") for _ in range(2)]
    sts = [torch.ones(nb, dtype=torch.int32, device="cuda") for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for rep in range(3):
        for k in range(2):
            gpu.ctx.compress_blocks(datas[k], off, ln, nb, comps[k], off, None, olens[k], sts[k], None, None,
                                    stream=streams[k].cuda_stream, max_block_len=B)
    torch.cuda.synchronize()
    for k in range(2):
        assert int(sts[k].abs().sum()) == 0
        h = datas[k].cpu().numpy()
        g = comps[k].cpu().numpy()
        gl = olens[k].cpu().numpy().view(np.uint32)
        for b in range(0, nb, 7):
            want = oracle.compress_block(h[b * B:(b + 1) * B].tobytes(), cap=B)
            assert (0, g[b * B: b * B + int(gl[b])].tobytes()) == want, (k, b)
    backs = [torch.zeros(nb * B, dtype=torch.uint8, device="cuda") for _ in range(2)]
    dl = [torch.zeros(nb, dtype=torch.int32, device="cuda") for _ in range(2)]
    for rep in range(3):
        for k in range(2):
            gpu.ctx.decompress_blocks(comps[k], off, olens[k], nb, backs[k], off, ln, ln, dl[k], sts[k], None, stream=streams[k].cuda_stream)
    torch.cuda.synchronize()
    for k in range(2):
        assert int(sts[k].abs().sum()) == 0 and torch.equal(backs[k], datas[k])


def test_raw_mirror_compress2_with_history(gpu, oracle):
    parity.check_raw_mirror_with_history(gpu, oracle)


def test_seeded_structural_fuzz(gpu, oracle):
    parity.check_fuzz_blocks(gpu, oracle, seed=21, count=600)
    parity.check_fuzz_frames(gpu, oracle, seed=21, count=300)
    parity.check_fuzz_frame_batches(gpu, oracle, seed=21, count=150)
    parity.check_fuzz_frame_batches(gpu, oracle, seed=22, count=100, device_api="cuda")
    parity.check_fuzz_block_batches(gpu, oracle, seed=21, count=60, use_torch_device="cuda")


def test_examples_dolz4_delz4(gpu, oracle, tmp_path):                       # examples/dolz4.rs, examples/delz4.rs
    parity.check_examples(oracle, tmp_path, None, scale=16)


def test_dependent_frame_beyond_the_position_limit_panics_alone(gpu, oracle, monkeypatch):   # src/raw/compress/mod.rs:67
    monkeypatch.setenv("LZF_B200_TEST_POS_LIMIT", "200000")
    with gpu.fresh() as b:
        parity.check_dependent_frame_position_limit(b, oracle)
