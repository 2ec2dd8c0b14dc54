"""Pins the CPU oracle (oracle/lzf_oracle.c) against everything the reference's own tests hold for
the hot path (tests/golden/, lifted by make_golden.py) and against independent implementations:
python xxhash for XXH32 and liblz4 1.9.4, which tests/output_equivalence.rs names as the expected
output of the compressor."""
import ctypes
import io
import struct

import numpy as np
import pytest

xxhash = pytest.importorskip("xxhash")


def big_pattern(n):
    i = np.arange(n, dtype=np.uint64).astype(np.uint8)
    return ((i * np.uint8(0xA) + np.uint8(33)) ^ np.uint8(0xA2)).astype(np.uint8)     # src/lib.rs:101


def test_decode_kats(oracle, vectors):                         # src/raw/decompress.rs:153-175
    for kat in vectors["decode_kats"]:
        st, out, n = oracle.decompress_raw(bytes(kat["input"]), out_limit=(1 << 62))
        if kat["output"] is None:
            assert st != 0, kat["name"]
        else:
            assert st == 0 and out == bytes(kat["output"]), kat["name"]
    assert oracle.decompress_raw(bytes([0x10, 97, 2, 0]))[0] == oracle.INVALID_DEDUP_OFFSET
    assert oracle.decompress_raw(bytes([0x40, 97, 1, 0]))[0] == oracle.UNEXPECTED_END


def test_roundtrip_strings(oracle, vectors):                   # src/lib.rs:24-95
    for name, strings in vectors["roundtrip_strings"].items():
        for s in strings:
            data = s.encode("latin-1")
            table = oracle.TABLE_U16 if len(data) <= 0xFFFF else oracle.TABLE_U32
            st, comp = oracle.compress_block(data, table=table)
            assert st == 0
            st, out, _ = oracle.decompress_raw(comp)
            assert st == 0 and out == data, (name, s)
            if name == "compression_works":
                assert len(comp) < len(data)


def test_empty_input_writes_nothing(oracle):                   # src/raw/compress/mod.rs:171
    assert oracle.compress_block(b"") == (0, b"")
    assert oracle.decompress_raw(b"") == (0, b"", 0)


def test_big_compression(oracle):                              # src/lib.rs:97-106 (80 000 000 bytes)
    data = big_pattern(80_000_000)
    st, comp = oracle.compress_block(data)
    assert st == 0 and len(comp) < len(data) // 100
    st, out, n = oracle.decompress_raw(comp, cap=len(data) + len(comp) + 16)
    assert st == 0 and n == len(data) and np.array_equal(np.frombuffer(out, dtype=np.uint8), data)


def test_config1_zero_block_kat(oracle):                       # BASELINE config 1
    z = bytes(65536)
    st, comp = oracle.compress_block(z)
    assert st == 0 and len(comp) == 267
    assert comp == bytes([0x1F, 0, 1, 0]) + b"\xff" * 256 + bytes([0xE7, 0x50, 0, 0, 0, 0, 0])
    rc, frame = oracle.frame_compress(z)
    assert rc == 0 and len(frame) == 286
    assert frame[:11] == bytes([0x04, 0x22, 0x4D, 0x18, 0x64, 0x70, 0xB9, 0x0B, 0x01, 0, 0])
    assert frame[-8:] == bytes([0, 0, 0, 0, 0x1C, 0xE8, 0x64, 0x0F])
    assert oracle.frame_decompress(frame)[:3] == (0, 0, z)


def test_xxh32_matches_python_xxhash(oracle):
    assert oracle.xxh32(b"") == 0x02CC5D05
    assert oracle.xxh32(bytes(65536)) == 0x0F64E81C
    rng = np.random.default_rng(0)
    for n in list(range(0, 70)) + [255, 256, 1000, 65536, 1 << 20]:
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle.xxh32(d) == xxhash.xxh32(d, seed=0).intdigest(), n
    for flg_bd, hc in [((0x64, 0x70), 0xB9), ((0x64, 0x40), 0xA7), ((0x60, 0x70), 0x73), ((0x74, 0x70), 0x8E)]:
        assert (oracle.xxh32(bytes(flg_bd)) >> 8) & 0xFF == hc


def test_issue15_dependent_blocks_roundtrip(oracle, issue15_input):        # tests/issue-15.rs
    rc, frame = oracle.frame_compress(issue15_input, independent_blocks=False, block_size=64 << 10)
    assert rc == 0
    rc, det, plain, cons = oracle.frame_decompress(frame)
    assert (rc, plain, cons) == (0, issue15_input, len(frame))


def test_roundtrip_fuzz_corpus(oracle, corpora):               # fuzz/fuzz_targets/roundtrip_fuzz.rs
    for name, data in corpora["roundtrip_fuzz"]:
        rc, frame = oracle.frame_compress(data, content_checksum=True, independent_blocks=True)
        assert rc == 0, name
        rc, det, plain, cons = oracle.frame_decompress(frame)
        assert (rc, plain) == (0, data), name


class _Prefs(ctypes.Structure):       # LZ4F_preferences_t of liblz4 1.9.x
    _fields_ = [("blockSizeID", ctypes.c_int), ("blockMode", ctypes.c_int), ("contentChecksumFlag", ctypes.c_int),
                ("frameType", ctypes.c_int), ("contentSize", ctypes.c_ulonglong), ("dictID", ctypes.c_uint),
                ("blockChecksumFlag", ctypes.c_int), ("compressionLevel", ctypes.c_int), ("autoFlush", ctypes.c_uint),
                ("favorDecSpeed", ctypes.c_uint), ("reserved", ctypes.c_uint * 3)]


def lz4f_compress(lib, data, **kw):
    p = _Prefs()
    for k, v in kw.items():
        setattr(p, k, v)
    lib.LZ4F_compressFrameBound.restype = ctypes.c_size_t
    lib.LZ4F_compressFrameBound.argtypes = [ctypes.c_size_t, ctypes.c_void_p]
    lib.LZ4F_compressFrame.restype = ctypes.c_size_t
    lib.LZ4F_compressFrame.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p]
    cap = lib.LZ4F_compressFrameBound(len(data), ctypes.byref(p))
    dst = ctypes.create_string_buffer(cap)
    n = lib.LZ4F_compressFrame(dst, cap, data, len(data), ctypes.byref(p))
    assert n <= cap
    return dst.raw[:n]


def test_interop_decode_corpus(oracle, corpora, liblz4):       # fuzz/fuzz_targets/interop_decode.rs
    """C lz4 (level 4, i.e. LZ4HC, content checksum on) -> our decoder: also pins XXH32 of the
    content checksum and the header checksum against the C implementation."""
    for name, data in corpora["interop_decode"]:
        frame = lz4f_compress(liblz4, data, compressionLevel=4, contentChecksumFlag=1)
        rc, det, plain, cons = oracle.frame_decompress(frame)
        assert (rc, plain) == (0, data), name
    # dependent blocks + block checksums written by C
    data = corpora["interop_decode"][-1][1] * 40
    frame = lz4f_compress(liblz4, data, blockSizeID=4, blockMode=0, contentChecksumFlag=1, blockChecksumFlag=1)
    assert oracle.frame_decompress(frame)[:3] == (0, 0, data)


def test_decode_corpus_never_crashes_and_agrees_with_c(oracle, corpora, liblz4):      # fuzz/fuzz_targets/decode.rs
    liblz4.LZ4F_createDecompressionContext.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint]
    liblz4.LZ4F_decompress.restype = ctypes.c_size_t
    liblz4.LZ4F_decompress.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p,
                                       ctypes.POINTER(ctypes.c_size_t), ctypes.c_void_p]
    liblz4.LZ4F_isError.argtypes = [ctypes.c_size_t]
    liblz4.LZ4F_freeDecompressionContext.argtypes = [ctypes.c_void_p]
    n_ok = 0
    statuses = set()
    for name, blob in corpora["decode"]:
        rc, det, plain, cons = oracle.frame_decompress(blob, cap=1 << 24)
        statuses.add(rc)
        if rc != 0:
            continue
        # frames our decoder accepts and C also accepts must decode to the same bytes
        dctx = ctypes.c_void_p()
        assert liblz4.LZ4F_createDecompressionContext(ctypes.byref(dctx), 100) == 0
        dst = ctypes.create_string_buffer(max(len(plain), 1) + (4 << 20))
        dn, sn = ctypes.c_size_t(len(dst)), ctypes.c_size_t(len(blob))
        r = liblz4.LZ4F_decompress(dctx, dst, ctypes.byref(dn), blob, ctypes.byref(sn), None)
        liblz4.LZ4F_freeDecompressionContext(dctx)
        if not liblz4.LZ4F_isError(r) and r == 0:
            assert dst.raw[: dn.value] == plain, name
            n_ok += 1
    assert n_ok >= 1
    assert {oracle.F_WRONG_MAGIC, oracle.F_INPUT_ERROR, oracle.F_CODEC_ERROR} <= statuses


def test_compress_equals_c_lz4_where_they_coincide(oracle, liblz4):       # tests/output_equivalence.rs:94-97, README:4
    """Same-size classes as SURVEY §8(c): U32 table inputs > 64 KiB + 11 whose tail the end rule of
    the two implementations treats alike (long match / literal tails probed with step 1)."""
    from lz_fear_b200 import workloads as W
    n_equal = 0
    cases = [bytes(70000), W.lowent(200000, 3).numpy().tobytes(), W.text(300000, 4).numpy().tobytes(),
             big_pattern(500000).tobytes(), W.text(1 << 20, 5).numpy().tobytes()]
    for data in cases:
        st, comp = oracle.compress_block(data)
        cap = liblz4.LZ4_compressBound(len(data))
        dst = ctypes.create_string_buffer(cap)
        n = liblz4.LZ4_compress_default(data, dst, len(data), cap)
        ref = dst.raw[:n]
        # sizes within 1 % always; bytes identical in the common case
        assert abs(len(comp) - len(ref)) <= max(16, len(ref) // 100)
        n_equal += comp == ref
        out = ctypes.create_string_buffer(len(data))
        assert liblz4.LZ4_decompress_safe(comp, out, len(comp), len(data)) == len(data) and out.raw == data
    assert n_equal >= 3


def test_frame_settings_and_errors(oracle):
    data = bytes(range(256)) * 1000
    for bs in (64 << 10, 256 << 10, 1 << 20, 4 << 20):
        for bc in (False, True):
            for cc in (False, True):
                rc, frame = oracle.frame_compress(data, block_size=bs, block_checksums=bc, content_checksum=cc)
                assert rc == 0
                assert oracle.frame_decompress(frame)[:3] == (0, 0, data)
    assert oracle.frame_compress(data, block_size=12345)[0] == oracle.F_INVALID_BLOCK_SIZE     # compress.rs:183
    assert oracle.frame_compress(data, block_size=8 << 20)[0] == oracle.F_INVALID_BLOCK_SIZE
    assert oracle.frame_compress(data, block_size=16 << 20)[0] == oracle.F_PANIC              # header.rs:55 unwrap
    rc, frame = oracle.frame_compress(data, block_size=64 << 10, block_checksums=True)
    bad = bytearray(frame); bad[20] ^= 1
    assert oracle.frame_decompress(bytes(bad))[0] == oracle.F_BLOCK_CHECKSUM_FAIL
    bad = bytearray(frame); bad[-1] ^= 1
    assert oracle.frame_decompress(bytes(bad))[0] == oracle.F_FRAME_CHECKSUM_FAIL
    bad = bytearray(frame); bad[6] ^= 1
    assert oracle.frame_decompress(bytes(bad))[0] == oracle.F_HEADER_CHECKSUM_FAIL
    assert oracle.frame_decompress(frame[:-3])[0] == oracle.F_INPUT_ERROR
    assert oracle.frame_decompress(b"\x00" * 8)[0] == oracle.F_WRONG_MAGIC
    # incompressible input is stored (compress.rs:250-255)
    rnd = np.random.default_rng(1).integers(0, 256, 100000, dtype=np.uint8).tobytes()
    rc, frame = oracle.frame_compress(rnd, block_size=64 << 10)
    assert struct.unpack_from("<I", frame, 7)[0] == (65536 | 0x80000000)
    assert oracle.frame_decompress(frame)[:3] == (0, 0, rnd)


def test_empty_block_ends_read_to_end(oracle):                 # src/framed/decompress.rs:54-61,286
    hdr = bytes([0x04, 0x22, 0x4D, 0x18, 0x60, 0x40, 0x82])
    assert oracle.parse_header(hdr)[0] == 0
    blk_a = bytes([0x30]) + b"abc"
    frame = hdr + struct.pack("<I", 4) + blk_a + struct.pack("<I", 1) + b"\x00" + struct.pack("<I", 4) + blk_a + struct.pack("<I", 0)
    rc, det, plain, cons = oracle.frame_decompress(frame)
    assert (rc, plain) == (0, b"abc") and cons == len(hdr) + 8 + 5
