"""ctypes loader for the CPU oracle — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (rust-lz-fear_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "liblzf_oracle.so")

OK, UNEXPECTED_END, MEMORY_LIMIT_EXCEEDED, ZERO_DEDUP_OFFSET, INVALID_DEDUP_OFFSET, WRITER_FULL, OUTPUT_CAP, PANIC = range(8)
F_OK = 0
F_INPUT_ERROR, F_CODEC_ERROR, F_HEADER_PARSE_ERROR, F_WRONG_MAGIC, F_HEADER_CHECKSUM_FAIL = 10, 11, 12, 13, 14
F_BLOCK_CHECKSUM_FAIL, F_FRAME_CHECKSUM_FAIL, F_BLOCK_LENGTH_OVERFLOW, F_BLOCK_SIZE_OVERFLOW = 15, 16, 17, 18
F_INVALID_BLOCK_SIZE, F_WRITE_ERROR, F_PANIC = 20, 21, 22
TABLE_U32, TABLE_U16 = 0, 1


_NATIVE_SO = os.path.join(_DIR, "_native", "liblzf_oracle_native.so")
_native = None


def native_lib():
    """The same sources built with -march=native ON THIS MACHINE (bench.py's CPU-baseline legs): oracle/_native is
    never shipped between machines (.gpurunignore), so the first call on a box compiles it there."""
    global _native
    if _native is None:
        src = [os.path.join(_DIR, f) for f in ("lzf_oracle.c", "lzf_oracle.h")]
        if not (os.path.exists(_NATIVE_SO) and all(os.path.getmtime(_NATIVE_SO) >= os.path.getmtime(x) for x in src)):
            subprocess.check_call(["make", "-C", _DIR, "-s", "-B", "_native/liblzf_oracle_native.so"])
        _native = C.CDLL(_NATIVE_SO)
        _declare_mt(_native)
    return _native


def _declare_mt(L):
    L.lzfo_compress_blocks_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.lzfo_decompress_blocks_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.lzfo_liblz4_blocks_mt.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]


def build(force=False):
    src = [os.path.join(_DIR, f) for f in ("lzf_oracle.c", "lzf_oracle.h")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src)):
        return _SO
    subprocess.check_call(["make", "-C", _DIR, "-s", "-B", "liblzf_oracle.so"])
    return _SO


class Settings(C.Structure):
    _fields_ = [
        ("independent_blocks", C.c_int), ("block_checksums", C.c_int), ("content_checksum", C.c_int),
        ("block_size", C.c_uint64), ("dictionary", C.c_void_p), ("dictionary_len", C.c_uint64),
        ("has_dictionary_id", C.c_int), ("dictionary_id", C.c_uint32),
        ("has_content_size", C.c_int), ("content_size", C.c_uint64), ("hashlog", C.c_uint32),
    ]


class FrameInfo(C.Structure):
    _fields_ = [
        ("flags", C.c_uint8), ("block_maxsize", C.c_uint64), ("has_content_size", C.c_int),
        ("content_size", C.c_uint64), ("has_dictionary_id", C.c_int), ("dictionary_id", C.c_uint32),
        ("header_len", C.c_size_t),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        try:
            _lib = C.CDLL(_SO)
        except OSError:
            build(force=True)
            _lib = C.CDLL(_SO)
        L = _lib
        L.lzfo_xxh32.restype = C.c_uint32
        L.lzfo_xxh32.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
        L.lzfo_compress_bound.restype = C.c_size_t
        L.lzfo_compress_bound.argtypes = [C.c_size_t]
        L.lzfo_compress_block.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint, C.c_void_p, C.c_size_t,
                                          C.POINTER(C.c_size_t)]
        L.lzfo_table_new.restype = C.c_void_p
        L.lzfo_table_new.argtypes = [C.c_int, C.c_uint]
        L.lzfo_table_free.argtypes = [C.c_void_p]
        L.lzfo_table_offset.argtypes = [C.c_void_p, C.c_size_t]
        L.lzfo_compress2.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                     C.POINTER(C.c_size_t)]
        L.lzfo_decompress_raw.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                          C.c_size_t, C.POINTER(C.c_size_t)]
        L.lzfo_settings_default.argtypes = [C.POINTER(Settings)]
        L.lzfo_frame_bound.restype = C.c_size_t
        L.lzfo_frame_bound.argtypes = [C.POINTER(Settings), C.c_size_t]
        L.lzfo_frame_compress.argtypes = [C.POINTER(Settings), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                          C.POINTER(C.c_size_t)]
        L.lzfo_frame_parse_header.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(FrameInfo), C.POINTER(C.c_int)]
        L.lzfo_frame_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                            C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
        _declare_mt(L)
    return _lib


def _buf(b):
    """bytes-like -> (ctypes pointer, length, keepalive)"""
    a = np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b
    a = np.ascontiguousarray(a)
    return a.ctypes.data if a.size else None, a.size, a


def xxh32(data, seed=0):
    p, n, _k = _buf(data)
    return lib().lzfo_xxh32(p, n, seed)


def compress_block(data, table=TABLE_U32, hashlog=12, cap=None):
    """raw::compress2 with a fresh table; cap=None -> unbounded writer (Vec)."""
    p, n, _k = _buf(data)
    if cap is None:
        cap = lib().lzfo_compress_bound(n)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    w = C.c_size_t(0)
    st = lib().lzfo_compress_block(p, n, table, hashlog, out.ctypes.data, cap, C.byref(w))
    return st, out[: w.value].tobytes()


class Table:
    """An EncoderTable that lives across compress2 calls (src/raw/compress/mod.rs:19-101)."""

    def __init__(self, kind=TABLE_U32, hashlog=12):
        self.h = lib().lzfo_table_new(kind, hashlog)

    def offset(self, by):                      # EncoderTable::offset  :72-74
        lib().lzfo_table_offset(self.h, by)

    def __del__(self):
        if getattr(self, "h", None):
            lib().lzfo_table_free(self.h)
            self.h = None


def compress2(data, cursor, table, cap=None):
    """raw::compress2(input, cursor, &mut table, writer): input[..cursor] is match-only history."""
    p, n, _k = _buf(data)
    if cap is None:
        cap = lib().lzfo_compress_bound(n)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    w = C.c_size_t(0)
    st = lib().lzfo_compress2(p, n, cursor, table.h, out.ctypes.data, cap, C.byref(w))
    return st, out[: w.value].tobytes()


def decompress_raw(data, prefix=b"", out_limit=None, cap=None, initial=b""):
    """raw::decompress_raw; `initial` = pre-existing Vec contents. Returns (status, output bytes incl. initial)."""
    p, n, _k = _buf(data)
    pp, pn, _k2 = _buf(prefix)
    if out_limit is None:
        out_limit = (1 << 62)
    if cap is None:
        cap = min(out_limit, 1 << 26) + n + len(initial) + 16
    out = np.empty(max(cap, 1), dtype=np.uint8)
    out[: len(initial)] = np.frombuffer(initial, dtype=np.uint8)
    olen = C.c_size_t(len(initial))
    st = lib().lzfo_decompress_raw(p, n, pp, pn, out.ctypes.data, cap, out_limit, C.byref(olen))
    return st, out[: min(olen.value, cap)].tobytes(), olen.value


def make_settings(independent_blocks=True, block_checksums=False, content_checksum=True,
                  block_size=4 << 20, dictionary=None, dictionary_id=None, content_size=None, hashlog=12):
    s = Settings()
    lib().lzfo_settings_default(C.byref(s))
    s.independent_blocks = int(independent_blocks)
    s.block_checksums = int(block_checksums)
    s.content_checksum = int(content_checksum)
    s.block_size = block_size
    keep = None
    if dictionary is not None:
        keep = np.frombuffer(bytes(dictionary), dtype=np.uint8).copy() if len(dictionary) else np.zeros(1, np.uint8)
        s.dictionary = keep.ctypes.data
        s.dictionary_len = len(dictionary)
    if dictionary_id is not None:
        s.has_dictionary_id = 1
        s.dictionary_id = dictionary_id
    if content_size is not None:
        s.has_content_size = 1
        s.content_size = content_size
    s.hashlog = hashlog
    return s, keep


def frame_compress(data, **kw):
    s, _keep = make_settings(**kw)
    p, n, _k = _buf(data)
    cap = lib().lzfo_frame_bound(C.byref(s), n)
    out = np.empty(cap, dtype=np.uint8)
    w = C.c_size_t(0)
    rc = lib().lzfo_frame_compress(C.byref(s), p, n, out.ctypes.data, cap, C.byref(w))
    return rc, out[: w.value].tobytes()


def frame_decompress(data, dictionary=b"", cap=None):
    """Returns (rc, detail, plaintext, consumed)."""
    p, n, _k = _buf(data)
    dp, dn, _k2 = _buf(dictionary)
    if cap is None:
        cap = max(1 << 20, 300 * n + (4 << 20))
    out = np.empty(cap, dtype=np.uint8)
    w = C.c_size_t(0)
    c = C.c_size_t(0)
    d = C.c_int(0)
    rc = lib().lzfo_frame_decompress(p, n, dp, dn, out.ctypes.data, cap, C.byref(w), C.byref(c), C.byref(d))
    return rc, d.value, out[: w.value].tobytes(), c.value


def parse_header(data):
    """LZ4FrameReader::new -> (rc, detail, FrameInfo)"""
    p, n, _k = _buf(data)
    info = FrameInfo()
    d = C.c_int(0)
    rc = lib().lzfo_frame_parse_header(p, n, C.byref(info), C.byref(d))
    return rc, d.value, info


def compress_blocks_mt(inp, in_off, in_len, out, out_off, hashlog=12, nthreads=1, native=False):
    nb = len(in_len)
    out_len = np.zeros(nb, dtype=np.uint32)
    status = np.zeros(nb, dtype=np.int32)
    (native_lib() if native else lib()).lzfo_compress_blocks_mt(inp.ctypes.data, in_off.ctypes.data, in_len.ctypes.data, nb, hashlog,
                                  out.ctypes.data, out_off.ctypes.data, out_len.ctypes.data, status.ctypes.data,
                                  nthreads)
    return out_len, status


def decompress_blocks_mt(inp, in_off, in_len, out, out_off, out_cap, out_limit, nthreads=1, native=False):
    nb = len(in_len)
    out_len = np.zeros(nb, dtype=np.uint32)
    status = np.zeros(nb, dtype=np.int32)
    (native_lib() if native else lib()).lzfo_decompress_blocks_mt(inp.ctypes.data, in_off.ctypes.data, in_len.ctypes.data, nb, out.ctypes.data,
                                    out_off.ctypes.data, out_cap.ctypes.data, out_limit.ctypes.data,
                                    out_len.ctypes.data, status.ctypes.data, nthreads)
    return out_len, status


def liblz4_blocks_mt(compress, inp, in_off, in_len, out, out_off, out_cap, nthreads=1, native=True):
    """C lz4 (liblz4.so.1) over the same blocks on the same thread pool -> (out_len, status), or None when the
    library is not installed."""
    nb = len(in_len)
    out_len = np.zeros(nb, dtype=np.uint32)
    status = np.zeros(nb, dtype=np.int32)
    rc = (native_lib() if native else lib()).lzfo_liblz4_blocks_mt(
        1 if compress else 0, inp.ctypes.data, in_off.ctypes.data, in_len.ctypes.data, nb, out.ctypes.data,
        out_off.ctypes.data, out_cap.ctypes.data, out_len.ctypes.data, status.ctypes.data, nthreads)
    return None if rc != 0 else (out_len, status)
