"""ctypes binding of liblzfear_b200.so (include/lzfear_b200.h).

There is no CPU implementation behind this module: if the CUDA shared library has not been built,
or no CUDA device is present, everything here raises.  `load_library(path)` exists so that the
test-suite can point the very same binding at another build of the same C ABI.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DEFAULT_LIB = os.path.join(_HERE, "liblzfear_b200.so")

# call-level codes
SUCCESS, ERR_INVALID_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_OOM, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
# per-block codec status (raw::DecodeError order, src/raw/decompress.rs:7-17)
OK, UNEXPECTED_END, MEMORY_LIMIT_EXCEEDED, ZERO_DEDUP_OFFSET, INVALID_DEDUP_OFFSET, WRITER_FULL, OUTPUT_CAP, PANIC = range(8)
# frame-level status
F_OK = 0
F_INPUT_ERROR, F_CODEC_ERROR, F_HEADER_PARSE_ERROR, F_WRONG_MAGIC, F_HEADER_CHECKSUM_FAIL = 10, 11, 12, 13, 14
F_BLOCK_CHECKSUM_FAIL, F_FRAME_CHECKSUM_FAIL, F_BLOCK_LENGTH_OVERFLOW, F_BLOCK_SIZE_OVERFLOW = 15, 16, 17, 18
F_INVALID_BLOCK_SIZE, F_WRITE_ERROR, F_PANIC = 20, 21, 22
P_UNIMPLEMENTED_BLOCKSIZE, P_UNSUPPORTED_VERSION, P_RESERVED_FLAG_BITS, P_RESERVED_BD_BITS = 1, 2, 3, 4
TABLE_U32, TABLE_U16 = 0, 1
INCOMPRESSIBLE = 0x80000000
ABI_VERSION = 3
OPT_SEGMENT_BYTES = 1


class NativeLibraryError(RuntimeError):
    """liblzfear_b200.so is missing / not loadable, or there is no CUDA device."""


class LzfCallError(RuntimeError):
    def __init__(self, code, what):
        super().__init__("lzf call failed (%d): %s" % (code, what))
        self.code = code


class Settings(C.Structure):
    _fields_ = [
        ("independent_blocks", C.c_int32), ("block_checksums", C.c_int32), ("content_checksum", C.c_int32),
        ("block_size", C.c_uint64), ("dictionary", C.c_void_p), ("dictionary_len", C.c_uint64),
        ("has_dictionary_id", C.c_int32), ("dictionary_id", C.c_uint32),
        ("has_content_size", C.c_int32), ("content_size", C.c_uint64), ("hashlog", C.c_uint32),
    ]


class FrameInfo(C.Structure):
    _fields_ = [
        ("flags", C.c_uint8), ("block_maxsize", C.c_uint64), ("has_content_size", C.c_int32),
        ("content_size", C.c_uint64), ("has_dictionary_id", C.c_int32), ("dictionary_id", C.c_uint32),
        ("header_len", C.c_size_t),
    ]


class Xxh32State(C.Structure):
    _fields_ = [("acc", C.c_uint32 * 4), ("buf", C.c_uint8 * 16), ("buflen", C.c_uint32), ("total", C.c_uint64)]


_P = C.c_void_p
_PROTOTYPES = {
    "lzf_xxh32_init": (None, [C.POINTER(Xxh32State)]),
    "lzf_xxh32_update": (C.c_int, [_P, C.POINTER(Xxh32State), _P, C.c_size_t]),
    "lzf_xxh32_finish": (C.c_uint32, [C.POINTER(Xxh32State)]),
    "lzf_abi_version": (C.c_int, []),
    "lzf_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "lzf_destroy": (None, [_P]),
    "lzf_last_error": (C.c_char_p, [_P]),
    "lzf_launch_count": (C.c_uint64, [_P]),
    "lzf_set_option": (C.c_int, [_P, C.c_int, C.c_uint64]),
    "lzf_trim": (C.c_int, [_P]),
    "lzf_compress_blocks": (C.c_int, [_P, _P, _P, _P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "lzf_decompress_blocks": (C.c_int, [_P, _P, _P, _P, C.c_uint32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "lzf_xxh32_ranges": (C.c_int, [_P, _P, _P, _P, C.c_uint32, _P, _P]),
    "lzf_raw_compress_into": (C.c_int, [_P, _P, C.c_size_t, C.c_uint32, C.c_uint32, _P, C.c_size_t,
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_int32)]),
    "lzf_raw_decompress": (C.c_int, [_P, _P, C.c_size_t, _P, C.c_size_t, _P, C.c_size_t, C.c_size_t,
                                     C.POINTER(C.c_size_t), C.POINTER(C.c_int32)]),
    "lzf_table_create": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.POINTER(_P)]),
    "lzf_table_destroy": (None, [_P, _P]),
    "lzf_table_reset": (C.c_int, [_P, _P]),
    "lzf_table_offset": (C.c_int, [_P, _P, C.c_uint64]),
    "lzf_raw_compress2": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t, _P, _P, C.c_size_t,
                                    C.POINTER(C.c_size_t), C.POINTER(C.c_int32)]),
    "lzf_compress_bound": (C.c_size_t, [C.c_size_t]),
    "lzf_settings_default": (None, [C.POINTER(Settings)]),
    "lzf_frame_bound": (C.c_size_t, [C.POINTER(Settings), C.c_size_t]),
    "lzf_frame_compress": (C.c_int, [_P, C.POINTER(Settings), _P, C.c_size_t, _P, C.c_size_t,
                                     C.POINTER(C.c_size_t), C.POINTER(C.c_int32)]),
    "lzf_frames_compress": (C.c_int, [_P, C.POINTER(Settings), _P, _P, _P, C.c_uint32, _P, _P, _P, _P, _P]),
    "lzf_frames_compress_device": (C.c_int, [_P, C.POINTER(Settings), _P, _P, _P, C.c_uint32, _P, _P, _P, _P, _P]),
    "lzf_frame_parse_header": (C.c_int, [_P, C.c_size_t, C.POINTER(FrameInfo), C.POINTER(C.c_int32)]),
    "lzf_frame_decompress": (C.c_int, [_P, _P, C.c_size_t, _P, C.c_size_t, _P, C.c_size_t, C.POINTER(C.c_size_t),
                                       C.POINTER(C.c_size_t), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "lzf_frames_decompress": (C.c_int, [_P, _P, _P, _P, C.c_uint32, _P, _P, _P, _P, _P, _P]),
    "lzf_frames_decompress_device": (C.c_int, [_P, _P, _P, _P, C.c_uint32, _P, _P, _P, _P, _P, _P]),
}
EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None
_lib_path = None


def library_path():
    return _lib_path or _DEFAULT_LIB


def load_library(path=None):
    """Loads (once) the C-ABI shared library; raises NativeLibraryError when it is not there."""
    global _lib, _lib_path
    if path is None and _lib is not None:
        return _lib
    # LZF_B200_LIB: explicit opt-in to another build of the SAME CUDA library (kernel tuning variants)
    path = path or os.environ.get("LZF_B200_LIB") or _DEFAULT_LIB
    if not os.path.exists(path):
        raise NativeLibraryError(
            "%s not found: build it with `python rust-lz-fear_b200/build.py` (there is no CPU fallback)" % path)
    try:
        lib = C.CDLL(path)
    except OSError as e:
        raise NativeLibraryError("cannot load %s: %s" % (path, e)) from e
    for name, (res, args) in _PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NativeLibraryError("%s does not export %s" % (path, name)) from e
        fn.restype = res
        fn.argtypes = args
    if lib.lzf_abi_version() != ABI_VERSION:
        raise NativeLibraryError("ABI version mismatch in %s" % path)
    _lib, _lib_path = lib, path
    return lib


def _np_ptr(a):
    return a.ctypes.data if a is not None and a.size else None


def _as_u8(b):
    if isinstance(b, np.ndarray):
        a = b if b.dtype == np.uint8 else b.view(np.uint8)
        return np.ascontiguousarray(a).reshape(-1)
    return np.frombuffer(b, dtype=np.uint8)


def make_settings(independent_blocks=True, block_checksums=False, content_checksum=True, block_size=4 << 20,
                  dictionary=None, dictionary_id=None, content_size=None, content_size_from_input=False, hashlog=12):
    """-> (Settings, keepalive)"""
    s = Settings()
    load_library().lzf_settings_default(C.byref(s))
    s.independent_blocks = int(bool(independent_blocks))
    s.block_checksums = int(bool(block_checksums))
    s.content_checksum = int(bool(content_checksum))
    s.block_size = int(block_size)
    keep = None
    if dictionary is not None:
        keep = np.frombuffer(bytes(dictionary), dtype=np.uint8).copy() if len(dictionary) else np.zeros(1, np.uint8)
        s.dictionary = keep.ctypes.data
        s.dictionary_len = len(dictionary)
    if dictionary_id is not None:
        s.has_dictionary_id = 1
        s.dictionary_id = int(dictionary_id)
    if content_size_from_input:
        s.has_content_size = 2
    elif content_size is not None:
        s.has_content_size = 1
        s.content_size = int(content_size)
    s.hashlog = int(hashlog)
    return s, keep


class Context:
    """One lzf_ctx: a CUDA device, its streams and scratch.  Not thread-safe; one per thread/stream."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = _P()
        rc = self._lib.lzf_create(int(device), C.byref(h))
        if rc == ERR_NO_DEVICE:
            raise NativeLibraryError("no CUDA device: the lz-fear B200 codec has no CPU fallback")
        if rc != SUCCESS:
            raise LzfCallError(rc, "lzf_create")
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lzf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != SUCCESS:
            raise LzfCallError(rc, (self._lib.lzf_last_error(self._h) or b"").decode("utf-8", "replace"))

    @property
    def launch_count(self):
        return int(self._lib.lzf_launch_count(self._h))

    # ---- single block, host buffers (raw::compress2 / raw::decompress_raw shape) ----------------
    def raw_compress_into(self, data, cap=None, table=TABLE_U32, hashlog=12):
        """-> (status, compressed bytes).  cap=None: unbounded writer (a Vec in the reference)."""
        a = _as_u8(data)
        if cap is None:
            cap = int(self._lib.lzf_compress_bound(a.size))
        out = np.empty(max(int(cap), 1), dtype=np.uint8)
        w = C.c_size_t(0)
        st = C.c_int32(0)
        self._check(self._lib.lzf_raw_compress_into(self._h, _np_ptr(a), a.size, table, hashlog, out.ctypes.data,
                                                    int(cap), C.byref(w), C.byref(st)))
        return st.value, out[: w.value].tobytes()

    def trim(self):
        """Gives the context's grow-only device / pinned scratch back to the driver."""
        self._check(self._lib.lzf_trim(self._h))

    def set_option(self, option, value):
        self._check(self._lib.lzf_set_option(self._h, int(option), int(value)))

    def table_new(self, table=TABLE_U32, hashlog=12):
        """A device-resident EncoderTable (U32Table::default() / U16Table::default())."""
        h = _P()
        self._check(self._lib.lzf_table_create(self._h, table, hashlog, C.byref(h)))
        return h

    def table_free(self, t):
        self._lib.lzf_table_destroy(self._h, t)

    def table_reset(self, t):
        self._check(self._lib.lzf_table_reset(self._h, t))

    def table_offset(self, t, by):
        self._check(self._lib.lzf_table_offset(self._h, t, by))

    def raw_compress2(self, data, cursor, table, cap=None):
        """raw::compress2(input, cursor, &mut table, writer) -> (status, compressed bytes)."""
        a = _as_u8(data)
        if cap is None:
            cap = self._lib.lzf_compress_bound(a.size)
        out = np.empty(max(int(cap), 1), dtype=np.uint8)
        w, st = C.c_size_t(0), C.c_int32(0)
        self._check(self._lib.lzf_raw_compress2(self._h, _np_ptr(a), a.size, int(cursor), table, out.ctypes.data, int(cap),
                                                C.byref(w), C.byref(st)))
        return st.value, out[: w.value].tobytes()

    def raw_decompress(self, data, prefix=b"", out_limit=None, cap=None):
        """-> (status, decoded bytes (truncated to cap), decoded length)."""
        a = _as_u8(data)
        p = _as_u8(prefix)
        if out_limit is None:
            out_limit = 0xFFFFFFFF
        if cap is None:
            cap = min(int(out_limit), 1 << 26) + a.size + 16
        out = np.empty(max(int(cap), 1), dtype=np.uint8)
        n = C.c_size_t(0)
        st = C.c_int32(0)
        self._check(self._lib.lzf_raw_decompress(self._h, _np_ptr(a), a.size, _np_ptr(p), p.size, out.ctypes.data,
                                                 int(cap), int(out_limit), C.byref(n), C.byref(st)))
        return st.value, out[: min(n.value, int(cap))].tobytes(), n.value

    # ---- frames, host buffers ------------------------------------------------------------------
    def frame_bound(self, settings, n):
        return int(self._lib.lzf_frame_bound(C.byref(settings), int(n)))

    def frame_compress(self, data, settings=None, cap=None, **kw):
        """CompressionSettings::compress over a host buffer -> (frame status, frame bytes)."""
        keep = None
        if settings is None:
            settings, keep = make_settings(**kw)
        a = _as_u8(data)
        if cap is None:
            cap = self.frame_bound(settings, a.size)
        out = np.empty(max(int(cap), 1), dtype=np.uint8)
        w = C.c_size_t(0)
        st = C.c_int32(0)
        self._check(self._lib.lzf_frame_compress(self._h, C.byref(settings), _np_ptr(a), a.size, out.ctypes.data,
                                                 int(cap), C.byref(w), C.byref(st)))
        del keep
        return st.value, out[: w.value].tobytes()

    def frame_decompress(self, data, dictionary=b"", cap=None, view=False):
        """decompress_frame over a host buffer -> (status, detail, plaintext, consumed).  view=True: the plaintext comes
        back as a uint8 numpy array over the call's own output buffer (no copy) instead of bytes."""
        a = _as_u8(data)
        d = _as_u8(dictionary)
        if cap is None:
            cap = max(1 << 20, 300 * a.size + (4 << 20))
        out = np.empty(max(int(cap), 1), dtype=np.uint8)
        w, cons = C.c_size_t(0), C.c_size_t(0)
        st, det = C.c_int32(0), C.c_int32(0)
        self._check(self._lib.lzf_frame_decompress(self._h, _np_ptr(a), a.size, _np_ptr(d), d.size, out.ctypes.data,
                                                   int(cap), C.byref(w), C.byref(cons), C.byref(st), C.byref(det)))
        return st.value, det.value, (out[: w.value] if view else out[: w.value].tobytes()), cons.value

    def frames_compress(self, inp, in_off, in_len, out, out_off, out_cap, settings):
        """Batched host-buffer frame compress. numpy arrays; returns (out_len u64[], status i32[])."""
        nf = len(in_len)
        in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
        in_len = np.ascontiguousarray(in_len, dtype=np.uint64)
        out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
        out_cap = np.ascontiguousarray(out_cap, dtype=np.uint64)
        out_len = np.zeros(nf, dtype=np.uint64)
        status = np.zeros(nf, dtype=np.int32)
        self._check(self._lib.lzf_frames_compress(self._h, C.byref(settings), _ptr(inp), _np_ptr(in_off), _np_ptr(in_len),
                                                  nf, _ptr(out), _np_ptr(out_off), _np_ptr(out_cap), _np_ptr(out_len),
                                                  _np_ptr(status)))
        return out_len, status

    def frames_decompress(self, inp, in_off, in_len, out, out_off, out_cap):
        """Batched host-buffer frame decompress; returns (out_len u64[], status i32[], detail i32[])."""
        nf = len(in_len)
        in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
        in_len = np.ascontiguousarray(in_len, dtype=np.uint64)
        out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
        out_cap = np.ascontiguousarray(out_cap, dtype=np.uint64)
        out_len = np.zeros(nf, dtype=np.uint64)
        status = np.zeros(nf, dtype=np.int32)
        detail = np.zeros(nf, dtype=np.int32)
        self._check(self._lib.lzf_frames_decompress(self._h, _ptr(inp), _np_ptr(in_off), _np_ptr(in_len), nf, _ptr(out),
                                                    _np_ptr(out_off), _np_ptr(out_cap), _np_ptr(out_len), _np_ptr(status),
                                                    _np_ptr(detail)))
        return out_len, status, detail

    # ---- frames, device buffers (torch tensors or raw device pointers) --------------------------
    def frames_compress_device(self, d_in, in_off, in_len, d_out, out_off, out_cap, settings):
        nf = len(in_len)
        in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
        in_len = np.ascontiguousarray(in_len, dtype=np.uint64)
        out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
        out_cap = np.ascontiguousarray(out_cap, dtype=np.uint64)
        out_len = np.zeros(nf, dtype=np.uint64)
        status = np.zeros(nf, dtype=np.int32)
        self._check(self._lib.lzf_frames_compress_device(self._h, C.byref(settings), _ptr(d_in), _np_ptr(in_off),
                                                         _np_ptr(in_len), nf, _ptr(d_out), _np_ptr(out_off),
                                                         _np_ptr(out_cap), _np_ptr(out_len), _np_ptr(status)))
        return out_len, status

    def frames_decompress_device(self, d_in, in_off, in_len, d_out, out_off, out_cap):
        nf = len(in_len)
        in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
        in_len = np.ascontiguousarray(in_len, dtype=np.uint64)
        out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
        out_cap = np.ascontiguousarray(out_cap, dtype=np.uint64)
        out_len = np.zeros(nf, dtype=np.uint64)
        status = np.zeros(nf, dtype=np.int32)
        detail = np.zeros(nf, dtype=np.int32)
        self._check(self._lib.lzf_frames_decompress_device(self._h, _ptr(d_in), _np_ptr(in_off), _np_ptr(in_len), nf,
                                                           _ptr(d_out), _np_ptr(out_off), _np_ptr(out_cap),
                                                           _np_ptr(out_len), _np_ptr(status), _np_ptr(detail)))
        return out_len, status, detail

    # ---- batched blocks, device pointers, asynchronous on `stream` ------------------------------
    def compress_blocks(self, d_in, d_in_off, d_in_len, nblocks, d_out, d_out_off, d_out_cap, d_out_len, d_status,
                        d_xxh_plain=None, d_xxh_stored=None, hashlog=12, table=TABLE_U32, stream=0, max_block_len=0):
        self._check(self._lib.lzf_compress_blocks(self._h, _ptr(d_in), _ptr(d_in_off), _ptr(d_in_len), int(nblocks),
                                                  int(hashlog), int(table), int(max_block_len), _ptr(d_out), _ptr(d_out_off),
                                                  _ptr(d_out_cap),
                                                  _ptr(d_out_len), _ptr(d_status), _ptr(d_xxh_plain), _ptr(d_xxh_stored),
                                                  _P(int(stream)) if stream else None))

    def decompress_blocks(self, d_in, d_in_off, d_in_len, nblocks, d_out, d_out_off, d_out_cap, d_out_limit, d_out_len,
                          d_status, d_xxh_plain=None, d_prefix=None, d_prefix_off=None, d_prefix_len=None, stream=0):
        self._check(self._lib.lzf_decompress_blocks(self._h, _ptr(d_in), _ptr(d_in_off), _ptr(d_in_len), int(nblocks),
                                                    _ptr(d_prefix), _ptr(d_prefix_off), _ptr(d_prefix_len), _ptr(d_out),
                                                    _ptr(d_out_off), _ptr(d_out_cap), _ptr(d_out_limit), _ptr(d_out_len),
                                                    _ptr(d_status), _ptr(d_xxh_plain),
                                                    _P(int(stream)) if stream else None))

    # ---- streaming XXH32 over host buffers (twox-hash XxHash32::with_seed(0) shape) -------------
    def xxh32_new(self):
        st = Xxh32State()
        self._lib.lzf_xxh32_init(C.byref(st))
        return st

    def xxh32_update(self, st, data):
        a = _as_u8(data)
        self._check(self._lib.lzf_xxh32_update(self._h, C.byref(st), _np_ptr(a), a.size))

    def xxh32_finish(self, st):
        return int(self._lib.lzf_xxh32_finish(C.byref(st)))

    def xxh32(self, data):
        st = self.xxh32_new()
        self.xxh32_update(st, data)
        return self.xxh32_finish(st)

    def xxh32_ranges(self, d_data, d_off, d_len, nranges, d_hash, stream=0):
        self._check(self._lib.lzf_xxh32_ranges(self._h, _ptr(d_data), _ptr(d_off), _ptr(d_len), int(nranges),
                                               _ptr(d_hash), _P(int(stream)) if stream else None))


def _ptr(x):
    """torch tensor / numpy array / int address / None -> address or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x or None
    if isinstance(x, np.ndarray):
        return x.ctypes.data if x.size else None
    if hasattr(x, "data_ptr"):
        return x.data_ptr() or None
    raise TypeError("expected a tensor, ndarray, address or None, got %r" % type(x))


def parse_frame_header(data):
    """LZ4FrameReader::new over the first bytes of a frame -> (status, detail, FrameInfo).  Host-only."""
    a = _as_u8(data)
    info = FrameInfo()
    det = C.c_int32(0)
    rc = load_library().lzf_frame_parse_header(_np_ptr(a), a.size, C.byref(info), C.byref(det))
    return rc, det.value, info
