"""`lz_fear::raw` — the block codec surface of the reference (src/raw/mod.rs:12-16), served by the
sm_100a kernels through the C ABI.

    compress2(input, cursor, table, writer)        src/raw/compress/mod.rs:165-238
    compress_into(input, out) -> int               (BASELINE north_star name; = compress2 into NoPartialWrites(out))
    decompress_raw(input, prefix, output, limit)   src/raw/decompress.rs:58-78
    DecodeError + 4 variants                       src/raw/decompress.rs:7-17
    EncoderTable / U32Table / U16Table             src/raw/compress/mod.rs:19-101
"""
from . import _native as N

_default_ctx = None


def default_context():
    """Process-wide context on cuda:0 (created on first use; raises without the CUDA library / a GPU)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = N.Context(0)
    return _default_ctx


def set_default_context(ctx):
    global _default_ctx
    _default_ctx = ctx


class DecodeError(Exception):
    """raw::DecodeError (src/raw/decompress.rs:7-17)."""
    code = 0


class UnexpectedEnd(DecodeError):
    code = N.UNEXPECTED_END


class MemoryLimitExceeded(DecodeError):
    code = N.MEMORY_LIMIT_EXCEEDED


class ZeroDeduplicationOffset(DecodeError):
    code = N.ZERO_DEDUP_OFFSET


class InvalidDeduplicationOffset(DecodeError):
    code = N.INVALID_DEDUP_OFFSET


_DECODE_ERRORS = {e.code: e for e in (UnexpectedEnd, MemoryLimitExceeded, ZeroDeduplicationOffset,
                                      InvalidDeduplicationOffset)}


def decode_error_from_status(status):
    return _DECODE_ERRORS[status]()


class WriterFull(OSError):
    """io::ErrorKind::ConnectionAborted out of NoPartialWrites (src/framed/compress.rs:298-301)."""


class EncoderTable:
    """EncoderTable (src/raw/compress/mod.rs:19-25).  While a table is fresh (Default::default(), never used with a
    history) it only selects the flavour: the kernel zeroes its own copy.  The first call that carries state — cursor > 0,
    offset(), or a second compress2 with the same table — makes it a device-resident object (lzf_table_*) that lives
    across calls with the reference's replace / offset semantics."""
    kind = N.TABLE_U32
    _LIMIT = 0xFFFFFFFF

    def __init__(self, hashlog=12):
        self.hashlog = hashlog
        self._fresh = True
        self._handle = None
        self._ctx = None
        self._pending_offset = 0

    @classmethod
    def payload_size_limit(cls):
        return cls._LIMIT

    def offset(self, offset):
        """EncoderTable::offset (:72-74,97-99): positions of later calls are shifted by `offset`."""
        self._pending_offset += int(offset)
        self._fresh = False

    def _device(self, ctx):
        if self._handle is None:
            self._ctx = ctx
            self._handle = ctx.table_new(self.kind, self.hashlog)
        if self._pending_offset:
            ctx.table_offset(self._handle, self._pending_offset)
            self._pending_offset = 0
        return self._handle

    def __del__(self):
        if getattr(self, "_handle", None) is not None and self._ctx is not None:
            try:
                self._ctx.table_free(self._handle)
            except Exception:            # noqa: BLE001 — interpreter shutdown
                pass
            self._handle = None


class U32Table(EncoderTable):          # src/raw/compress/mod.rs:27-36,63-76
    kind = N.TABLE_U32
    _LIMIT = 0xFFFFFFFF


class U16Table(EncoderTable):          # src/raw/compress/mod.rs:78-101
    kind = N.TABLE_U16
    _LIMIT = 0xFFFF


def _compress(input, table, cap, ctx, cursor=0):
    ctx = ctx or default_context()
    table = table or U32Table()
    if len(input) > table.payload_size_limit():
        raise AssertionError("assertion failed: input.len() <= T::payload_size_limit()")   # mod.rs:167
    if cursor == 0 and table._fresh:
        status, out = ctx.raw_compress_into(input, cap=cap, table=table.kind, hashlog=table.hashlog)
    else:
        status, out = ctx.raw_compress2(input, cursor, table._device(ctx), cap=cap)
    if status == N.PANIC:
        raise AssertionError("EncoderTable contract violated")
    return status, out


def compress2(input, cursor, table, writer, ctx=None):
    """raw::compress2(input, cursor, &mut table, writer) (src/raw/compress/mod.rs:165-238): input[..cursor] is match-only
    history, the table keeps its entries from call to call (dependent blocks, src/framed/compress.rs:220,270-275).
    `writer` needs a .write(bytes) method; an unbounded writer never refuses."""
    if table is not None and table._fresh and cursor == 0 and len(input):
        # first use of a fresh table at cursor 0: the kernel's own zeroed table would do, but the caller may come back
        # with the same table, so the state has to be kept from the start
        table._fresh = False
    status, out = _compress(input, table, None, ctx, cursor)
    assert status == N.OK
    writer.write(out)


def compress_into(input, out, table=None, ctx=None):
    """compress2 through NoPartialWrites(out): returns the number of bytes written into `out`
    (bytearray / memoryview) or raises WriterFull when the block does not fit."""
    status, data = _compress(input, table, len(out), ctx)
    if status == N.WRITER_FULL:
        raise WriterFull("ConnectionAborted")
    assert status == N.OK
    out[: len(data)] = data
    return len(data)


def decompress_raw(input, prefix, output, output_limit, ctx=None):
    """raw::decompress_raw: appends to `output` (bytearray); bytes already in it are history."""
    ctx = ctx or default_context()
    existing = bytes(output)
    hist = bytes(prefix) + existing if existing else bytes(prefix)
    limit = max(0, int(output_limit) - len(existing))
    limit = min(limit, 0xFFFFFFFF)
    status, data, n = ctx.raw_decompress(input, prefix=hist, out_limit=limit, cap=limit + len(input) + 16)
    if status in _DECODE_ERRORS:
        raise decode_error_from_status(status)
    assert status == N.OK and n == len(data)
    output += data
