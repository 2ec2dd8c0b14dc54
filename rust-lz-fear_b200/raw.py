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
    """EncoderTable (src/raw/compress/mod.rs:19-25).  The table itself lives in shared memory on the
    GPU for the duration of one block; host objects only select the flavour and carry `hashlog`."""
    kind = N.TABLE_U32
    _LIMIT = 0xFFFFFFFF

    def __init__(self, hashlog=12):
        self.hashlog = hashlog
        self._fresh = True

    @classmethod
    def payload_size_limit(cls):
        return cls._LIMIT


class U32Table(EncoderTable):          # src/raw/compress/mod.rs:27-36,63-76
    kind = N.TABLE_U32
    _LIMIT = 0xFFFFFFFF


class U16Table(EncoderTable):          # src/raw/compress/mod.rs:78-101
    kind = N.TABLE_U16
    _LIMIT = 0xFFFF


def _compress(input, table, cap, ctx):
    ctx = ctx or default_context()
    table = table or U32Table()
    if len(input) > table.payload_size_limit():
        raise AssertionError("assertion failed: input.len() <= T::payload_size_limit()")   # mod.rs:167
    status, out = ctx.raw_compress_into(input, cap=cap, table=table.kind, hashlog=table.hashlog)
    if status == N.PANIC:
        raise AssertionError("EncoderTable contract violated")
    return status, out


def compress2(input, cursor, table, writer, ctx=None):
    """raw::compress2 with cursor 0 and a fresh table (the independent-block call of the framed path,
    src/framed/compress.rs:242-243,265-270).  `writer` needs a .write(bytes) method."""
    if cursor != 0 or (table is not None and not table._fresh):
        raise NotImplementedError("prefix / carried-over table state (dependent blocks) is not on the GPU path yet")
    status, out = _compress(input, table, None, ctx)
    assert status == N.OK
    if table is not None and len(input):
        table._fresh = False
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
