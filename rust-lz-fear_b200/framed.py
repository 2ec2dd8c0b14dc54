"""`lz_fear::framed` — the frame-format surface of the reference (src/framed/mod.rs:11-24) over the
B200 codec: CompressionSettings (src/framed/compress.rs:36-157), LZ4FrameReader / LZ4FrameIoReader /
decompress_frame (src/framed/decompress.rs:46-288).  Readers/writers are Python file-like objects.
"""
import io

from . import _native as N
from . import raw as _raw

MAGIC = 0x184D2204            # src/framed/mod.rs:16
INCOMPRESSIBLE = 1 << 31      # src/framed/mod.rs:18
WINDOW_SIZE = 64 * 1024       # src/framed/mod.rs:20


class CompressionError(Exception):
    """CompressionError (src/framed/compress.rs:15-23)."""


class ReadError(CompressionError):
    pass


class WriteError(CompressionError):
    pass


class InvalidBlockSize(CompressionError):
    pass


class DecompressionError(Exception):
    """DecompressionError (src/framed/decompress.rs:16-36)."""


class InputError(DecompressionError):
    pass


class CodecError(DecompressionError):
    def __init__(self, inner):
        super().__init__(type(inner).__name__)
        self.inner = inner


class HeaderParseError(DecompressionError):
    KINDS = {N.P_UNIMPLEMENTED_BLOCKSIZE: "UnimplementedBlocksize", N.P_UNSUPPORTED_VERSION: "UnsupportedVersion",
             N.P_RESERVED_FLAG_BITS: "ReservedFlagBitsSet", N.P_RESERVED_BD_BITS: "ReservedBdBitsSet"}

    def __init__(self, kind):
        super().__init__(self.KINDS.get(kind, str(kind)))
        self.kind = kind


class WrongMagic(DecompressionError):
    pass


class HeaderChecksumFail(DecompressionError):
    pass


class BlockChecksumFail(DecompressionError):
    pass


class FrameChecksumFail(DecompressionError):
    pass


class BlockLengthOverflow(DecompressionError):
    pass


class BlockSizeOverflow(DecompressionError):
    pass


def _raise_frame_status(status, detail=0):
    if status == N.F_OK:
        return
    if status == N.F_CODEC_ERROR:
        raise CodecError(_raw.decode_error_from_status(detail))
    if status == N.F_HEADER_PARSE_ERROR:
        raise HeaderParseError(detail)
    table = {N.F_INPUT_ERROR: InputError, N.F_WRONG_MAGIC: WrongMagic, N.F_HEADER_CHECKSUM_FAIL: HeaderChecksumFail,
             N.F_BLOCK_CHECKSUM_FAIL: BlockChecksumFail, N.F_FRAME_CHECKSUM_FAIL: FrameChecksumFail,
             N.F_BLOCK_LENGTH_OVERFLOW: BlockLengthOverflow, N.F_BLOCK_SIZE_OVERFLOW: BlockSizeOverflow,
             N.F_INVALID_BLOCK_SIZE: InvalidBlockSize, N.F_WRITE_ERROR: WriteError}
    if status == N.F_PANIC:
        raise AssertionError("the reference panics on this input (src/framed/header.rs:55 unwrap)")
    raise table[status]()


class CompressionSettings:
    """Builder with the reference's setters and defaults (src/framed/compress.rs:44-133)."""

    def __init__(self, ctx=None):
        self._independent_blocks = True
        self._block_checksums = False
        self._content_checksum = True
        self._block_size = 4 * 1024 * 1024
        self._dictionary = None
        self._dictionary_id = None
        self._ctx = ctx

    @classmethod
    def default(cls):
        return cls()

    def independent_blocks(self, v):
        self._independent_blocks = bool(v)
        return self

    def block_checksums(self, v):
        self._block_checksums = bool(v)
        return self

    def content_checksum(self, v):
        self._content_checksum = bool(v)
        return self

    def block_size(self, v):
        self._block_size = int(v)
        return self

    def dictionary(self, id, dict):
        self._dictionary_id = int(id)
        self._dictionary = bytes(dict)
        return self

    def dictionary_id_nonsense_override(self, id):
        self._dictionary_id = None if id is None else int(id)
        return self

    # ---- compress / compress_with_size / compress_with_size_unchecked  (compress.rs:137-157)
    def compress(self, reader, writer):
        self._compress_internal(reader, writer, None)

    def compress_with_size_unchecked(self, reader, writer, content_size):
        self._compress_internal(reader, writer, int(content_size))

    def compress_with_size(self, reader, writer):
        start = reader.tell()
        end = reader.seek(0, io.SEEK_END)
        reader.seek(start, io.SEEK_SET)
        self._compress_internal(reader, writer, end - start)

    def _settings(self, content_size):
        return N.make_settings(independent_blocks=self._independent_blocks, block_checksums=self._block_checksums,
                               content_checksum=self._content_checksum, block_size=self._block_size,
                               dictionary=self._dictionary, dictionary_id=self._dictionary_id,
                               content_size=content_size)

    def _compress_internal(self, reader, writer, content_size):
        ctx = self._ctx or _raw.default_context()
        try:
            data = reader.read()
        except OSError as e:
            raise ReadError(str(e)) from e
        s, keep = self._settings(content_size)
        status, frame = ctx.frame_compress(data, settings=s)
        del keep
        _raise_frame_status(status)
        try:
            writer.write(frame)
        except OSError as e:
            raise WriteError(str(e)) from e

    def compress_to_bytes(self, data):
        out = io.BytesIO()
        self.compress(io.BytesIO(bytes(data)), out)
        return out.getvalue()


def _read_exact(reader, n):
    buf = b""
    while len(buf) < n:
        chunk = reader.read(n - len(buf))
        if not chunk:
            raise InputError("failed to fill whole buffer")
        buf += chunk
    return buf


class LZ4FrameReader:
    """LZ4FrameReader (src/framed/decompress.rs:81-279): parses the header on construction, then
    decodes one block per decode_block() call through the GPU block decoder."""

    def __init__(self, reader, ctx=None):
        self._ctx = ctx
        self.reader = reader
        hdr = _read_exact(reader, 4)
        status, detail, info = N.parse_frame_header(hdr)
        if status == N.F_WRONG_MAGIC:
            raise WrongMagic(hex(int.from_bytes(hdr, "little")))
        # feed the header parser one field at a time, like the reference reads from its stream
        while status == N.F_INPUT_ERROR:
            hdr += _read_exact(reader, 1)
            status, detail, info = N.parse_frame_header(hdr)
        _raise_frame_status(status, detail)
        self.flags = info.flags
        self.block_maxsize = int(info.block_maxsize)
        self.content_size = int(info.content_size) if info.has_content_size else None
        self._dictionary_id = int(info.dictionary_id) if info.has_dictionary_id else None
        c = self._context()
        self.content_hasher = c.xxh32_new() if self.flags & 0x04 else None
        self.carryover_window = None if self.flags & 0x20 else bytearray()
        self.finished = False

    def _context(self):
        return self._ctx or _raw.default_context()

    def block_size(self):
        return self.block_maxsize

    def frame_size(self):
        return self.content_size

    def dictionary_id(self):
        return self._dictionary_id

    def into_read_with_dictionary(self, dictionary):
        return LZ4FrameIoReader(self, bytes(dictionary))

    def into_read(self):
        return self.into_read_with_dictionary(b"")

    def decode_block(self, output, dictionary=b""):
        """src/framed/decompress.rs:197-279.  `output` must be an empty bytearray."""
        assert len(output) == 0, "You must pass an empty buffer to this interface."
        if self.finished:
            return
        c = self._context()
        block_length = int.from_bytes(_read_exact(self.reader, 4), "little")
        if block_length == 0:
            if self.content_hasher is not None:
                hasher, self.content_hasher = self.content_hasher, None
                checksum = int.from_bytes(_read_exact(self.reader, 4), "little")
                if c.xxh32_finish(hasher) != checksum:
                    raise FrameChecksumFail()
            self.finished = True
            return
        is_compressed = (block_length & INCOMPRESSIBLE) == 0
        block_length &= ~INCOMPRESSIBLE
        if block_length > self.block_maxsize:
            raise BlockSizeOverflow()
        buf = _read_exact(self.reader, block_length)
        if self.flags & 0x10:
            checksum = int.from_bytes(_read_exact(self.reader, 4), "little")
            if c.xxh32(buf) != checksum:
                raise BlockChecksumFail()
        if self.carryover_window is not None:
            if len(self.carryover_window) == 0:
                self.carryover_window += dictionary
            dec_prefix = bytes(self.carryover_window)
        else:
            dec_prefix = bytes(dictionary)
        if is_compressed:
            try:
                _raw.decompress_raw(buf, dec_prefix, output, self.block_maxsize, ctx=c)
            except _raw.DecodeError as e:
                raise CodecError(e) from e
        else:
            output += buf
        window = self.carryover_window
        if window is not None:
            outlen = len(output)
            if outlen < WINDOW_SIZE:
                surplus = len(window) + outlen - WINDOW_SIZE
                if surplus >= 0:
                    del window[:surplus]
                window += output
            else:
                del window[:]
                window += output[outlen - WINDOW_SIZE:]
            assert len(window) <= WINDOW_SIZE
        if len(output) > self.block_maxsize:
            raise BlockSizeOverflow()
        if self.content_hasher is not None:
            c.xxh32_update(self.content_hasher, output)


class LZ4FrameIoReader:
    """LZ4FrameIoReader (src/framed/decompress.rs:46-77): Read + BufRead over a frame."""

    def __init__(self, frame_reader, dictionary):
        self.frame_reader = frame_reader
        self.bytes_taken = 0
        self.buffer = bytearray()
        self.dictionary = dictionary

    def fill_buf(self):
        if self.bytes_taken == len(self.buffer):
            del self.buffer[:]
            self.frame_reader.decode_block(self.buffer, self.dictionary)
            self.bytes_taken = 0
        return bytes(self.buffer[self.bytes_taken:])

    def consume(self, amt):
        self.bytes_taken += amt
        assert self.bytes_taken <= len(self.buffer), "You consumed more bytes than I even gave you!"

    def read(self, n=-1):
        if n is None or n < 0:
            return self.read_to_end()
        mybuf = self.fill_buf()
        take = min(len(mybuf), n)
        self.consume(take)
        return mybuf[:take]

    def read_to_end(self):
        """std::io::Read::read_to_end: stops at the first read() that returns 0 bytes."""
        out = bytearray()
        while True:
            mybuf = self.fill_buf()
            if len(mybuf) == 0:
                return bytes(out)
            out += mybuf
            self.consume(len(mybuf))


def decompress_frame(reader, ctx=None):
    """decompress_frame (src/framed/decompress.rs:283-288): the whole frame in ONE batched GPU call
    when the reader is fully buffered; same result as LZ4FrameReader(..).into_read().read_to_end()."""
    data = reader.read()
    c = ctx or _raw.default_context()
    status, detail, plain, _consumed = c.frame_decompress(data)
    _raise_frame_status(status, detail)
    return plain
