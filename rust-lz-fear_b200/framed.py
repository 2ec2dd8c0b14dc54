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

    # bytes of plaintext handed to the GPU per launch of the streaming path (whole blocks; >= one block)
    STREAM_CHUNK_BYTES = 1 << 30

    def _compress_internal(self, reader, writer, content_size):
        """compress_internal (compress.rs:159-282).  Independent blocks without a dictionary are STREAMED: the reader is
        consumed in chunks of whole blocks, every chunk is one batched launch, and the next chunk is read and compressed
        on a second thread while this one's bytes are written.  The block records of a chunk compressed on its own are
        the records the whole frame would hold (independent blocks share nothing, compress.rs:265-270); the content
        checksum runs over the plaintext as it streams by (compress.rs:172,233-235).  Dependent blocks and dictionaries
        carry a table from block to block: they go through the one-shot call."""
        ctx = self._ctx or _raw.default_context()
        s, keep = self._settings(content_size)
        streamable = self._independent_blocks and not self._dictionary and self._block_size in (64 << 10, 256 << 10, 1 << 20, 4 << 20)
        chunk_bytes = max(self._block_size, self.STREAM_CHUNK_BYTES // self._block_size * self._block_size) if streamable else -1
        try:
            first = _read_up_to(reader, chunk_bytes)
        except OSError as e:
            raise ReadError(str(e)) from e
        if not streamable or len(first) < chunk_bytes:
            status, frame = ctx.frame_compress(first, settings=s)       # everything fits one call
            del keep
            _raise_frame_status(status)
            try:
                writer.write(frame)
            except OSError as e:
                raise WriteError(str(e)) from e
            return
        self._compress_streaming(ctx, reader, writer, first, content_size, chunk_bytes)

    def _compress_streaming(self, ctx, reader, writer, first, content_size, chunk_bytes):
        import threading
        import queue
        # the header as compress_internal writes it (:163-200): taken from an empty frame with the same settings
        s_full, keep_full = self._settings(content_size)
        status, empty = ctx.frame_compress(b"", settings=s_full)
        _raise_frame_status(status)
        hlen = len(empty) - 4 - (4 if self._content_checksum else 0)
        header = empty[:hlen]
        # chunks are compressed as frames of their own without a content checksum: header | records | EndMark
        s_chunk, keep_chunk = N.make_settings(independent_blocks=True, block_checksums=self._block_checksums, content_checksum=False,
                                              block_size=self._block_size)
        chunk_hlen = 7
        hasher = ctx.xxh32_new() if self._content_checksum else None
        q = queue.Queue(maxsize=2)
        stop = threading.Event()            # set when the consumer gives up (a failing writer): the producer must not block
        worker_ctx = N.Context(ctx.device) if hasattr(ctx, "device") else ctx

        def put(item):
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.05)
                    return True
                except queue.Full:
                    pass
            return False

        def produce():
            try:
                chunk = first
                while chunk:
                    st, fr = worker_ctx.frame_compress(chunk, settings=s_chunk)
                    if hasher is not None and st == N.F_OK:
                        worker_ctx.xxh32_update(hasher, chunk)
                    if not put((st, fr)) or st != N.F_OK:
                        return
                    # a reader may return short counts (pipes, sockets): a chunk is cut at EOF only, never inside a block
                    chunk = _read_up_to(reader, chunk_bytes)
                put(None)
            except OSError as e:
                put(ReadError(str(e)))
            except Exception as e:          # noqa: BLE001 — handed to the consumer
                put(e)

        t = threading.Thread(target=produce, daemon=True)
        t.start()
        try:
            writer.write(header)
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, Exception):
                    raise item
                st, fr = item
                _raise_frame_status(st)
                writer.write(memoryview(fr)[chunk_hlen:len(fr) - 4])   # the block records between the chunk's header and EndMark
            writer.write((0).to_bytes(4, "little"))                     # EndMark :277
            if hasher is not None:
                writer.write(ctx.xxh32_finish(hasher).to_bytes(4, "little"))   # :279-281
        except OSError as e:
            raise WriteError(str(e)) from e
        finally:
            stop.set()
            t.join()
            if worker_ctx is not ctx:
                worker_ctx.close()
            del keep_full, keep_chunk

    def compress_to_bytes(self, data):
        out = io.BytesIO()
        self.compress(io.BytesIO(bytes(data)), out)
        return out.getvalue()


def _read_up_to(reader, n):
    """Read::read until `n` bytes are there or the stream ends (n < 0: everything) — what the reference's block loop does
    with its own buffer (compress.rs:222-231: read until the block is full or read() returns 0)."""
    if n < 0:
        return reader.read()
    first = reader.read(n)
    if len(first) == n or not first:
        return first
    parts, total = [first], len(first)
    while total < n:
        more = reader.read(n - total)
        if not more:
            break
        parts.append(more)
        total += len(more)
    return b"".join(parts)


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


class _ReadAhead:
    """The streaming half of LZ4FrameIoReader: a second thread reads the compressed stream AHEAD of the caller, cuts it
    at block boundaries (the length-word chase of decompress.rs:205-226), and decodes each batch of whole blocks with ONE
    batched launch while the caller is still consuming the previous batch.  A batch travels as a frame of its own —
    the original FLG/BD without the content-checksum flag, the block records as they are, an EndMark — so block
    checksums, stored blocks, dependent blocks (their window = the `dictionary` of the batch frame, decompress.rs:
    238-269) and every error keep the semantics of the frame decoder.  The content checksum runs over the plaintext
    as it streams by (decompress.rs:276-278,207-211)."""

    def __init__(self, fr, dictionary, batch_bytes):
        import queue
        import threading
        self.fr = fr
        self.q = queue.Queue(maxsize=2)
        self._queue_full = queue.Full
        self.stop = threading.Event()                        # set by close(): a reader dropped mid-frame must not strand the thread
        self.batch_bytes = batch_bytes
        base = fr._context()
        self.ctx = N.Context(base.device) if hasattr(base, "device") else base
        self.own_ctx = self.ctx is not base
        flg = 0x40 | (fr.flags & 0x30)                       # version 1, independence and block-checksum bits only
        bd = {64 << 10: 0x40, 256 << 10: 0x50, 1 << 20: 0x60, 4 << 20: 0x70}[fr.block_maxsize]
        st = self.ctx.xxh32_new()
        self.ctx.xxh32_update(st, bytes([flg, bd]))          # < 16 bytes: no launch, the stripes stay on the host side
        self.header = (MAGIC).to_bytes(4, "little") + bytes([flg, bd, (self.ctx.xxh32_finish(st) >> 8) & 0xFF])
        self.window = bytes(dictionary)                      # history in front of the next batch
        self.dependent = not (fr.flags & 0x20)
        self.thread = threading.Thread(target=self._produce, daemon=True)
        self.thread.start()

    def _put(self, item):
        while not self.stop.is_set():
            try:
                self.q.put(item, timeout=0.05)
                return True
            except self._queue_full:
                pass
        return False

    def _read_batch(self, carried):
        """Whole block records up to batch_bytes -> (records, number of blocks, tail); tail = "more" | ("end", content
        checksum or None) | "overflow" (a length word beyond block_maxsize, decompress.rs:220-222) | ("error", exception)."""
        r, fr = self.fr.reader, self.fr
        parts, total, nblocks = ([carried] if carried else []), len(carried), 0
        try:
            while total < self.batch_bytes:
                word = _read_exact(r, 4)
                n = int.from_bytes(word, "little")
                if n == 0:
                    cks = int.from_bytes(_read_exact(r, 4), "little") if fr.flags & 0x04 else None
                    return b"".join(parts), nblocks, ("end", cks)
                n &= ~INCOMPRESSIBLE
                if n > fr.block_maxsize:
                    parts.append(word)                       # the frame decoder reports BlockSizeOverflow at this record
                    return b"".join(parts), nblocks + 1, "overflow"
                body = _read_exact(r, n + (4 if fr.flags & 0x10 else 0))
                parts += [word, body]
                total += 4 + len(body)
                nblocks += 1
        except InputError as e:
            return b"".join(parts), nblocks, ("error", e)    # the whole blocks in front of the cut are still delivered
        return b"".join(parts), nblocks, "more"

    def _count_blocks(self, records):
        n, p, extra = 0, 0, 4 if self.fr.flags & 0x10 else 0
        while p + 4 <= len(records):
            p += 4 + (int.from_bytes(records[p:p + 4], "little") & ~INCOMPRESSIBLE) + extra
            n += 1
        return n

    def _produce(self):
        fr = self.fr
        carried = b""
        try:
            while not self.stop.is_set():
                records, nblocks, tail = self._read_batch(carried)
                if carried:
                    nblocks += self._count_blocks(carried)
                carried = b""
                if records:
                    frame = b"".join((self.header, records, b"" if tail == "overflow" else (0).to_bytes(4, "little")))
                    status, detail, plain, consumed = self.ctx.frame_decompress(frame, dictionary=self.window,
                                                                                cap=max(1, nblocks) * fr.block_maxsize, view=True)
                    if len(plain):
                        if fr.content_hasher is not None:
                            self.ctx.xxh32_update(fr.content_hasher, plain)
                        if self.dependent:
                            self.window = (self.window + plain[-WINDOW_SIZE:].tobytes())[-WINDOW_SIZE:]
                        self._put(("data", memoryview(plain)))     # the batch's own output buffer: no copy on the way to the caller
                    if status != N.F_OK:
                        self._put(("status", (status, detail)))
                        return
                    if consumed < len(frame) - 4:
                        # a block that decoded to nothing: the reader hands an empty buffer to its caller (read_to_end
                        # stops there, decompress.rs:54-61) and goes on with the next block if it is asked again
                        self._put(("empty", None))
                        carried = frame[consumed:len(frame) - 4]
                        if tail == "more":
                            continue
                        # the stream's tail is already known: decode the rest before acting on it
                        while carried:
                            fr2 = self.header + carried + (0).to_bytes(4, "little")
                            status, detail, plain, consumed = self.ctx.frame_decompress(
                                fr2, dictionary=self.window, cap=max(1, self._count_blocks(carried)) * fr.block_maxsize, view=True)
                            if len(plain):
                                if fr.content_hasher is not None:
                                    self.ctx.xxh32_update(fr.content_hasher, plain)
                                if self.dependent:
                                    self.window = (self.window + plain[-WINDOW_SIZE:].tobytes())[-WINDOW_SIZE:]
                                self._put(("data", memoryview(plain)))
                            if status != N.F_OK:
                                self._put(("status", (status, detail)))
                                return
                            if consumed < len(fr2) - 4:
                                self._put(("empty", None))
                                carried = fr2[consumed:len(fr2) - 4]
                            else:
                                carried = b""
                if tail == "more":
                    continue
                if tail == "overflow":
                    self._put(("error", BlockSizeOverflow()))
                    return
                if tail[0] == "error":
                    self._put(("error", tail[1]))
                    return
                cks = tail[1]                                # EndMark: verify the content checksum (:206-215)
                if fr.content_hasher is not None and cks is not None:
                    hasher, fr.content_hasher = fr.content_hasher, None
                    if self.ctx.xxh32_finish(hasher) != cks:
                        self._put(("error", FrameChecksumFail()))
                        return
                self._put(("end", None))
                return
        except Exception as e:           # noqa: BLE001 — handed to the consumer
            self._put(("error", e))

    def next(self):
        """-> (plaintext of the next batch, frame finished?).  (b"", False) = a block that decoded to nothing; raises what
        the frame decoder would."""
        kind, val = self.q.get()
        if kind == "data":
            return val, False
        if kind == "empty":
            return b"", False
        self.close()
        if kind == "end":
            self.fr.finished = True
            return b"", True
        if kind == "status":
            _raise_frame_status(*val)
        raise val

    def close(self):
        self.stop.set()
        self.thread.join(timeout=60)
        if self.own_ctx and not self.thread.is_alive():
            self.ctx.close()
            self.own_ctx = False


class LZ4FrameIoReader:
    """LZ4FrameIoReader (src/framed/decompress.rs:46-77): Read + BufRead over a frame.  With `read_ahead` (default) the
    compressed stream is read and decoded in batches of whole blocks ahead of the caller (see _ReadAhead); read_ahead=0
    keeps the reference's one decode_block() per fill_buf()."""

    READ_AHEAD_BYTES = 256 << 20

    def __init__(self, frame_reader, dictionary, read_ahead=None):
        self.frame_reader = frame_reader
        self.bytes_taken = 0
        self.buffer = bytearray()           # (read-ahead: the current batch itself, a read-only view; else one decoded block)
        self.dictionary = dictionary
        self.read_ahead = self.READ_AHEAD_BYTES if read_ahead is None else int(read_ahead)
        self._ahead = None
        self._done = False

    def fill_buf(self):
        """BufRead::fill_buf: the unconsumed part of the current buffer (a zero-copy view), refilled when it is empty."""
        if self.bytes_taken == len(self.buffer):
            self.bytes_taken = 0
            if self.read_ahead > 0 and self.frame_reader.block_maxsize in (64 << 10, 256 << 10, 1 << 20, 4 << 20):
                self.buffer = b""
                if self._done or self.frame_reader.finished:
                    return b""
                if self._ahead is None:
                    self._ahead = _ReadAhead(self.frame_reader, self.dictionary, self.read_ahead)
                try:
                    chunk, finished = self._ahead.next()
                except Exception:
                    self._done = True
                    raise
                if finished:
                    self._done = True
                self.buffer = chunk                      # bytes or a memoryview over the batch's output buffer
            else:
                self.buffer = bytearray()
                self.frame_reader.decode_block(self.buffer, self.dictionary)
        return memoryview(self.buffer)[self.bytes_taken:]

    def close(self):
        """Stops the read-ahead thread of a reader that is dropped before the end of its frame (idempotent; also run by
        `with` and on garbage collection)."""
        if self._ahead is not None:
            self._ahead.close()
        self._done = True

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:                # noqa: BLE001 — interpreter shutdown
            pass

    def consume(self, amt):
        self.bytes_taken += amt
        assert self.bytes_taken <= len(self.buffer), "You consumed more bytes than I even gave you!"

    def read(self, n=-1):
        if n is None or n < 0:
            return self.read_to_end()
        mybuf = self.fill_buf()
        take = min(len(mybuf), n)
        self.consume(take)
        return bytes(mybuf[:take])

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
