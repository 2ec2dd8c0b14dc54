//! Safe wrappers over `lz-fear-b200-sys`.  The names, argument meaning and error behaviour are the ones of
//! `lz-fear` 0.2 (reference: /root/reference/src), so that `lz-fear` itself can keep `#![forbid(unsafe_code)]` and only
//! swap the bodies of its two hot call sites behind a cargo feature (see rust/patches/lz-fear-b200-feature.patch):
//!
//!   * `compress2(&in_buffer, window_offset, &mut table, &mut NoPartialWrites(..))`   src/framed/compress.rs:242-243
//!   * `raw::decompress_raw(buf, dec_prefix, output, self.block_maxsize)`             src/framed/decompress.rs:247-248
//!
//! There is no CPU fallback: every call needs a CUDA device and fails with an `io::Error` otherwise.
//! (Written for a toolchain this repository's build image does not have — it is checked against the C header by
//! tests/test_abi.py, not compiled here.)
use lz_fear_b200_sys as sys;
use std::ffi::CStr;
use std::io;
use std::sync::Mutex;

/// One context (CUDA device 0 unless `LZFEAR_B200_DEVICE` says otherwise) shared by the process; calls serialise on it.
struct Ctx(*mut sys::lzf_ctx);
unsafe impl Send for Ctx {}

fn with_ctx<R>(f: impl FnOnce(*mut sys::lzf_ctx) -> R) -> io::Result<R> {
    static CTX: Mutex<Option<Ctx>> = Mutex::new(None);
    let mut guard = CTX.lock().unwrap();
    if guard.is_none() {
        let device = std::env::var("LZFEAR_B200_DEVICE").ok().and_then(|v| v.parse().ok()).unwrap_or(0);
        let mut p: *mut sys::lzf_ctx = std::ptr::null_mut();
        let rc = unsafe { sys::lzf_create(device, &mut p) };
        if rc != sys::LZF_SUCCESS {
            return Err(io::Error::new(io::ErrorKind::Other, format!("lzf_create failed ({}): no CUDA device / driver — there is no CPU fallback", rc)));
        }
        *guard = Some(Ctx(p));
    }
    Ok(f(guard.as_ref().unwrap().0))
}

fn call_error(ctx: *mut sys::lzf_ctx, rc: i32) -> io::Error {
    let msg = unsafe { CStr::from_ptr(sys::lzf_last_error(ctx)) }.to_string_lossy().into_owned();
    io::Error::new(io::ErrorKind::Other, format!("liblzfear_b200 call failed ({}): {}", rc, msg))
}

pub mod raw {
    use super::*;

    /// `raw::DecodeError` (src/raw/decompress.rs:7-17), same variants, same order as the C status codes 1..4.
    #[derive(Debug, thiserror::Error, PartialEq, Eq)]
    pub enum DecodeError {
        #[error("input ended prematurely")]
        UnexpectedEnd,
        #[error("the output would exceed the configured limit")]
        MemoryLimitExceeded,
        #[error("deduplication offset of zero")]
        ZeroDeduplicationOffset,
        #[error("deduplication offset points before the start of the history")]
        InvalidDeduplicationOffset,
    }

    /// Which `EncoderTable` of the reference a device table mirrors (src/raw/compress/mod.rs:27-36,78-101).
    pub trait EncoderTable {
        const KIND: u32;
        fn payload_size_limit() -> usize;
        #[doc(hidden)]
        fn handle(&mut self) -> io::Result<*mut sys::lzf_table>;
        /// `EncoderTable::offset` (:72-74)
        fn offset(&mut self, offset: usize);
    }

    macro_rules! device_table {
        ($name:ident, $kind:expr, $limit:expr) => {
            /// A device-resident table with the reference's semantics; `Default` = a zeroed table.
            pub struct $name { h: *mut sys::lzf_table, pending_offset: u64 }
            unsafe impl Send for $name {}
            impl Default for $name {
                fn default() -> Self { $name { h: std::ptr::null_mut(), pending_offset: 0 } }
            }
            impl EncoderTable for $name {
                const KIND: u32 = $kind;
                fn payload_size_limit() -> usize { $limit }
                fn handle(&mut self) -> io::Result<*mut sys::lzf_table> {
                    if self.h.is_null() {
                        let mut h: *mut sys::lzf_table = std::ptr::null_mut();
                        let rc = with_ctx(|c| unsafe { sys::lzf_table_create(c, $kind, 12, &mut h) })?;
                        if rc != sys::LZF_SUCCESS { return Err(io::Error::new(io::ErrorKind::Other, "lzf_table_create")); }
                        self.h = h;
                    }
                    if self.pending_offset != 0 {
                        let by = std::mem::take(&mut self.pending_offset);
                        let h = self.h;
                        with_ctx(|c| unsafe { sys::lzf_table_offset(c, h, by) })?;
                    }
                    Ok(self.h)
                }
                fn offset(&mut self, offset: usize) { self.pending_offset += offset as u64; }
            }
            impl Drop for $name {
                fn drop(&mut self) {
                    if !self.h.is_null() {
                        let h = self.h;
                        let _ = with_ctx(|c| unsafe { sys::lzf_table_destroy(c, h) });
                    }
                }
            }
        };
    }
    device_table!(U32Table, sys::LZF_TABLE_U32 as u32, std::u32::MAX as usize);
    device_table!(U16Table, sys::LZF_TABLE_U16 as u32, std::u16::MAX as usize);

    /// `compress2(input, cursor, &mut table, NoPartialWrites(out))` (src/raw/compress/mod.rs:165-238 with the bounded
    /// writer of src/framed/compress.rs:294-308): `input[..cursor]` is match-only history, the table keeps its entries.
    /// `Err(ConnectionAborted)` = the writer refused a write: the caller stores the block raw (compress.rs:250-255).
    pub fn compress2<T: EncoderTable>(input: &[u8], cursor: usize, table: &mut T, out: &mut [u8]) -> io::Result<usize> {
        let h = table.handle()?;
        let (mut written, mut status) = (0usize, 0i32);
        let rc = with_ctx(|c| {
            let rc = unsafe { sys::lzf_raw_compress2(c, input.as_ptr(), input.len(), cursor, h, out.as_mut_ptr(), out.len(), &mut written, &mut status) };
            if rc != sys::LZF_SUCCESS { Err(call_error(c, rc)) } else { Ok(()) }
        })?;
        rc?;
        match status {
            sys::LZF_OK => Ok(written),
            sys::LZF_WRITER_FULL => Err(io::ErrorKind::ConnectionAborted.into()),
            _ => panic!("EncoderTable contract violated"),          // src/raw/compress/mod.rs:67,92,167
        }
    }

    /// The name BASELINE.json's north_star uses: `compress2` from a fresh `U32Table` at cursor 0 into a slice.
    pub fn compress_into(input: &[u8], out: &mut [u8]) -> io::Result<usize> {
        let (mut written, mut status) = (0usize, 0i32);
        let rc = with_ctx(|c| {
            let rc = unsafe { sys::lzf_raw_compress_into(c, input.as_ptr(), input.len(), sys::LZF_TABLE_U32 as u32, 12, out.as_mut_ptr(), out.len(), &mut written, &mut status) };
            if rc != sys::LZF_SUCCESS { Err(call_error(c, rc)) } else { Ok(()) }
        })?;
        rc?;
        match status {
            sys::LZF_OK => Ok(written),
            sys::LZF_WRITER_FULL => Err(io::ErrorKind::ConnectionAborted.into()),
            _ => panic!("EncoderTable contract violated"),
        }
    }

    /// `raw::decompress_raw(input, prefix, output, output_limit)` (src/raw/decompress.rs:58-138): appends to `output`;
    /// bytes already in it are addressable history, `prefix` lies behind them.
    pub fn decompress_raw(input: &[u8], prefix: &[u8], output: &mut Vec<u8>, output_limit: usize) -> Result<(), DecodeError> {
        // history = prefix ++ what is already in the Vec; the limit counts the whole Vec (:72)
        let joined: Vec<u8>;
        let hist: &[u8] = if output.is_empty() { prefix } else { joined = [prefix, &output[..]].concat(); &joined };
        let old = output.len();
        let limit = output_limit.saturating_sub(old);
        // literals are not limit-checked (:65-67): at most input.len() of them on top of the limit
        let cap = limit.saturating_add(input.len());
        output.resize(old + cap, 0);
        let (mut n, mut status) = (0usize, 0i32);
        let res = with_ctx(|c| unsafe {
            sys::lzf_raw_decompress(c, input.as_ptr(), input.len(), hist.as_ptr(), hist.len(), output[old..].as_mut_ptr(), cap, limit, &mut n, &mut status)
        });
        let rc = res.expect("liblzfear_b200: no CUDA device (there is no CPU fallback)");
        assert_eq!(rc, sys::LZF_SUCCESS, "liblzfear_b200 call failed");
        output.truncate(old + n.min(cap));
        match status {
            sys::LZF_OK => Ok(()),
            sys::LZF_UNEXPECTED_END => Err(DecodeError::UnexpectedEnd),
            sys::LZF_MEMORY_LIMIT_EXCEEDED => Err(DecodeError::MemoryLimitExceeded),
            sys::LZF_ZERO_DEDUP_OFFSET => Err(DecodeError::ZeroDeduplicationOffset),
            _ => Err(DecodeError::InvalidDeduplicationOffset),
        }
    }
}

pub mod framed {
    use super::*;
    use std::io::{Read, Write};

    /// `CompressionError` (src/framed/compress.rs:15-23)
    #[derive(Debug, thiserror::Error)]
    pub enum CompressionError {
        #[error("error reading from the input you gave me")]
        ReadError(io::Error),
        #[error("error writing to the output you gave me")]
        WriteError(#[from] io::Error),
        #[error("the block size you asked for is not supported")]
        InvalidBlockSize,
    }

    /// `DecompressionError` (src/framed/decompress.rs:16-36)
    #[derive(Debug, thiserror::Error)]
    pub enum DecompressionError {
        #[error("error reading from the input you gave me")]
        InputError(#[from] io::Error),
        #[error("the raw LZ4 decompression failed (data corruption?)")]
        CodecError(raw::DecodeError),
        #[error("invalid header")]
        HeaderParseError(i32),
        #[error("wrong magic number in file header")]
        WrongMagic(u32),
        #[error("the header checksum was invalid")]
        HeaderChecksumFail,
        #[error("a block checksum was invalid")]
        BlockChecksumFail,
        #[error("the frame checksum was invalid")]
        FrameChecksumFail,
        #[error("stream contains a compressed block with a size so large we can't even compute it")]
        BlockLengthOverflow,
        #[error("a block decompressed to more data than allowed")]
        BlockSizeOverflow,
    }

    /// Plaintext handed to the GPU per launch of the streaming `compress` (whole blocks; at least one block).
    pub const STREAM_CHUNK_BYTES: usize = 1 << 30;

    /// Read until `n` bytes are there or the stream ends — the reference's `take(block_size).read_to_end(..)` (:227).
    fn read_up_to<R: Read>(reader: &mut R, n: usize, buf: &mut Vec<u8>) -> io::Result<usize> {
        reader.by_ref().take(n as u64).read_to_end(buf)
    }

    /// `CompressionSettings` (src/framed/compress.rs:36-133): same builder, same defaults (:44-55).
    #[derive(Clone)]
    pub struct CompressionSettings<'a> {
        independent_blocks: bool,
        block_checksums: bool,
        content_checksum: bool,
        block_size: usize,
        dictionary: Option<&'a [u8]>,
        dictionary_id: Option<u32>,
    }
    impl<'a> Default for CompressionSettings<'a> {
        fn default() -> Self {
            Self { independent_blocks: true, block_checksums: false, content_checksum: true, block_size: 4 * 1024 * 1024, dictionary: None, dictionary_id: None }
        }
    }
    impl<'a> CompressionSettings<'a> {
        pub fn independent_blocks(&mut self, v: bool) -> &mut Self { self.independent_blocks = v; self }
        pub fn block_checksums(&mut self, v: bool) -> &mut Self { self.block_checksums = v; self }
        pub fn content_checksum(&mut self, v: bool) -> &mut Self { self.content_checksum = v; self }
        pub fn block_size(&mut self, v: usize) -> &mut Self { self.block_size = v; self }
        pub fn dictionary(&mut self, id: u32, dict: &'a [u8]) -> &mut Self { self.dictionary_id = Some(id); self.dictionary = Some(dict); self }
        pub fn dictionary_id_nonsense_override(&mut self, id: Option<u32>) -> &mut Self { self.dictionary_id = id; self }

        fn settings(&self, content_size: Option<u64>) -> sys::lzf_settings {
            let mut s: sys::lzf_settings = unsafe { std::mem::zeroed() };
            unsafe { sys::lzf_settings_default(&mut s) };
            s.independent_blocks = self.independent_blocks as i32;
            s.block_checksums = self.block_checksums as i32;
            s.content_checksum = self.content_checksum as i32;
            s.block_size = self.block_size as u64;
            if let Some(d) = self.dictionary { s.dictionary = d.as_ptr(); s.dictionary_len = d.len() as u64; }
            if let Some(id) = self.dictionary_id { s.has_dictionary_id = 1; s.dictionary_id = id; }
            if let Some(n) = content_size { s.has_content_size = 1; s.content_size = n; }
            s
        }

        /// `compress` (:137-140)
        pub fn compress<R: Read, W: Write>(&self, reader: R, writer: W) -> Result<(), CompressionError> {
            self.compress_stream(reader, writer, None)
        }
        /// `compress_with_size_unchecked` (:142-145)
        pub fn compress_with_size_unchecked<R: Read, W: Write>(&self, reader: R, writer: W, content_size: u64) -> Result<(), CompressionError> {
            self.compress_stream(reader, writer, Some(content_size))
        }

        /// `compress_internal` (:159-282).  Independent blocks without a dictionary are STREAMED: the reader is consumed in
        /// chunks of whole blocks (`STREAM_CHUNK_BYTES`), every chunk is one batched launch, and its block records are
        /// spliced into the frame — the records of a chunk compressed on its own are the records the whole frame would hold
        /// (independent blocks share nothing, :265-270); the content checksum runs over the plaintext as it streams by
        /// (:172,233-235).  Dependent blocks and dictionaries carry a table from block to block: one call for the whole input.
        /// (Same scheme as `CompressionSettings._compress_streaming` of the Python mirror, which the test tiers execute.)
        fn compress_stream<R: Read, W: Write>(&self, mut reader: R, mut writer: W, content_size: Option<u64>) -> Result<(), CompressionError> {
            let streamable = self.independent_blocks && self.dictionary.is_none()
                && [64usize << 10, 256 << 10, 1 << 20, 4 << 20].contains(&self.block_size);
            let chunk_bytes = if streamable { std::cmp::max(self.block_size, STREAM_CHUNK_BYTES / self.block_size * self.block_size) } else { usize::MAX };
            let mut chunk = Vec::new();
            read_up_to(&mut reader, chunk_bytes, &mut chunk).map_err(CompressionError::ReadError)?;
            if !streamable || chunk.len() < chunk_bytes {
                return self.compress_buffer(&chunk, writer, content_size);          // everything fits one call
            }
            // the header as compress_internal writes it (:163-200): taken from an empty frame with the same settings
            let mut empty = Vec::new();
            self.compress_buffer(&[], &mut empty, content_size)?;
            let trailer = 4 + if self.content_checksum { 4 } else { 0 };
            writer.write_all(&empty[..empty.len() - trailer])?;
            // a chunk travels as a frame of its own without a content checksum: 7 header bytes | block records | EndMark
            let mut chunk_settings = self.clone();
            chunk_settings.content_checksum = false;
            chunk_settings.dictionary_id = None;
            let mut hasher = if self.content_checksum { Some(hash_new()) } else { None };
            let mut framed = Vec::new();
            while !chunk.is_empty() {
                framed.clear();
                chunk_settings.compress_buffer(&chunk, &mut framed, None)?;
                if let Some(h) = hasher.as_mut() { hash_update(h, &chunk)?; }
                writer.write_all(&framed[7..framed.len() - 4])?;
                chunk.clear();
                read_up_to(&mut reader, chunk_bytes, &mut chunk).map_err(CompressionError::ReadError)?;
            }
            writer.write_all(&0u32.to_le_bytes())?;                                  // EndMark :277
            if let Some(h) = hasher {
                writer.write_all(&unsafe { sys::lzf_xxh32_finish(&h) }.to_le_bytes())?;   // :279-281
            }
            Ok(())
        }
        /// `compress_with_size` (:147-157): the length comes from seeking, as in the reference
        pub fn compress_with_size<R: Read + io::Seek, W: Write>(&self, mut reader: R, writer: W) -> Result<(), CompressionError> {
            let start = reader.seek(io::SeekFrom::Current(0)).map_err(CompressionError::ReadError)?;
            let end = reader.seek(io::SeekFrom::End(0)).map_err(CompressionError::ReadError)?;
            reader.seek(io::SeekFrom::Start(start)).map_err(CompressionError::ReadError)?;
            self.compress_with_size_unchecked(reader, writer, end - start)
        }

        fn compress_buffer<W: Write>(&self, input: &[u8], mut writer: W, content_size: Option<u64>) -> Result<(), CompressionError> {
            let s = self.settings(content_size);
            let cap = unsafe { sys::lzf_frame_bound(&s, input.len()) };
            let mut out = vec![0u8; cap];
            let (mut written, mut status) = (0usize, 0i32);
            let rc = with_ctx(|c| {
                let rc = unsafe { sys::lzf_frame_compress(c, &s, input.as_ptr(), input.len(), out.as_mut_ptr(), cap, &mut written, &mut status) };
                if rc != sys::LZF_SUCCESS { Err(call_error(c, rc)) } else { Ok(()) }
            })?;
            rc?;
            match status {
                sys::LZF_F_OK => { writer.write_all(&out[..written])?; Ok(()) }
                sys::LZF_F_INVALID_BLOCK_SIZE => Err(CompressionError::InvalidBlockSize),
                sys::LZF_F_PANIC => panic!("called `Option::unwrap()` on a `None` value"),   // src/framed/header.rs:55
                _ => Err(CompressionError::WriteError(io::ErrorKind::WriteZero.into())),
            }
        }
    }

    pub const WINDOW_SIZE: usize = 64 * 1024;
    const FLAG_INDEPENDENT_BLOCKS: u8 = 0x20;
    const FLAG_BLOCK_CHECKSUMS: u8 = 0x10;
    const FLAG_CONTENT_CHECKSUM: u8 = 0x04;

    fn hash_update(st: &mut sys::lzf_xxh32_state, data: &[u8]) -> io::Result<()> {
        let rc = with_ctx(|c| unsafe { sys::lzf_xxh32_update(c, st, data.as_ptr(), data.len()) })?;
        if rc != sys::LZF_SUCCESS { return Err(io::Error::new(io::ErrorKind::Other, "lzf_xxh32_update failed")); }
        Ok(())
    }
    fn hash_new() -> sys::lzf_xxh32_state {
        let mut st: sys::lzf_xxh32_state = unsafe { std::mem::zeroed() };
        unsafe { sys::lzf_xxh32_init(&mut st) };
        st
    }
    fn read_u32_le<R: Read>(reader: &mut R) -> io::Result<u32> {
        let mut w = [0u8; 4];
        reader.read_exact(&mut w)?;
        Ok(u32::from_le_bytes(w))
    }

    /// `LZ4FrameReader` (src/framed/decompress.rs:81-279): the header is parsed by `new`, every `decode_block` reads one
    /// block record and hands its payload to the GPU block decoder.  (The batched, read-ahead form of this loop is
    /// `lzf_frames_decompress`; this type keeps the reference's block-at-a-time contract for callers that rely on it.)
    pub struct LZ4FrameReader<R: Read> {
        reader: R,
        flags: u8,
        block_maxsize: usize,
        content_size: Option<u64>,
        dictionary_id: Option<u32>,
        content_hasher: Option<sys::lzf_xxh32_state>,
        carryover_window: Option<Vec<u8>>,
        read_buf: Vec<u8>,
        finished: bool,
    }

    impl<R: Read> LZ4FrameReader<R> {
        /// `LZ4FrameReader::new` (:101-161).  The header parser is fed one more byte at a time until it has the whole
        /// descriptor, so nothing behind the header is taken from `reader`.
        pub fn new(mut reader: R) -> Result<Self, DecompressionError> {
            let mut hdr = vec![0u8; 4];
            reader.read_exact(&mut hdr)?;
            let info = loop {
                let mut info: sys::lzf_frame_info = unsafe { std::mem::zeroed() };
                let mut detail = 0i32;
                let status = unsafe { sys::lzf_frame_parse_header(hdr.as_ptr(), hdr.len(), &mut info, &mut detail) };
                match status {
                    sys::LZF_F_OK => break info,
                    sys::LZF_F_INPUT_ERROR => {
                        let mut one = [0u8; 1];
                        reader.read_exact(&mut one)?;
                        hdr.push(one[0]);
                    }
                    sys::LZF_F_WRONG_MAGIC => return Err(DecompressionError::WrongMagic(u32::from_le_bytes([hdr[0], hdr[1], hdr[2], hdr[3]]))),
                    sys::LZF_F_HEADER_CHECKSUM_FAIL => return Err(DecompressionError::HeaderChecksumFail),
                    _ => return Err(DecompressionError::HeaderParseError(detail)),
                }
            };
            let flags = info.flags;
            Ok(LZ4FrameReader {
                reader,
                flags,
                block_maxsize: info.block_maxsize as usize,
                content_size: if info.has_content_size != 0 { Some(info.content_size) } else { None },
                dictionary_id: if info.has_dictionary_id != 0 { Some(info.dictionary_id) } else { None },
                content_hasher: if flags & FLAG_CONTENT_CHECKSUM != 0 { Some(hash_new()) } else { None },
                carryover_window: if flags & FLAG_INDEPENDENT_BLOCKS != 0 { None } else { Some(Vec::with_capacity(WINDOW_SIZE)) },
                read_buf: Vec::new(),
                finished: false,
            })
        }

        pub fn block_size(&self) -> usize { self.block_maxsize }
        pub fn frame_size(&self) -> Option<u64> { self.content_size }
        pub fn dictionary_id(&self) -> Option<u32> { self.dictionary_id }

        pub fn into_read(self) -> LZ4FrameIoReader<'static, R> { self.into_read_with_dictionary(&[]) }
        pub fn into_read_with_dictionary<'a>(self, dictionary: &'a [u8]) -> LZ4FrameIoReader<'a, R> {
            LZ4FrameIoReader { frame_reader: self, bytes_taken: 0, buffer: Vec::new(), dictionary }
        }

        /// `decode_block` (:197-279): `output` must be empty; it stays empty once the EndMark has been read.
        pub fn decode_block(&mut self, output: &mut Vec<u8>, dictionary: &[u8]) -> Result<(), DecompressionError> {
            assert!(output.is_empty(), "You must pass an empty buffer to this interface.");
            if self.finished { return Ok(()); }
            let word = read_u32_le(&mut self.reader)?;
            if word == 0 {
                // EndMark, then the content checksum if the header promised one (:206-215)
                if let Some(hasher) = self.content_hasher.take() {
                    let expected = read_u32_le(&mut self.reader)?;
                    if unsafe { sys::lzf_xxh32_finish(&hasher) } != expected { return Err(DecompressionError::FrameChecksumFail); }
                }
                self.finished = true;
                return Ok(());
            }
            let is_compressed = word & sys::LZF_INCOMPRESSIBLE == 0;
            let block_length = (word & !sys::LZF_INCOMPRESSIBLE) as usize;
            if block_length > self.block_maxsize { return Err(DecompressionError::BlockSizeOverflow); }      // :220-222
            let mut buf = std::mem::take(&mut self.read_buf);
            buf.resize(block_length, 0);
            self.reader.read_exact(&mut buf[..])?;
            if self.flags & FLAG_BLOCK_CHECKSUMS != 0 {                                                      // :229-234
                let expected = read_u32_le(&mut self.reader)?;
                let mut st = hash_new();
                hash_update(&mut st, &buf)?;
                if unsafe { sys::lzf_xxh32_finish(&st) } != expected { return Err(DecompressionError::BlockChecksumFail); }
            }
            {
                // dependent blocks: the last 64 KiB of plaintext (or, in front of the first block, the dictionary) are the
                // decoder's prefix (:238-246)
                let dec_prefix: &[u8] = match self.carryover_window.as_mut() {
                    Some(window) => {
                        if window.is_empty() { window.extend_from_slice(dictionary); }
                        &window[..]
                    }
                    None => dictionary,
                };
                if is_compressed {
                    raw::decompress_raw(&buf, dec_prefix, output, self.block_maxsize).map_err(DecompressionError::CodecError)?;   // :248
                } else {
                    output.extend_from_slice(&buf);
                }
            }
            if let Some(window) = self.carryover_window.as_mut() {                                           // :253-269
                let outlen = output.len();
                if outlen < WINDOW_SIZE {
                    let total = window.len() + outlen;
                    if total > WINDOW_SIZE { window.drain(..total - WINDOW_SIZE); }
                    window.extend_from_slice(&output[..]);
                } else {
                    window.clear();
                    window.extend_from_slice(&output[outlen - WINDOW_SIZE..]);
                }
            }
            self.read_buf = buf;
            if output.len() > self.block_maxsize { return Err(DecompressionError::BlockSizeOverflow); }      // :272-274
            if let Some(hasher) = self.content_hasher.as_mut() { hash_update(hasher, &output[..])?; }              // :276-278
            Ok(())
        }
    }

    /// `LZ4FrameIoReader` (src/framed/decompress.rs:46-77): `Read` + `BufRead` over a frame, one block per refill —
    /// `LZ4FrameReader::new(file)?.into_read()` of examples/delz4.rs:13.
    pub struct LZ4FrameIoReader<'a, R: Read> {
        frame_reader: LZ4FrameReader<R>,
        bytes_taken: usize,
        buffer: Vec<u8>,
        dictionary: &'a [u8],
    }
    impl<R: Read> io::BufRead for LZ4FrameIoReader<'_, R> {
        fn fill_buf(&mut self) -> io::Result<&[u8]> {
            if self.bytes_taken == self.buffer.len() {
                self.buffer.clear();
                self.bytes_taken = 0;
                self.frame_reader.decode_block(&mut self.buffer, self.dictionary)
                    .map_err(|e| match e { DecompressionError::InputError(err) => err, other => io::Error::new(io::ErrorKind::Other, other) })?;
            }
            Ok(&self.buffer[self.bytes_taken..])
        }
        fn consume(&mut self, amt: usize) {
            self.bytes_taken += amt;
            assert!(self.bytes_taken <= self.buffer.len(), "You consumed more bytes than I even gave you!");
        }
    }
    impl<R: Read> Read for LZ4FrameIoReader<'_, R> {
        fn read(&mut self, buf: &mut [u8]) -> io::Result<usize> {
            use io::BufRead;
            let mybuf = self.fill_buf()?;
            let n = std::cmp::min(mybuf.len(), buf.len());
            buf[..n].copy_from_slice(&mybuf[..n]);
            self.consume(n);
            Ok(n)
        }
    }

    /// `decompress_frame` (src/framed/decompress.rs:283-288): reads one frame from `reader`, returns its plaintext.
    pub fn decompress_frame<R: Read>(mut reader: R) -> Result<Vec<u8>, DecompressionError> {
        let mut input = Vec::new();
        reader.read_to_end(&mut input)?;
        decompress_frame_with_dictionary(&input, &[]).map(|(plain, _consumed)| plain)
    }

    /// One frame from a buffer -> (plaintext, bytes of `input` the frame occupied).  The capacity grows until the frame
    /// fits (the frame header's content size is a hint the reference never verifies, decompress.rs:165-166).
    pub fn decompress_frame_with_dictionary(input: &[u8], dictionary: &[u8]) -> Result<(Vec<u8>, usize), DecompressionError> {
        let mut cap = input.len().saturating_mul(4).max(1 << 16);
        loop {
            let mut out = vec![0u8; cap];
            let (mut written, mut consumed, mut status, mut detail) = (0usize, 0usize, 0i32, 0i32);
            let rc = with_ctx(|c| {
                let rc = unsafe { sys::lzf_frame_decompress(c, input.as_ptr(), input.len(), dictionary.as_ptr(), dictionary.len(),
                                                            out.as_mut_ptr(), cap, &mut written, &mut consumed, &mut status, &mut detail) };
                if rc != sys::LZF_SUCCESS { Err(call_error(c, rc)) } else { Ok(()) }
            })?;
            rc?;
            return match status {
                sys::LZF_F_OK => { out.truncate(written); Ok((out, consumed)) }
                sys::LZF_F_WRITE_ERROR => { cap = cap.saturating_mul(4); continue; }
                sys::LZF_F_INPUT_ERROR => Err(DecompressionError::InputError(io::ErrorKind::UnexpectedEof.into())),
                sys::LZF_F_CODEC_ERROR => Err(DecompressionError::CodecError(match detail {
                    1 => raw::DecodeError::UnexpectedEnd, 2 => raw::DecodeError::MemoryLimitExceeded,
                    3 => raw::DecodeError::ZeroDeduplicationOffset, _ => raw::DecodeError::InvalidDeduplicationOffset })),
                sys::LZF_F_HEADER_PARSE_ERROR => Err(DecompressionError::HeaderParseError(detail)),
                sys::LZF_F_WRONG_MAGIC => Err(DecompressionError::WrongMagic(u32::from_le_bytes([input[0], input[1], input[2], input[3]]))),
                sys::LZF_F_HEADER_CHECKSUM_FAIL => Err(DecompressionError::HeaderChecksumFail),
                sys::LZF_F_BLOCK_CHECKSUM_FAIL => Err(DecompressionError::BlockChecksumFail),
                sys::LZF_F_FRAME_CHECKSUM_FAIL => Err(DecompressionError::FrameChecksumFail),
                sys::LZF_F_BLOCK_LENGTH_OVERFLOW => Err(DecompressionError::BlockLengthOverflow),
                _ => Err(DecompressionError::BlockSizeOverflow),
            };
        }
    }
}
