// lz_fear.hpp — header-only C++ host mirror of the `lz-fear` crate's public surface over the C ABI of
// include/lzfear_b200.h.  Same names, argument meaning and error behaviour as the Rust originals
// (file:line of each original in the comments); Rust `Read`/`Write` become std::istream/std::ostream,
// `Result<_, E>` becomes an exception carrying the same variant.  There is no CPU implementation
// behind it: every call goes to the sm_100a kernels through liblzfear_b200.so.
#pragma once

#include <cstdint>
#include <cstring>
#include <istream>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lzfear_b200.h"

namespace lz_fear {

struct CallError : std::runtime_error {          // CUDA / argument failure below the codec (no Rust analogue)
    int code;
    CallError(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// One lzf_ctx (CUDA device + streams + scratch).  Not thread-safe; one per thread.
class Context {
  public:
    explicit Context(int device = 0) {
        const int rc = lzf_create(device, &c_);
        if (rc != LZF_SUCCESS) throw CallError(rc, rc == LZF_ERR_NO_DEVICE ? "no CUDA device (there is no CPU fallback)" : "lzf_create failed");
    }
    ~Context() { lzf_destroy(c_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    lzf_ctx* get() const { return c_; }
    void check(int rc) const { if (rc != LZF_SUCCESS) throw CallError(rc, lzf_last_error(c_)); }
    static Context& default_context() { static Context ctx(0); return ctx; }

  private:
    lzf_ctx* c_ = nullptr;
};

namespace raw {

// raw::DecodeError — src/raw/decompress.rs:7-17
struct DecodeError : std::runtime_error {
    enum Kind { UnexpectedEnd = 1, MemoryLimitExceeded = 2, ZeroDeduplicationOffset = 3, InvalidDeduplicationOffset = 4 } kind;
    explicit DecodeError(int k) : std::runtime_error(name(k)), kind((Kind)k) {}
    static const char* name(int k) {
        switch (k) { case 1: return "UnexpectedEnd"; case 2: return "MemoryLimitExceeded";
                     case 3: return "ZeroDeduplicationOffset"; default: return "InvalidDeduplicationOffset"; }
    }
};
// io::ErrorKind::ConnectionAborted out of NoPartialWrites — src/framed/compress.rs:298-301
struct WriterFull : std::runtime_error { WriterFull() : std::runtime_error("ConnectionAborted") {} };

// EncoderTable flavours — src/raw/compress/mod.rs:19-101.  A fresh table only selects the flavour (the kernel zeroes
// its own copy); the first call that carries state — cursor > 0, offset(), a second compress2 — makes it a
// device-resident object (lzf_table_*) with the reference's replace / offset semantics.
template <uint32_t kKind, uint64_t kLimit>
struct DeviceTable {
    static constexpr uint32_t kind = kKind;
    static size_t payload_size_limit() { return (size_t)kLimit; }
    bool fresh = true;
    DeviceTable() = default;
    DeviceTable(const DeviceTable&) = delete;
    DeviceTable& operator=(const DeviceTable&) = delete;
    ~DeviceTable() { if (h_) lzf_table_destroy(ctx_ ? ctx_->get() : nullptr, h_); }
    void offset(size_t by) { pending_ += by; fresh = false; }                  // EncoderTable::offset :72-74,97-99
    lzf_table* device(Context& ctx) {
        if (!h_) { ctx_ = &ctx; ctx.check(lzf_table_create(ctx.get(), kKind, 12, &h_)); }
        if (pending_) { ctx.check(lzf_table_offset(ctx.get(), h_, pending_)); pending_ = 0; }
        return h_;
    }
  private:
    lzf_table* h_ = nullptr;
    Context* ctx_ = nullptr;
    uint64_t pending_ = 0;
};
using U32Table = DeviceTable<LZF_TABLE_U32, 0xffffffffull>;
using U16Table = DeviceTable<LZF_TABLE_U16, 0xffffull>;

// compress2 through NoPartialWrites(out[..cap]) from a fresh table at cursor 0 — src/raw/compress/mod.rs:165-238,
// src/framed/compress.rs:242
template <class Table = U32Table>
inline size_t compress_into(const uint8_t* input, size_t n, uint8_t* out, size_t cap, Context& ctx = Context::default_context()) {
    if (n > Table::payload_size_limit()) throw std::logic_error("assertion failed: input.len() <= T::payload_size_limit()");
    size_t written = 0;
    int32_t status = 0;
    ctx.check(lzf_raw_compress_into(ctx.get(), input, n, Table::kind, 12, out, cap, &written, &status));
    if (status == LZF_WRITER_FULL) throw WriterFull();
    if (status != LZF_OK) throw std::logic_error("EncoderTable contract violated");
    return written;
}

// raw::compress2(input, cursor, &mut table, writer) — src/raw/compress/mod.rs:165-170: input[..cursor] is match-only
// history, the table keeps its entries from call to call (src/framed/compress.rs:220,270-275).  Any writer with
// write(const char*, std::streamsize); an unbounded writer never refuses.
template <class Writer, class Table>
inline void compress2(const uint8_t* input, size_t n, size_t cursor, Table& table, Writer& writer, Context& ctx = Context::default_context()) {
    if (n > Table::payload_size_limit()) throw std::logic_error("assertion failed: input.len() <= T::payload_size_limit()");
    std::vector<uint8_t> buf(lzf_compress_bound(n));
    size_t written = 0;
    int32_t status = 0;
    if (n) table.fresh = false;
    ctx.check(lzf_raw_compress2(ctx.get(), input, n, cursor, table.device(ctx), buf.data(), buf.size(), &written, &status));
    if (status == LZF_WRITER_FULL) throw WriterFull();
    if (status != LZF_OK) throw std::logic_error("EncoderTable contract violated");
    writer.write(reinterpret_cast<const char*>(buf.data()), (std::streamsize)written);
}

// raw::decompress_raw — src/raw/decompress.rs:58-78: appends to `output`; bytes already in it are history
inline void decompress_raw(const uint8_t* input, size_t n, const uint8_t* prefix, size_t plen, std::vector<uint8_t>& output,
                           size_t output_limit, Context& ctx = Context::default_context()) {
    std::vector<uint8_t> hist;
    const size_t old = output.size();
    if (old) {
        hist.assign(prefix, prefix + plen);
        hist.insert(hist.end(), output.begin(), output.end());
        prefix = hist.data();
        plen = hist.size();
    }
    size_t limit = output_limit > old ? output_limit - old : 0;
    if (limit > 0xffffffffull) limit = 0xffffffffull;
    size_t cap = limit + n + 16;
    if (cap > (size_t(1) << 32)) cap = size_t(1) << 32;
    output.resize(old + cap);
    size_t out_len = 0;
    int32_t status = 0;
    const int rc = lzf_raw_decompress(ctx.get(), input, n, prefix, plen, output.data() + old, cap, limit, &out_len, &status);
    output.resize(old + (out_len < cap ? out_len : cap));
    ctx.check(rc);
    if (status >= 1 && status <= 4) throw DecodeError(status);
}

}  // namespace raw

namespace framed {

constexpr uint32_t MAGIC = LZF_MAGIC;                 // src/framed/mod.rs:16
constexpr size_t WINDOW_SIZE = LZF_WINDOW_SIZE;       // src/framed/mod.rs:20

// CompressionError — src/framed/compress.rs:15-23 ; DecompressionError — src/framed/decompress.rs:16-36
struct FrameError : std::runtime_error {
    int status, detail;
    FrameError(int s, int d) : std::runtime_error(name(s)), status(s), detail(d) {}
    static const char* name(int s) {
        switch (s) {
            case LZF_F_INPUT_ERROR: return "InputError"; case LZF_F_CODEC_ERROR: return "CodecError";
            case LZF_F_HEADER_PARSE_ERROR: return "HeaderParseError"; case LZF_F_WRONG_MAGIC: return "WrongMagic";
            case LZF_F_HEADER_CHECKSUM_FAIL: return "HeaderChecksumFail"; case LZF_F_BLOCK_CHECKSUM_FAIL: return "BlockChecksumFail";
            case LZF_F_FRAME_CHECKSUM_FAIL: return "FrameChecksumFail"; case LZF_F_BLOCK_LENGTH_OVERFLOW: return "BlockLengthOverflow";
            case LZF_F_BLOCK_SIZE_OVERFLOW: return "BlockSizeOverflow"; case LZF_F_INVALID_BLOCK_SIZE: return "InvalidBlockSize";
            case LZF_F_WRITE_ERROR: return "WriteError"; default: return "panic";
        }
    }
};

inline std::vector<uint8_t> slurp(std::istream& r) {
    std::vector<uint8_t> v;
    char buf[1 << 16];
    while (r.read(buf, sizeof(buf)) || r.gcount()) v.insert(v.end(), buf, buf + r.gcount());
    return v;
}

// CompressionSettings — src/framed/compress.rs:36-157 (same setters, same defaults)
class CompressionSettings {
  public:
    CompressionSettings() { lzf_settings_default(&s_); }
    static CompressionSettings default_() { return CompressionSettings(); }
    CompressionSettings& independent_blocks(bool v) { s_.independent_blocks = v; return *this; }
    CompressionSettings& block_checksums(bool v) { s_.block_checksums = v; return *this; }
    CompressionSettings& content_checksum(bool v) { s_.content_checksum = v; return *this; }
    CompressionSettings& block_size(size_t v) { s_.block_size = v; return *this; }
    CompressionSettings& dictionary(uint32_t id, const std::vector<uint8_t>& dict) {
        dict_ = dict; s_.dictionary = dict_.data(); s_.dictionary_len = dict_.size(); s_.has_dictionary_id = 1; s_.dictionary_id = id; return *this;
    }
    CompressionSettings& dictionary_id_nonsense_override(bool has, uint32_t id = 0) { s_.has_dictionary_id = has; s_.dictionary_id = id; return *this; }

    void compress(std::istream& reader, std::ostream& writer, Context& ctx = Context::default_context()) const { run(reader, writer, ctx, 0, 0); }
    void compress_with_size_unchecked(std::istream& reader, std::ostream& writer, uint64_t content_size, Context& ctx = Context::default_context()) const {
        run(reader, writer, ctx, 1, content_size);
    }
    void compress_with_size(std::istream& reader, std::ostream& writer, Context& ctx = Context::default_context()) const { run(reader, writer, ctx, 2, 0); }

  private:
    void run(std::istream& reader, std::ostream& writer, Context& ctx, int size_mode, uint64_t content_size) const {
        const std::vector<uint8_t> in = slurp(reader);
        lzf_settings s = s_;
        s.has_content_size = size_mode;
        s.content_size = content_size;
        std::vector<uint8_t> out(lzf_frame_bound(&s, in.size()));
        size_t written = 0;
        int32_t status = 0;
        ctx.check(lzf_frame_compress(ctx.get(), &s, in.data(), in.size(), out.data(), out.size(), &written, &status));
        if (status != LZF_F_OK) throw FrameError(status, 0);
        writer.write(reinterpret_cast<const char*>(out.data()), (std::streamsize)written);
    }
    lzf_settings s_;
    std::vector<uint8_t> dict_;
};

class LZ4FrameIoReader;

// LZ4FrameReader — src/framed/decompress.rs:81-279
class LZ4FrameReader {
  public:
    explicit LZ4FrameReader(std::istream& reader, Context& ctx = Context::default_context()) : r_(reader), ctx_(ctx) {
        uint8_t hdr[19];
        size_t n = 0;
        int32_t detail = 0;
        int st = LZF_F_INPUT_ERROR;
        while (st == LZF_F_INPUT_ERROR) {                       // feed the parser field by field, like the reference reads
            if (n == sizeof(hdr) || !r_.read(reinterpret_cast<char*>(hdr + n), 1)) throw FrameError(LZF_F_INPUT_ERROR, 0);
            n++;
            st = n < 4 ? LZF_F_INPUT_ERROR : lzf_frame_parse_header(hdr, n, &info_, &detail);
        }
        if (st != LZF_F_OK) throw FrameError(st, detail);
        if (info_.flags & 0x04) { lzf_xxh32_init(&content_hasher_); hashing_ = true; }
        dependent_ = !(info_.flags & 0x20);
    }
    size_t block_size() const { return (size_t)info_.block_maxsize; }
    bool frame_size(uint64_t* v) const { if (info_.has_content_size) *v = info_.content_size; return info_.has_content_size; }
    bool dictionary_id(uint32_t* v) const { if (info_.has_dictionary_id) *v = info_.dictionary_id; return info_.has_dictionary_id; }

    // decode_block — src/framed/decompress.rs:197-279.  `output` must be empty.
    void decode_block(std::vector<uint8_t>& output, const std::vector<uint8_t>& dictionary = {}) {
        if (!output.empty()) throw std::logic_error("You must pass an empty buffer to this interface.");
        if (finished_) return;
        uint32_t block_length = read_u32();
        if (block_length == 0) {
            if (hashing_) {
                hashing_ = false;
                if (lzf_xxh32_finish(&content_hasher_) != read_u32()) throw FrameError(LZF_F_FRAME_CHECKSUM_FAIL, 0);
            }
            finished_ = true;
            return;
        }
        const bool is_compressed = !(block_length & LZF_INCOMPRESSIBLE);
        block_length &= ~LZF_INCOMPRESSIBLE;
        if (block_length > block_size()) throw FrameError(LZF_F_BLOCK_SIZE_OVERFLOW, 0);
        read_buf_.resize(block_length);
        if (block_length && !r_.read(reinterpret_cast<char*>(read_buf_.data()), block_length)) throw FrameError(LZF_F_INPUT_ERROR, 0);
        if (info_.flags & 0x10) {
            const uint32_t checksum = read_u32();
            lzf_xxh32_state h;
            lzf_xxh32_init(&h);
            ctx_.check(lzf_xxh32_update(ctx_.get(), &h, read_buf_.data(), read_buf_.size()));
            if (lzf_xxh32_finish(&h) != checksum) throw FrameError(LZF_F_BLOCK_CHECKSUM_FAIL, 0);
        }
        const std::vector<uint8_t>* prefix = &dictionary;
        if (dependent_) {
            if (window_.empty()) window_ = dictionary;
            prefix = &window_;
        }
        if (is_compressed) {
            try { raw::decompress_raw(read_buf_.data(), read_buf_.size(), prefix->data(), prefix->size(), output, block_size(), ctx_); }
            catch (const raw::DecodeError& e) { throw FrameError(LZF_F_CODEC_ERROR, (int)e.kind); }
        } else {
            output = read_buf_;
        }
        if (dependent_) {
            const size_t outlen = output.size();
            if (outlen < WINDOW_SIZE) {
                const size_t avail = window_.size() + outlen;
                if (avail >= WINDOW_SIZE) window_.erase(window_.begin(), window_.begin() + (avail - WINDOW_SIZE));
                window_.insert(window_.end(), output.begin(), output.end());
            } else {
                window_.assign(output.end() - WINDOW_SIZE, output.end());
            }
        }
        if (output.size() > block_size()) throw FrameError(LZF_F_BLOCK_SIZE_OVERFLOW, 0);
        if (hashing_) ctx_.check(lzf_xxh32_update(ctx_.get(), &content_hasher_, output.data(), output.size()));
    }

    LZ4FrameIoReader into_read();
    LZ4FrameIoReader into_read_with_dictionary(const std::vector<uint8_t>& dictionary);

  private:
    uint32_t read_u32() {
        uint8_t b[4];
        if (!r_.read(reinterpret_cast<char*>(b), 4)) throw FrameError(LZF_F_INPUT_ERROR, 0);
        return uint32_t(b[0]) | (uint32_t(b[1]) << 8) | (uint32_t(b[2]) << 16) | (uint32_t(b[3]) << 24);
    }
    std::istream& r_;
    Context& ctx_;
    lzf_frame_info info_{};
    lzf_xxh32_state content_hasher_{};
    bool hashing_ = false, dependent_ = false, finished_ = false;
    std::vector<uint8_t> read_buf_, window_;
};

// LZ4FrameIoReader — src/framed/decompress.rs:46-77 (Read + BufRead)
class LZ4FrameIoReader {
  public:
    LZ4FrameIoReader(LZ4FrameReader& fr, std::vector<uint8_t> dict) : fr_(fr), dict_(std::move(dict)) {}
    const uint8_t* fill_buf(size_t* n) {
        if (taken_ == buffer_.size()) {
            buffer_.clear();
            fr_.decode_block(buffer_, dict_);
            taken_ = 0;
        }
        *n = buffer_.size() - taken_;
        return buffer_.data() + taken_;
    }
    void consume(size_t amt) {
        taken_ += amt;
        if (taken_ > buffer_.size()) throw std::logic_error("You consumed more bytes than I even gave you!");
    }
    size_t read(uint8_t* buf, size_t len) {
        size_t n;
        const uint8_t* p = fill_buf(&n);
        const size_t take = n < len ? n : len;
        std::memcpy(buf, p, take);
        consume(take);
        return take;
    }
    void read_to_end(std::vector<uint8_t>& out) {            // stops at the first read that yields 0 bytes
        for (;;) {
            size_t n;
            const uint8_t* p = fill_buf(&n);
            if (n == 0) return;
            out.insert(out.end(), p, p + n);
            consume(n);
        }
    }

  private:
    LZ4FrameReader& fr_;
    std::vector<uint8_t> dict_, buffer_;
    size_t taken_ = 0;
};
inline LZ4FrameIoReader LZ4FrameReader::into_read() { return LZ4FrameIoReader(*this, {}); }
inline LZ4FrameIoReader LZ4FrameReader::into_read_with_dictionary(const std::vector<uint8_t>& d) { return LZ4FrameIoReader(*this, d); }

// decompress_frame — src/framed/decompress.rs:283-288: the whole frame in ONE batched GPU call
inline std::vector<uint8_t> decompress_frame(std::istream& reader, Context& ctx = Context::default_context()) {
    const std::vector<uint8_t> in = slurp(reader);
    lzf_frame_info info;
    int32_t detail = 0;
    int st = lzf_frame_parse_header(in.data(), in.size(), &info, &detail);
    if (st != LZF_F_OK) throw FrameError(st, detail);
    size_t cap = info.has_content_size && info.content_size < (uint64_t(1) << 36) ? (size_t)info.content_size + 64 : in.size() * 255 + (1 << 20);
    for (;;) {
        std::vector<uint8_t> out(cap);
        size_t written = 0, consumed = 0;
        int32_t status = 0;
        ctx.check(lzf_frame_decompress(ctx.get(), in.data(), in.size(), nullptr, 0, out.data(), out.size(), &written, &consumed, &status, &detail));
        if (status == LZF_F_WRITE_ERROR && cap < in.size() * 255 + (1 << 20)) { cap = in.size() * 255 + (1 << 20); continue; }
        if (status != LZF_F_OK) throw FrameError(status, detail);
        out.resize(written);
        return out;
    }
}

}  // namespace framed
}  // namespace lz_fear
