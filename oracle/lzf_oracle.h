/*
 * lzf_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the lz-fear hot path: the raw LZ4 block codec,
 * XXH32, and the frame glue around them.  It exists to CHECK the CUDA product
 * path; nothing under rust-lz-fear_b200/ links, loads or calls it.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may use it.
 *
 * Parity pinning: the reference is pure Rust and there is no Rust toolchain in
 * this image, so oracle/_ref cannot be built ("unbuildable here").  The oracle
 * is pinned instead against (tests/test_oracle.py):
 *   - the reference's own decode KATs      src/raw/decompress.rs:153-175
 *   - the reference's roundtrip tests      src/lib.rs:43-106
 *   - tests/issue-15.rs input (dependent 64 KiB blocks roundtrip)
 *   - the three fuzz corpora under fuzz/corpus (packed in tests/golden/)
 *   - liblz4 1.9.4 (LZ4_compress_default / LZ4_decompress_safe / LZ4F_*),
 *     which tests/output_equivalence.rs names as the expected output, on the
 *     input classes where the two are known to coincide
 *   - python xxhash 3.7.0 for XXH32 (twox-hash is not vendored in the reference)
 *
 * Every function cites the reference file:line it follows
 * (paths relative to /root/reference).
 */
#ifndef LZF_ORACLE_H
#define LZF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- raw codec status codes (src/raw/decompress.rs:7-17 + writer-full) ---- */
enum {
    LZFO_OK = 0,
    LZFO_UNEXPECTED_END = 1,          /* DecodeError::UnexpectedEnd */
    LZFO_MEMORY_LIMIT_EXCEEDED = 2,   /* DecodeError::MemoryLimitExceeded */
    LZFO_ZERO_DEDUP_OFFSET = 3,       /* DecodeError::ZeroDeduplicationOffset */
    LZFO_INVALID_DEDUP_OFFSET = 4,    /* DecodeError::InvalidDeduplicationOffset */
    LZFO_WRITER_FULL = 5,             /* io::ErrorKind::ConnectionAborted from NoPartialWrites */
    LZFO_OUTPUT_CAP = 6,              /* physical output buffer too small (no reference analogue) */
    LZFO_PANIC = 7                    /* a reference assert!/expect()/unwrap() would fire */
};

/* ---- frame-level codes (src/framed/decompress.rs:16-36, compress.rs:15-23) ---- */
enum {
    LZFO_F_OK = 0,
    LZFO_F_INPUT_ERROR = 10,          /* DecompressionError::InputError (short read) */
    LZFO_F_CODEC_ERROR = 11,          /* CodecError(raw::DecodeError) — detail in *codec */
    LZFO_F_HEADER_PARSE_ERROR = 12,   /* HeaderParseError(ParseError)  — detail in *codec */
    LZFO_F_WRONG_MAGIC = 13,
    LZFO_F_HEADER_CHECKSUM_FAIL = 14,
    LZFO_F_BLOCK_CHECKSUM_FAIL = 15,
    LZFO_F_FRAME_CHECKSUM_FAIL = 16,
    LZFO_F_BLOCK_LENGTH_OVERFLOW = 17,
    LZFO_F_BLOCK_SIZE_OVERFLOW = 18,
    LZFO_F_INVALID_BLOCK_SIZE = 20,   /* CompressionError::InvalidBlockSize */
    LZFO_F_WRITE_ERROR = 21,          /* CompressionError::WriteError (output cap) */
    LZFO_F_PANIC = 22
};
/* header::ParseError detail (src/framed/header.rs:18-28) */
enum {
    LZFO_P_UNIMPLEMENTED_BLOCKSIZE = 1,
    LZFO_P_UNSUPPORTED_VERSION = 2,
    LZFO_P_RESERVED_FLAG_BITS = 3,
    LZFO_P_RESERVED_BD_BITS = 4
};

/* ---- XXH32 (twox-hash XxHash32::with_seed; public algorithm) ---- */
typedef struct {
    uint32_t acc[4];
    uint8_t buf[16];
    uint32_t buflen;
    uint64_t total;
    uint32_t seed;
} lzfo_xxh32_state;
void lzfo_xxh32_init(lzfo_xxh32_state* s, uint32_t seed);
void lzfo_xxh32_update(lzfo_xxh32_state* s, const void* data, size_t n);
uint32_t lzfo_xxh32_finish(const lzfo_xxh32_state* s);
uint32_t lzfo_xxh32(const void* data, size_t n, uint32_t seed);

/* ---- EncoderTable (src/raw/compress/mod.rs:19-101) ---- */
enum { LZFO_TABLE_U32 = 0, LZFO_TABLE_U16 = 1 };
typedef struct lzfo_table lzfo_table;
/* hashlog: 12 is the reference's const HASHLOG; other values are the config-5 extension. */
lzfo_table* lzfo_table_new(int kind, unsigned hashlog);
lzfo_table* lzfo_table_clone(const lzfo_table* t);
void lzfo_table_free(lzfo_table* t);
size_t lzfo_table_payload_size_limit(const lzfo_table* t);
/* returns LZFO_OK or LZFO_PANIC ("EncoderTable contract violated"); *old gets the swapped-out position */
int lzfo_table_replace(lzfo_table* t, const uint8_t* input, size_t len, size_t pos, size_t* old);
void lzfo_table_offset(lzfo_table* t, size_t by);

/* worst-case size of compress2 output into an unbounded writer */
size_t lzfo_compress_bound(size_t n);

/* raw::compress2 (src/raw/compress/mod.rs:165-238) writing through NoPartialWrites(out[..cap])
 * (src/framed/compress.rs:294-308).  Returns LZFO_OK, LZFO_WRITER_FULL or LZFO_PANIC. */
int lzfo_compress2(const uint8_t* input, size_t len, size_t cursor, lzfo_table* table,
                   uint8_t* out, size_t cap, size_t* written);

/* convenience: fresh table, cursor 0 */
int lzfo_compress_block(const uint8_t* input, size_t len, int table_kind, unsigned hashlog,
                        uint8_t* out, size_t cap, size_t* written);

/* raw::decompress_raw (src/raw/decompress.rs:58-138).  `out` holds *out_len bytes of
 * pre-existing history on entry (Vec contents) and is appended to; `out_cap` is the physical
 * size of `out`.  Returns a raw status code. */
int lzfo_decompress_raw(const uint8_t* in, size_t n, const uint8_t* prefix, size_t plen,
                        uint8_t* out, size_t out_cap, size_t out_limit, size_t* out_len);

/* ---- frame glue ---- */
typedef struct {
    int independent_blocks;   /* default 1  (src/framed/compress.rs:47) */
    int block_checksums;      /* default 0 */
    int content_checksum;     /* default 1 */
    uint64_t block_size;      /* default 4 MiB */
    const uint8_t* dictionary; /* nullable */
    uint64_t dictionary_len;
    int has_dictionary_id;
    uint32_t dictionary_id;
    int has_content_size;     /* compress_with_size* */
    uint64_t content_size;
    uint32_t hashlog;         /* 0 => 12 (reference) */
} lzfo_settings;
void lzfo_settings_default(lzfo_settings* s);

size_t lzfo_frame_bound(const lzfo_settings* s, size_t n);
/* CompressionSettings::compress_internal (src/framed/compress.rs:159-282) */
int lzfo_frame_compress(const lzfo_settings* s, const uint8_t* in, size_t n,
                        uint8_t* out, size_t cap, size_t* written);

typedef struct {
    uint8_t flags;
    uint64_t block_maxsize;
    int has_content_size;
    uint64_t content_size;
    int has_dictionary_id;
    uint32_t dictionary_id;
    size_t header_len;
} lzfo_frame_info;
/* LZ4FrameReader::new (src/framed/decompress.rs:101-161); *detail gets the ParseError kind */
int lzfo_frame_parse_header(const uint8_t* in, size_t n, lzfo_frame_info* info, int* detail);
/* decompress_frame / into_read_with_dictionary + read_to_end (src/framed/decompress.rs:197-288).
 * On error, *written is the number of bytes of fully decoded blocks before the failure and
 * *detail the raw status / ParseError kind.  *consumed = bytes of `in` consumed on success. */
int lzfo_frame_decompress(const uint8_t* in, size_t n, const uint8_t* dict, size_t dlen,
                          uint8_t* out, size_t cap, size_t* written, size_t* consumed, int* detail);

/* ---- multi-threaded batch drivers (CPU baseline; one block per task) ---- */
int lzfo_compress_blocks_mt(const uint8_t* in, const uint64_t* in_off, const uint32_t* in_len,
                            uint32_t nblocks, unsigned hashlog, uint8_t* out, const uint64_t* out_off,
                            uint32_t* out_len, int32_t* status, int nthreads);
int lzfo_decompress_blocks_mt(const uint8_t* in, const uint64_t* in_off, const uint32_t* in_len,
                              uint32_t nblocks, uint8_t* out, const uint64_t* out_off,
                              const uint32_t* out_cap, const uint32_t* out_limit, uint32_t* out_len,
                              int32_t* status, int nthreads);

/* second CPU bar: C lz4 1.9.x (liblz4.so.1 via dlopen) on the same thread pool.  compress != 0:
 * LZ4_compress_default into out_cap[b] bytes (status 1 = does not fit); else LZ4_decompress_safe.
 * Returns -1 when the library is not installed. */
int lzfo_liblz4_blocks_mt(int compress, const uint8_t* in, const uint64_t* in_off, const uint32_t* in_len, uint32_t nblocks,
                          uint8_t* out, const uint64_t* out_off, const uint32_t* out_cap, uint32_t* out_len,
                          int32_t* status, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
