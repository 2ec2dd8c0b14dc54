/*
 * lzfear_b200.h — C ABI of the B200-native LZ4 block codec that drops in for the
 * hot path of the pure-Rust `lz-fear` crate.
 *
 * The reference has no FFI of its own (`#![forbid(unsafe_code)]`, src/lib.rs:1); the seam this
 * ABI replaces is the pair of generic Rust calls
 *     compress2(&in_buffer, window_offset, &mut table, &mut NoPartialWrites(&mut out[..n]))
 *                                                        src/framed/compress.rs:242-243
 *     raw::decompress_raw(buf, dec_prefix, output, self.block_maxsize)
 *                                                        src/framed/decompress.rs:247-248
 * batched over many independent blocks, plus the per-block loop around them
 * (src/framed/compress.rs:221-276, src/framed/decompress.rs:197-279) and the XXH32 call
 * sites (twox-hash; src/framed/compress.rs:172,197-199,233-235,260-262,279-281).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary
 *   - `d_` pointers are device memory on the ctx's GPU, everything else is host memory
 *   - functions return a call-level code (LZF_SUCCESS or < 0); codec results are reported
 *     PER BLOCK in `status[]` so one bad block never poisons a batch
 *   - there is NO CPU fallback: without a CUDA device every entry point fails with
 *     LZF_ERR_NO_DEVICE / LZF_ERR_CUDA
 *   - batched device calls are asynchronous on `stream` (a cudaStream_t passed as void*);
 *     host-buffer calls synchronise before returning
 *   - nothing here throws or aborts across the ABI
 *
 * Buffers
 *   - blocks and frames may start at ANY byte address; lengths are arbitrary
 *   - the kernels read whole aligned words / 16-byte TMA granules: of a block's input (and of the history in
 *     front of it) they may READ — never write — the other bytes of the aligned 16-byte granules that hold its
 *     first and its last byte.  Every allocator whose alignment and granularity are at least 16 bytes satisfies
 *     this (cudaMalloc: 256, pinned host pages, the usual sub-allocators); a buffer whose first or last granule is
 *     cut by an allocation boundary finer than 16 bytes must be padded by the caller
 *   - output bytes between out_len and the capacity of a block / frame are unspecified after a call
 *
 * Concurrency
 *   - a ctx may be used from several host threads; calls that share a pipeline slot serialise internally
 *   - batched device calls of one ctx run one after the other even when given different streams (they share the
 *     ctx's work queue and table scratch): a launch on another stream first waits for the previous one.  Use one
 *     ctx per stream for independent queues
 *
 * Environment (read ONCE, by lzf_create; tuning and test knobs, never needed for correct results)
 *     LZF_B200_CHUNK_BYTES, LZF_B200_FEED_SLICE, LZF_B200_FEED_MIN_BLOCKS, LZF_B200_ENC_U32,
 *     LZF_B200_ENC_SMEM_WARPS, LZF_B200_DEC_CTAS_PER_SM, LZF_B200_TRACE
 *     LZF_B200_TEST_POS_LIMIT (tests only): lowers the stream position at which a dependent-block frame answers
 *     LZF_F_PANIC ("EncoderTable contract violated") from 2^32 - 1, so that the rule can be tested with small frames
 */
#ifndef LZFEAR_B200_H
#define LZFEAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LZF_ABI_VERSION 3

/* ---- call-level return codes ---- */
enum {
    LZF_SUCCESS = 0,
    LZF_ERR_INVALID_ARG = -1,
    LZF_ERR_CUDA = -2,
    LZF_ERR_NO_DEVICE = -3,
    LZF_ERR_OOM = -4,
    LZF_ERR_UNSUPPORTED = -5
};

/* ---- per-block codec status ----
 * 1..4 = raw::DecodeError variants in declaration order (src/raw/decompress.rs:7-17);
 * 5    = the bounded writer refused a write (io::ErrorKind::ConnectionAborted from
 *        NoPartialWrites, src/framed/compress.rs:298-301) => the frame layer stores the block raw;
 * 6    = the physical output buffer was too small (reference Vec would simply have grown);
 * 7    = a reference assert!/expect() would have fired (e.g. U16Table with > 65535 bytes,
 *        src/raw/compress/mod.rs:167). */
enum {
    LZF_OK = 0,
    LZF_UNEXPECTED_END = 1,
    LZF_MEMORY_LIMIT_EXCEEDED = 2,
    LZF_ZERO_DEDUP_OFFSET = 3,
    LZF_INVALID_DEDUP_OFFSET = 4,
    LZF_WRITER_FULL = 5,
    LZF_OUTPUT_CAP = 6,
    LZF_PANIC = 7
};

/* ---- frame-level status (DecompressionError src/framed/decompress.rs:16-36,
 *      CompressionError src/framed/compress.rs:15-23) ---- */
enum {
    LZF_F_OK = 0,
    LZF_F_INPUT_ERROR = 10,
    LZF_F_CODEC_ERROR = 11,           /* detail = per-block codec status 1..4 */
    LZF_F_HEADER_PARSE_ERROR = 12,    /* detail = LZF_P_* */
    LZF_F_WRONG_MAGIC = 13,
    LZF_F_HEADER_CHECKSUM_FAIL = 14,
    LZF_F_BLOCK_CHECKSUM_FAIL = 15,
    LZF_F_FRAME_CHECKSUM_FAIL = 16,
    LZF_F_BLOCK_LENGTH_OVERFLOW = 17,
    LZF_F_BLOCK_SIZE_OVERFLOW = 18,
    LZF_F_INVALID_BLOCK_SIZE = 20,
    LZF_F_WRITE_ERROR = 21,           /* caller's output buffer too small */
    LZF_F_PANIC = 22                  /* reference would panic (e.g. header.rs:55 unwrap) */
};
enum {                                /* header::ParseError, src/framed/header.rs:18-28 */
    LZF_P_UNIMPLEMENTED_BLOCKSIZE = 1,
    LZF_P_UNSUPPORTED_VERSION = 2,
    LZF_P_RESERVED_FLAG_BITS = 3,
    LZF_P_RESERVED_BD_BITS = 4
};

/* EncoderTable implementations, src/raw/compress/mod.rs:27-36 (U32Table), :78-101 (U16Table) */
enum { LZF_TABLE_U32 = 0, LZF_TABLE_U16 = 1 };

/* high bit of a block length word: "stored, not compressed" (src/framed/mod.rs:18) */
#define LZF_INCOMPRESSIBLE 0x80000000u
#define LZF_MAGIC 0x184D2204u          /* src/framed/mod.rs:16 */
#define LZF_WINDOW_SIZE 65536u         /* src/framed/mod.rs:20 */

typedef struct lzf_ctx lzf_ctx;

int lzf_abi_version(void);
/* Creates a context bound to CUDA device `device`.  Fails (LZF_ERR_NO_DEVICE) without a GPU. */
int lzf_create(int device, lzf_ctx** ctx);
void lzf_destroy(lzf_ctx* ctx);
/* Human-readable text for the last failing call on this ctx (valid until the next call). */
const char* lzf_last_error(const lzf_ctx* ctx);
/* Options of a ctx.
 *   LZF_OPT_SEGMENT_BYTES   0 (default): every block is parsed exactly like the reference, by one warp; output bytes are
 *                           the reference's.  >= 65536: a compress call whose blocks cannot fill the GPU (fewer than half
 *                           the resident warps, e.g. one 64 MiB file = 16 blocks) cuts every block into segments of at
 *                           least this many bytes, parses them side by side — each from a table primed with the 64 KiB
 *                           in front of it — and stitches the sequence streams back into ONE LZ4 block.  The result is
 *                           valid LZ4 that decodes to the input, within a fraction of a percent of the reference's size,
 *                           but NOT its bytes (BASELINE.json north_star: "compressed size within 1 %"); launches that
 *                           fill the GPU are unaffected.  Applies to independent blocks without a dictionary. */
enum { LZF_OPT_SEGMENT_BYTES = 1 };
int lzf_set_option(lzf_ctx* ctx, int option, uint64_t value);
/* Releases the ctx's grow-only scratch (device staging buffers, descriptor arenas, table / segment scratch) — e.g.
 * after one very large call.  No call may be in flight on the ctx. */
int lzf_trim(lzf_ctx* ctx);
/* Number of kernels this ctx has launched so far (bench.py's gpu_launches). */
uint64_t lzf_launch_count(const lzf_ctx* ctx);

/* ------------------------------------------------------------------------------------------
 * Batched block compress — replaces the per-block `compress2` call of
 * src/framed/compress.rs:242-243 (body: src/raw/compress/mod.rs:165-238), each block from a fresh
 * zeroed table with cursor 0 (independent blocks, src/framed/compress.rs:265-270).
 *
 *   block b reads  d_in[d_in_off[b] .. + d_in_len[b]]
 *   and writes     d_out[d_out_off[b] .. + cap_b], cap_b = d_out_cap ? d_out_cap[b] : d_in_len[b]
 *                  (cap = own plaintext length is the NoPartialWrites bound of compress.rs:242)
 *   d_out_len[b]   bytes written when d_status[b] == LZF_OK
 *   d_status[b]    LZF_OK | LZF_WRITER_FULL (store raw; compress.rs:250-255) | LZF_PANIC
 *   d_xxh_plain    nullable: XXH32(seed 0) of the block's plaintext
 *   d_xxh_stored   nullable: XXH32 of the bytes the frame stores for this block — the compressed
 *                  bytes if LZF_OK, else the plaintext (block checksum, compress.rs:259-263)
 *   hashlog        0 or 12 = reference (src/raw/compress/mod.rs:15); 13..16 = larger-table extension
 *   table_kind     LZF_TABLE_U32 (framed path, compress.rs:202) or LZF_TABLE_U16 (src/lib.rs:26-27)
 *   max_block_len  a promise that no d_in_len[b] exceeds it (the frame's block_size, compress.rs:50,227), or 0 for
 *                  "unknown".  Up to 16 MiB it lets the kernel keep 17-bit table slots (twice as many blocks
 *                  resident per SM); the output is the same either way.  A block that breaks the promise gets
 *                  LZF_PANIC.
 * Output bytes are identical to the reference's for the same input.
 * ------------------------------------------------------------------------------------------ */
int lzf_compress_blocks(lzf_ctx* ctx,
                        const uint8_t* d_in, const uint64_t* d_in_off, const uint32_t* d_in_len,
                        uint32_t nblocks, uint32_t hashlog, uint32_t table_kind, uint32_t max_block_len,
                        uint8_t* d_out, const uint64_t* d_out_off, const uint32_t* d_out_cap,
                        uint32_t* d_out_len, int32_t* d_status,
                        uint32_t* d_xxh_plain, uint32_t* d_xxh_stored, void* stream);

/* ------------------------------------------------------------------------------------------
 * Batched block decompress — replaces `raw::decompress_raw` at src/framed/decompress.rs:247-248
 * (body: src/raw/decompress.rs:58-138) and the stored-block copy at :249-251.
 *
 *   block b reads  d_in[d_in_off[b] .. + (d_in_len[b] & 0x7fffffff)]; if bit 31 of d_in_len[b]
 *                  is set the block is stored and is copied verbatim
 *   prefix         nullable triple: history behind the output (dictionary / carry-over window,
 *                  src/raw/decompress.rs:84-99); block b uses d_prefix[d_prefix_off[b] .. + d_prefix_len[b]]
 *   output         starts empty at d_out[d_out_off[b]], physical capacity d_out_cap[b]
 *   d_out_limit[b] the reference's soft `output_limit` (decompress.rs:55-57,72-74)
 *   d_out_len[b]   decoded length (may exceed d_out_cap[b] with status LZF_OUTPUT_CAP: the
 *                  length the reference's growing Vec would have reached)
 *   d_status[b]    LZF_OK or 1..4 with the reference's check precedence, or LZF_OUTPUT_CAP
 *   d_xxh_plain    nullable: XXH32(seed 0) of the decoded bytes (valid when status == LZF_OK)
 * ------------------------------------------------------------------------------------------ */
int lzf_decompress_blocks(lzf_ctx* ctx,
                          const uint8_t* d_in, const uint64_t* d_in_off, const uint32_t* d_in_len,
                          uint32_t nblocks,
                          const uint8_t* d_prefix, const uint64_t* d_prefix_off, const uint32_t* d_prefix_len,
                          uint8_t* d_out, const uint64_t* d_out_off, const uint32_t* d_out_cap,
                          const uint32_t* d_out_limit, uint32_t* d_out_len, int32_t* d_status,
                          uint32_t* d_xxh_plain, void* stream);

/* Batched XXH32 (seed 0) of device byte ranges, one hash per range (content checksums of whole
 * frames: src/framed/compress.rs:233-235,279-281; src/framed/decompress.rs:276-278,207-211). */
int lzf_xxh32_ranges(lzf_ctx* ctx, const uint8_t* d_data, const uint64_t* d_off, const uint64_t* d_len,
                     uint32_t nranges, uint32_t* d_hash, void* stream);

/* Streaming XXH32 (seed 0) over HOST buffers, for Read/Write-style callers that see the plaintext one
 * block at a time (twox-hash `XxHash32::with_seed(0)` / `Hasher::write` / `finish`,
 * src/framed/compress.rs:172,233-235,279-281; src/framed/decompress.rs:138-139,276-278,207-211).
 * The 16-byte stripes are hashed on the GPU; only the < 16-byte carry and the final avalanche
 * (a dozen integer ops) are host glue. */
typedef struct {
    uint32_t acc[4];
    uint8_t buf[16];
    uint32_t buflen;
    uint64_t total;
} lzf_xxh32_state;
void lzf_xxh32_init(lzf_xxh32_state* st);
int lzf_xxh32_update(lzf_ctx* ctx, lzf_xxh32_state* st, const uint8_t* data, size_t n);
uint32_t lzf_xxh32_finish(const lzf_xxh32_state* st);

/* ------------------------------------------------------------------------------------------
 * Single-block host-pointer conveniences mirroring the Rust raw API.
 *   lzf_raw_compress_into  = compress2(input, 0, &mut fresh table, NoPartialWrites(out[..cap]))
 *                            (the crate has no `compress_into`; see SURVEY.md "Facts")
 *   lzf_raw_decompress     = decompress_raw(input, prefix, &mut Vec (empty), output_limit)
 * Return value: call-level code; *status gets the per-block codec status.
 * ------------------------------------------------------------------------------------------ */
int lzf_raw_compress_into(lzf_ctx* ctx, const uint8_t* in, size_t n, uint32_t table_kind, uint32_t hashlog,
                          uint8_t* out, size_t cap, size_t* written, int32_t* status);
int lzf_raw_decompress(lzf_ctx* ctx, const uint8_t* in, size_t n, const uint8_t* prefix, size_t plen,
                       uint8_t* out, size_t out_cap, size_t out_limit, size_t* out_len, int32_t* status);
/* compress2 with history and a carried table — the full signature of src/raw/compress/mod.rs:165-170,
 *     compress2(input, cursor, &mut table, NoPartialWrites(out[..cap]))
 * as src/framed/compress.rs:220,243,270-275 uses it for dependent blocks: input[..cursor] is match-only history,
 * `table` keeps its entries from call to call.  The table is a device-resident EncoderTable (:19-25):
 *   lzf_table_create   U32Table::default() / U16Table::default()  (:32-36,83-87); hashlog 0/12 = reference
 *   lzf_table_reset    *table = T::default()
 *   lzf_table_offset   EncoderTable::offset(by)  (:72-74): positions of later calls are shifted by `by`
 * *status: LZF_OK, LZF_WRITER_FULL, or LZF_PANIC where the reference would panic (:167 the size assert; :67,92
 * "EncoderTable contract violated" — raised for a position that is actually INSERTED beyond the slot width, i.e. never
 * for the last 7 bytes of the input, and not when the writer refused an earlier sequence first: both orders are
 * reproduced).  A refused write leaves the table as the reference leaves it (updated up to that point); after
 * LZF_PANIC the table is unspecified, as it is behind a Rust panic. */
typedef struct lzf_table lzf_table;
int lzf_table_create(lzf_ctx* ctx, uint32_t table_kind, uint32_t hashlog, lzf_table** table);
void lzf_table_destroy(lzf_ctx* ctx, lzf_table* table);
int lzf_table_reset(lzf_ctx* ctx, lzf_table* table);
int lzf_table_offset(lzf_ctx* ctx, lzf_table* table, uint64_t by);
int lzf_raw_compress2(lzf_ctx* ctx, const uint8_t* in, size_t n, size_t cursor, lzf_table* table,
                      uint8_t* out, size_t cap, size_t* written, int32_t* status);
/* worst-case compress2 output for n input bytes into an unbounded writer */
size_t lzf_compress_bound(size_t n);

/* ------------------------------------------------------------------------------------------
 * Frame layer — CompressionSettings (src/framed/compress.rs:36-157) as a POD.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t independent_blocks;     /* default 1  compress.rs:47 */
    int32_t block_checksums;        /* default 0  compress.rs:48 */
    int32_t content_checksum;       /* default 1  compress.rs:49 */
    uint64_t block_size;            /* default 4 MiB; 64 KiB/256 KiB/1 MiB/4 MiB  compress.rs:50, header.rs:73-80 */
    const uint8_t* dictionary;      /* nullable   compress.rs:113-117 */
    uint64_t dictionary_len;
    int32_t has_dictionary_id;      /* dictionary() sets both; dictionary_id_nonsense_override() only this */
    uint32_t dictionary_id;
    int32_t has_content_size;       /* compress_with_size / compress_with_size_unchecked  compress.rs:142-157 */
    uint64_t content_size;
    uint32_t hashlog;               /* 0/12 = reference */
} lzf_settings;

void lzf_settings_default(lzf_settings* s);
/* upper bound of a frame's size for n plaintext bytes under `s` */
size_t lzf_frame_bound(const lzf_settings* s, size_t n);

/* CompressionSettings::compress / compress_with_size* over host buffers
 * (compress_internal, src/framed/compress.rs:159-282).  *status gets an LZF_F_* code. */
int lzf_frame_compress(lzf_ctx* ctx, const lzf_settings* s, const uint8_t* in, size_t n,
                       uint8_t* out, size_t cap, size_t* written, int32_t* status);

/* Many frames in one call (host buffers): frame f = in[in_off[f] .. + in_len[f]] is written to
 * out[out_off[f] ..] (capacity out_cap[f]); out_len[f]/status[f] per frame.  All blocks of all
 * frames go through ONE compress launch. */
int lzf_frames_compress(lzf_ctx* ctx, const lzf_settings* s, const uint8_t* in, const uint64_t* in_off,
                        const uint64_t* in_len, uint32_t nframes, uint8_t* out, const uint64_t* out_off,
                        const uint64_t* out_cap, uint64_t* out_len, int32_t* status);

/* Same, device-resident plaintext and frames; offsets/lengths/status are HOST arrays. Synchronises. */
int lzf_frames_compress_device(lzf_ctx* ctx, const lzf_settings* s, const uint8_t* d_in, const uint64_t* in_off,
                               const uint64_t* in_len, uint32_t nframes, uint8_t* d_out, const uint64_t* out_off,
                               const uint64_t* out_cap, uint64_t* out_len, int32_t* status);

/* LZ4FrameReader::new (src/framed/decompress.rs:101-161) */
typedef struct {
    uint8_t flags;                  /* Flags bits, header.rs:8-16 */
    uint64_t block_maxsize;         /* LZ4FrameReader::block_size() */
    int32_t has_content_size;
    uint64_t content_size;          /* LZ4FrameReader::frame_size() */
    int32_t has_dictionary_id;
    uint32_t dictionary_id;         /* LZ4FrameReader::dictionary_id() */
    size_t header_len;
} lzf_frame_info;
int lzf_frame_parse_header(const uint8_t* in, size_t n, lzf_frame_info* info, int32_t* detail);

/* decompress_frame / into_read_with_dictionary(..).read_to_end (src/framed/decompress.rs:180-288)
 * over host buffers.  *status = LZF_F_*, *detail = codec status / ParseError kind, *written =
 * plaintext bytes of the blocks decoded before a failure (what read_to_end had appended),
 * *consumed = bytes of `in` the reader consumed. */
int lzf_frame_decompress(lzf_ctx* ctx, const uint8_t* in, size_t n, const uint8_t* dict, size_t dlen,
                         uint8_t* out, size_t cap, size_t* written, size_t* consumed,
                         int32_t* status, int32_t* detail);

/* Many frames in one call (host buffers); all blocks of all frames go through ONE decompress launch. */
int lzf_frames_decompress(lzf_ctx* ctx, const uint8_t* in, const uint64_t* in_off, const uint64_t* in_len,
                          uint32_t nframes, uint8_t* out, const uint64_t* out_off, const uint64_t* out_cap,
                          uint64_t* out_len, int32_t* status, int32_t* detail);

/* Same, device-resident frames and plaintext; offsets/lengths/status are HOST arrays. Synchronises. */
int lzf_frames_decompress_device(lzf_ctx* ctx, const uint8_t* d_in, const uint64_t* in_off, const uint64_t* in_len,
                                 uint32_t nframes, uint8_t* d_out, const uint64_t* out_off, const uint64_t* out_cap,
                                 uint64_t* out_len, int32_t* status, int32_t* detail);

#ifdef __cplusplus
}
#endif
#endif
