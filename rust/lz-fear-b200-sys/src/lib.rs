//! Raw FFI declarations of `liblzfear_b200.so` (include/lzfear_b200.h): the B200 LZ4 block codec that stands in for
//! lz-fear's `raw::compress2` / `raw::decompress_raw` and the per-block loops of its framed layer.
//!
//! GENERATED from the C header by rust/gen_sys.py — edit the header, not this file.  Everything here is `unsafe`;
//! the safe surface lives in the sibling crate `lz-fear-b200`.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const LZF_SUCCESS: i32 = 0;
pub const LZF_ERR_INVALID_ARG: i32 = -1;
pub const LZF_ERR_CUDA: i32 = -2;
pub const LZF_ERR_NO_DEVICE: i32 = -3;
pub const LZF_ERR_OOM: i32 = -4;
pub const LZF_ERR_UNSUPPORTED: i32 = -5;
pub const LZF_OK: i32 = 0;
pub const LZF_UNEXPECTED_END: i32 = 1;
pub const LZF_MEMORY_LIMIT_EXCEEDED: i32 = 2;
pub const LZF_ZERO_DEDUP_OFFSET: i32 = 3;
pub const LZF_INVALID_DEDUP_OFFSET: i32 = 4;
pub const LZF_WRITER_FULL: i32 = 5;
pub const LZF_OUTPUT_CAP: i32 = 6;
pub const LZF_PANIC: i32 = 7;
pub const LZF_F_OK: i32 = 0;
pub const LZF_F_INPUT_ERROR: i32 = 10;
pub const LZF_F_CODEC_ERROR: i32 = 11;
pub const LZF_F_HEADER_PARSE_ERROR: i32 = 12;
pub const LZF_F_WRONG_MAGIC: i32 = 13;
pub const LZF_F_HEADER_CHECKSUM_FAIL: i32 = 14;
pub const LZF_F_BLOCK_CHECKSUM_FAIL: i32 = 15;
pub const LZF_F_FRAME_CHECKSUM_FAIL: i32 = 16;
pub const LZF_F_BLOCK_LENGTH_OVERFLOW: i32 = 17;
pub const LZF_F_BLOCK_SIZE_OVERFLOW: i32 = 18;
pub const LZF_F_INVALID_BLOCK_SIZE: i32 = 20;
pub const LZF_F_WRITE_ERROR: i32 = 21;
pub const LZF_F_PANIC: i32 = 22;
pub const LZF_P_UNIMPLEMENTED_BLOCKSIZE: i32 = 1;
pub const LZF_P_UNSUPPORTED_VERSION: i32 = 2;
pub const LZF_P_RESERVED_FLAG_BITS: i32 = 3;
pub const LZF_P_RESERVED_BD_BITS: i32 = 4;
pub const LZF_TABLE_U32: i32 = 0;
pub const LZF_TABLE_U16: i32 = 1;
pub const LZF_OPT_SEGMENT_BYTES: i32 = 1;
pub const LZF_ABI_VERSION: u32 = 3;
pub const LZF_INCOMPRESSIBLE: u32 = 0x80000000;
pub const LZF_MAGIC: u32 = 0x184D2204;
pub const LZF_WINDOW_SIZE: u32 = 0x10000;

#[repr(C)]
pub struct lzf_ctx { _private: [u8; 0] }
#[repr(C)]
pub struct lzf_table { _private: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy)]
pub struct lzf_xxh32_state {
    pub acc: [u32; 4],
    pub buf: [u8; 16],
    pub buflen: u32,
    pub total: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct lzf_settings {
    pub independent_blocks: i32,
    pub block_checksums: i32,
    pub content_checksum: i32,
    pub block_size: u64,
    pub dictionary: *const u8,
    pub dictionary_len: u64,
    pub has_dictionary_id: i32,
    pub dictionary_id: u32,
    pub has_content_size: i32,
    pub content_size: u64,
    pub hashlog: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct lzf_frame_info {
    pub flags: u8,
    pub block_maxsize: u64,
    pub has_content_size: i32,
    pub content_size: u64,
    pub has_dictionary_id: i32,
    pub dictionary_id: u32,
    pub header_len: usize,
}

#[link(name = "lzfear_b200")]
extern "C" {
    pub fn lzf_abi_version() -> c_int;
    pub fn lzf_create(device: c_int, ctx: *mut *mut lzf_ctx) -> c_int;
    pub fn lzf_destroy(ctx: *mut lzf_ctx);
    pub fn lzf_last_error(ctx: *const lzf_ctx) -> *const c_char;
    pub fn lzf_set_option(ctx: *mut lzf_ctx, option: c_int, value: u64) -> c_int;
    pub fn lzf_trim(ctx: *mut lzf_ctx) -> c_int;
    pub fn lzf_launch_count(ctx: *const lzf_ctx) -> u64;
    pub fn lzf_compress_blocks(ctx: *mut lzf_ctx, d_in: *const u8, d_in_off: *const u64, d_in_len: *const u32, nblocks: u32, hashlog: u32, table_kind: u32, max_block_len: u32, d_out: *mut u8, d_out_off: *const u64, d_out_cap: *const u32, d_out_len: *mut u32, d_status: *mut i32, d_xxh_plain: *mut u32, d_xxh_stored: *mut u32, stream: *mut c_void) -> c_int;
    pub fn lzf_decompress_blocks(ctx: *mut lzf_ctx, d_in: *const u8, d_in_off: *const u64, d_in_len: *const u32, nblocks: u32, d_prefix: *const u8, d_prefix_off: *const u64, d_prefix_len: *const u32, d_out: *mut u8, d_out_off: *const u64, d_out_cap: *const u32, d_out_limit: *const u32, d_out_len: *mut u32, d_status: *mut i32, d_xxh_plain: *mut u32, stream: *mut c_void) -> c_int;
    pub fn lzf_xxh32_ranges(ctx: *mut lzf_ctx, d_data: *const u8, d_off: *const u64, d_len: *const u64, nranges: u32, d_hash: *mut u32, stream: *mut c_void) -> c_int;
    pub fn lzf_xxh32_init(st: *mut lzf_xxh32_state);
    pub fn lzf_xxh32_update(ctx: *mut lzf_ctx, st: *mut lzf_xxh32_state, data: *const u8, n: usize) -> c_int;
    pub fn lzf_xxh32_finish(st: *const lzf_xxh32_state) -> u32;
    pub fn lzf_raw_compress_into(ctx: *mut lzf_ctx, input: *const u8, n: usize, table_kind: u32, hashlog: u32, out: *mut u8, cap: usize, written: *mut usize, status: *mut i32) -> c_int;
    pub fn lzf_raw_decompress(ctx: *mut lzf_ctx, input: *const u8, n: usize, prefix: *const u8, plen: usize, out: *mut u8, out_cap: usize, out_limit: usize, out_len: *mut usize, status: *mut i32) -> c_int;
    pub fn lzf_table_create(ctx: *mut lzf_ctx, table_kind: u32, hashlog: u32, table: *mut *mut lzf_table) -> c_int;
    pub fn lzf_table_destroy(ctx: *mut lzf_ctx, table: *mut lzf_table);
    pub fn lzf_table_reset(ctx: *mut lzf_ctx, table: *mut lzf_table) -> c_int;
    pub fn lzf_table_offset(ctx: *mut lzf_ctx, table: *mut lzf_table, by: u64) -> c_int;
    pub fn lzf_raw_compress2(ctx: *mut lzf_ctx, input: *const u8, n: usize, cursor: usize, table: *mut lzf_table, out: *mut u8, cap: usize, written: *mut usize, status: *mut i32) -> c_int;
    pub fn lzf_compress_bound(n: usize) -> usize;
    pub fn lzf_settings_default(s: *mut lzf_settings);
    pub fn lzf_frame_bound(s: *const lzf_settings, n: usize) -> usize;
    pub fn lzf_frame_compress(ctx: *mut lzf_ctx, s: *const lzf_settings, input: *const u8, n: usize, out: *mut u8, cap: usize, written: *mut usize, status: *mut i32) -> c_int;
    pub fn lzf_frames_compress(ctx: *mut lzf_ctx, s: *const lzf_settings, input: *const u8, in_off: *const u64, in_len: *const u64, nframes: u32, out: *mut u8, out_off: *const u64, out_cap: *const u64, out_len: *mut u64, status: *mut i32) -> c_int;
    pub fn lzf_frames_compress_device(ctx: *mut lzf_ctx, s: *const lzf_settings, d_in: *const u8, in_off: *const u64, in_len: *const u64, nframes: u32, d_out: *mut u8, out_off: *const u64, out_cap: *const u64, out_len: *mut u64, status: *mut i32) -> c_int;
    pub fn lzf_frame_parse_header(input: *const u8, n: usize, info: *mut lzf_frame_info, detail: *mut i32) -> c_int;
    pub fn lzf_frame_decompress(ctx: *mut lzf_ctx, input: *const u8, n: usize, dict: *const u8, dlen: usize, out: *mut u8, cap: usize, written: *mut usize, consumed: *mut usize, status: *mut i32, detail: *mut i32) -> c_int;
    pub fn lzf_frames_decompress(ctx: *mut lzf_ctx, input: *const u8, in_off: *const u64, in_len: *const u64, nframes: u32, out: *mut u8, out_off: *const u64, out_cap: *const u64, out_len: *mut u64, status: *mut i32, detail: *mut i32) -> c_int;
    pub fn lzf_frames_decompress_device(ctx: *mut lzf_ctx, d_in: *const u8, in_off: *const u64, in_len: *const u64, nframes: u32, d_out: *mut u8, out_off: *const u64, out_cap: *const u64, out_len: *mut u64, status: *mut i32, detail: *mut i32) -> c_int;
}
