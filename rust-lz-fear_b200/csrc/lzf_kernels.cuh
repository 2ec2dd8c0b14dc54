// lzf_kernels.cuh — argument blocks and host-side launchers of the sm_100a kernels
// (lzf_compress.cu, lzf_decompress.cu, lzf_frame.cu), shared with the C-ABI layer (lzf_api.cu).
#pragma once

#include "lzf_common.cuh"

namespace lzf {
struct EncodeArgs {
    const uint8_t* in; const uint64_t* in_off; const uint32_t* in_len; uint32_t nblocks;
    uint32_t hashlog; uint32_t table_kind;
    uint8_t* out; const uint64_t* out_off; const uint32_t* out_cap;
    uint32_t* out_len; int32_t* status; uint32_t* xxh_plain; uint32_t* xxh_stored;
    uint32_t* work_counter; uint8_t* global_tables; uint32_t max_block_len;
    // internal (frame layer): dependent blocks and dictionaries (src/framed/compress.rs:202-214,220,271-275).
    //   prefix_len[b]  bytes of addressable history physically in front of block b (window / dictionary):
    //                  compress2's `cursor` (window_offset, compress.rs:222,243)
    //   abs_base[b]    stream position of the first history byte (the table stores stream positions, :65)
    //   prime_len[b]   dictionary bytes at the start of b's history whose positions 0, 3, 6, ... are
    //                  inserted before the block is parsed (compress.rs:204-214); first block of a chain only
    //   chains         consecutive blocks [chain_first[c], + chain_count[c]) share one table and are parsed in
    //                  order by one warp; null = every block is its own chain with a fresh table
    const uint32_t* prefix_len; const uint32_t* abs_base; const uint32_t* prime_len;
    const uint32_t* chain_first; const uint32_t* chain_count; uint32_t nchains;
    uint64_t max_pos;   // largest stream position + 1 any block reaches (0 = max_block_len): picks the slot width
    // internal (host-buffer pipeline): the plaintext is still arriving over PCIe while the kernel runs.  Bytes
    // [0, *progress * slice_bytes) of EVERY block are in place (the host copies one slice of all blocks after the
    // other and bumps *progress after each); a warp waits before it reads past that.  null = everything is there.
    const uint32_t* progress; uint32_t slice_bytes;
    // internal (raw API with a carried table, lzf_raw_compress2): the first chain starts from these slots (one u32
    // stream position per slot) instead of a zeroed table and writes them back when its last block is done
    uint32_t* table_io;
    // tuning knobs, read from the environment ONCE by lzf_create (0 = default): plain u32 slots where packed ones
    // would do / 1 + the number of warps per CTA whose table may live in shared memory
    uint32_t tune_u32_slots; uint32_t tune_smem_warps_p1;
    // internal (segmented parse): where the closing literal-only sequence of block b starts in its output, and how
    // many literals it holds (null = not wanted)
    uint32_t* fin_pos; uint32_t* fin_lit;
    // internal (lzf_raw_compress2 next to the table's position limit): parse on even though stream positions leave the
    // slot width (they wrap); the host decides from the sequence stream where the reference's expect() fires
    uint32_t allow_slot_wrap;
};
// Segmented parse (lzf_set_option LZF_OPT_SEGMENT_BYTES): a launch with too few blocks to fill the GPU cuts every block
// into S segments that are parsed side by side, each from a table primed with the 64 KiB in front of it, and stitches
// the S sequence streams back into one LZ4 block.  Valid LZ4 of (almost) the reference's size — not its bytes.
struct SegmentPlanArgs {
    uint32_t nblocks, nseg, seg_len, seg_cap;            // S segments per block of seg_len bytes, seg_cap bytes of scratch each
    const uint64_t* in_off; const uint32_t* in_len;      // the blocks
    uint64_t* seg_in_off; uint32_t* seg_in_len; uint32_t* seg_prefix; uint64_t* seg_out_off; uint32_t* seg_out_cap;
    uint32_t* seg_chain_first; uint32_t* seg_chain_count; uint32_t* seg_abs;
};
struct StitchJob { uint64_t lit_src, lit_dst, body_src, body_dst; uint32_t lit_len, body_len; };
struct StitchArgs {
    uint32_t nblocks, nseg, seg_len;
    const uint8_t* in; const uint64_t* in_off; const uint32_t* in_len;                   // plaintext blocks
    const uint8_t* seg; const uint64_t* seg_out_off; const uint32_t* seg_out_len;        // per-segment streams
    const int32_t* seg_status; const uint32_t* fin_pos; const uint32_t* fin_lit;
    uint8_t* out; const uint64_t* out_off; const uint32_t* out_cap;                      // the blocks' outputs
    uint32_t* out_len; int32_t* status;
    StitchJob* jobs;                                                                     // nblocks * (nseg + 1)
    uint64_t* hash_off; uint64_t* hash_len; uint64_t* plain_off; uint64_t* plain_len;    // nullable: absolute ranges for the block checksums
};
struct StageArgs {      // [dictionary | block] staging copies for blocks whose history is the dictionary
    uint32_t n; const uint8_t* dict; uint32_t dlen;
    const uint8_t* in; const uint64_t* src_off; const uint32_t* len; uint8_t* dst; const uint64_t* dst_off;
};
struct DecodeArgs {
    const uint8_t* in; const uint64_t* in_off; const uint32_t* in_len; uint32_t nblocks;
    const uint8_t* prefix; const uint64_t* prefix_off; const uint32_t* prefix_len;
    uint8_t* out; const uint64_t* out_off; const uint32_t* out_cap; const uint32_t* out_limit;
    uint32_t* out_len; int32_t* status; uint32_t* xxh_plain;
    uint32_t* work_counter;
    // internal (frame layer): dependent blocks.  prefix_abs: prefix_off[] holds absolute device addresses
    // (history of a dependent block = the previous blocks' output in front of its own).  wait_for[b] >= 0:
    // block b may only start once block wait_for[b] (always a lower index) has set done[wait_for[b]].
    int prefix_abs; const int32_t* wait_for; uint32_t* done;
    uint32_t tune_ctas_per_sm;      // tuning knob read once by lzf_create: cap of the resident CTAs per SM (0 = occupancy)
};
struct LayoutArgs {
    uint32_t nframes;
    const uint32_t* first_block; const uint32_t* nblocks;
    const uint32_t* blk_in_len; const uint32_t* blk_comp_len; const int32_t* blk_status;
    int block_checksums; int content_checksum;
    const uint8_t* headers;
    uint8_t* out; const uint64_t* out_off; const uint64_t* out_cap;
    const uint32_t* content_hash;
    uint64_t* blk_dst;
    uint64_t* frame_len; int32_t* frame_status;
};
struct AssembleArgs {
    uint32_t nblocks;
    const uint8_t* in; const uint64_t* blk_in_off; const uint32_t* blk_in_len;
    const uint8_t* comp; const uint64_t* blk_comp_off; const uint32_t* blk_comp_len; const int32_t* blk_status;
    const uint32_t* blk_xxh_stored;
    const uint64_t* blk_dst; uint8_t* out;
};
struct WalkFrame {
    int32_t header_status; int32_t header_detail;
    uint32_t flags; uint32_t nblocks;
    uint64_t block_maxsize;
    int32_t term_status;
    uint32_t content_checksum;
    uint64_t consumed;
    uint64_t content_size; uint32_t dictionary_id; uint32_t has_fields;
};
struct WalkArgs {
    uint32_t nframes; int mode;
    const uint8_t* in; const uint64_t* in_off; const uint64_t* in_len;
    WalkFrame* frames;
    const uint32_t* first_block;
    const uint64_t* out_off; const uint64_t* out_cap;
    uint64_t* blk_in_off; uint32_t* blk_len_word; uint32_t* blk_checksum;
    uint64_t* blk_out_off; uint32_t* blk_out_cap; uint32_t* blk_out_limit;
    uint64_t* blk_payload_len; uint64_t* blk_end;
};
}  // namespace lzf

extern "C" {
int lzf_launch_encode(const lzf::EncodeArgs* args, int num_sms, cudaStream_t stream);
size_t lzf_encode_global_table_bytes(const lzf::EncodeArgs* args, int num_sms);
int lzf_launch_decode(const lzf::DecodeArgs* args, int num_sms, cudaStream_t stream);
int lzf_launch_xxh32_ranges(const uint8_t* data, const uint64_t* off, const uint64_t* len, uint32_t nranges,
                            uint32_t* hash, cudaStream_t s, int long_ranges = 0);
int lzf_launch_xxh32_stripes(const uint8_t* data, uint64_t nstripes, uint32_t* acc, cudaStream_t s);
int lzf_launch_stage_dict(const lzf::StageArgs* a, uint32_t max_block_len, cudaStream_t s);
int lzf_launch_layout(const lzf::LayoutArgs* a, cudaStream_t s);
int lzf_launch_assemble(const lzf::AssembleArgs* a, uint32_t max_block_len, cudaStream_t s);
int lzf_launch_walk(const lzf::WalkArgs* a, cudaStream_t s);
int lzf_launch_segment_plan(const lzf::SegmentPlanArgs* a, cudaStream_t s);
int lzf_launch_stitch(const lzf::StitchArgs* a, cudaStream_t s);
}

