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
size_t lzf_encode_global_table_warps(int num_sms);
int lzf_launch_decode(const lzf::DecodeArgs* args, int num_sms, cudaStream_t stream);
int lzf_launch_xxh32_ranges(const uint8_t* data, const uint64_t* off, const uint64_t* len, uint32_t nranges,
                            uint32_t* hash, cudaStream_t s);
int lzf_launch_xxh32_stripes(const uint8_t* data, uint64_t nstripes, uint32_t* acc, cudaStream_t s);
int lzf_launch_layout(const lzf::LayoutArgs* a, cudaStream_t s);
int lzf_launch_assemble(const lzf::AssembleArgs* a, uint32_t max_block_len, cudaStream_t s);
int lzf_launch_walk(const lzf::WalkArgs* a, cudaStream_t s);
}

