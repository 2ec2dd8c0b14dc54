// simt_kernels.cpp — TEST INFRASTRUCTURE ONLY: compiles the product kernel sources for the CPU
// SIMT emulator (simt_emu.h) and exposes them to pytest through a C ABI with HOST pointers.
#include "simt_emu.h"

#include "../../rust-lz-fear_b200/csrc/lzf_compress.cu"
#include "../../rust-lz-fear_b200/csrc/lzf_decompress.cu"
#include "../../rust-lz-fear_b200/csrc/lzf_frame.cu"
#include "../../rust-lz-fear_b200/csrc/lzf_api.cu"

extern "C" {

int simt_encode_blocks(const uint8_t* in, const uint64_t* in_off, const uint32_t* in_len, uint32_t nblocks,
                       uint32_t hashlog, uint32_t table_kind, uint8_t* out, const uint64_t* out_off,
                       const uint32_t* out_cap, uint32_t* out_len, int32_t* status, uint32_t* xxh_plain,
                       uint32_t* xxh_stored, uint32_t max_block_len, int num_sms) {
    lzf::EncodeArgs a;
    memset(&a, 0, sizeof(a));
    a.in = in; a.in_off = in_off; a.in_len = in_len; a.nblocks = nblocks;
    a.hashlog = hashlog ? hashlog : 12; a.table_kind = table_kind;
    a.out = out; a.out_off = out_off; a.out_cap = out_cap; a.out_len = out_len; a.status = status;
    a.xxh_plain = xxh_plain; a.xxh_stored = xxh_stored;
    uint32_t counter = 0;
    a.work_counter = &counter;
    a.max_block_len = max_block_len;
    std::vector<uint8_t> gt(lzf_encode_global_table_bytes(&a, num_sms) + 16);
    a.global_tables = gt.data();
    return lzf_launch_encode(&a, num_sms, nullptr);
}

int simt_decode_blocks(const uint8_t* in, const uint64_t* in_off, const uint32_t* in_len, uint32_t nblocks,
                       const uint8_t* prefix, const uint64_t* prefix_off, const uint32_t* prefix_len,
                       uint8_t* out, const uint64_t* out_off, const uint32_t* out_cap, const uint32_t* out_limit,
                       uint32_t* out_len, int32_t* status, uint32_t* xxh_plain) {
    lzf::DecodeArgs a;
    memset(&a, 0, sizeof(a));
    a.in = in; a.in_off = in_off; a.in_len = in_len; a.nblocks = nblocks;
    a.prefix = prefix; a.prefix_off = prefix_off; a.prefix_len = prefix_len;
    a.out = out; a.out_off = out_off; a.out_cap = out_cap; a.out_limit = out_limit;
    a.out_len = out_len; a.status = status; a.xxh_plain = xxh_plain;
    uint32_t counter = 0;
    a.work_counter = &counter;
    return lzf_launch_decode(&a, 2, nullptr);
}

int simt_xxh32_ranges(const uint8_t* data, const uint64_t* off, const uint64_t* len, uint32_t n, uint32_t* hash) {
    return lzf_launch_xxh32_ranges(data, off, len, n, hash, nullptr);
}

}  // extern "C"
