// lzf_frame.cuh — frame-format pieces shared by host shim and device kernels.
// Header layout and checks follow src/framed/header.rs and LZ4FrameReader::new
// (src/framed/decompress.rs:101-161).
#pragma once

#include <stdint.h>
#include <stddef.h>

#include "../../include/lzfear_b200.h"

#ifdef __CUDACC__
#define LZF_HD __host__ __device__
#else
#define LZF_HD
#endif

namespace lzf {

constexpr uint8_t kFlagIndependent = 0x20;      // header.rs:10
constexpr uint8_t kFlagBlockChecksums = 0x10;   // header.rs:11
constexpr uint8_t kFlagContentSize = 0x08;      // header.rs:12
constexpr uint8_t kFlagContentChecksum = 0x04;  // header.rs:13
constexpr uint8_t kFlagDictionaryId = 0x01;     // header.rs:14
constexpr uint32_t kMaxHeaderLen = 19;

// Scalar XXH32 (seed 0) for the handful of header bytes (compress.rs:197-199, decompress.rs:112-133).
// Bulk hashing (block / content checksums) is done by the warp kernels, not by this function.
LZF_HD inline uint32_t xxh32_scalar(const uint8_t* p, size_t n) {
    const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    auto rotl = [](uint32_t x, int r) { return (x << r) | (x >> (32 - r)); };
    auto rd32 = [](const uint8_t* q) {
        return uint32_t(q[0]) | (uint32_t(q[1]) << 8) | (uint32_t(q[2]) << 16) | (uint32_t(q[3]) << 24);
    };
    const uint8_t* end = p + n;
    uint32_t h;
    if (n >= 16) {
        uint32_t a0 = P1 + P2, a1 = P2, a2 = 0, a3 = 0u - P1;
        while (end - p >= 16) {
            a0 = rotl(a0 + rd32(p) * P2, 13) * P1;
            a1 = rotl(a1 + rd32(p + 4) * P2, 13) * P1;
            a2 = rotl(a2 + rd32(p + 8) * P2, 13) * P1;
            a3 = rotl(a3 + rd32(p + 12) * P2, 13) * P1;
            p += 16;
        }
        h = rotl(a0, 1) + rotl(a1, 7) + rotl(a2, 12) + rotl(a3, 18);
    } else {
        h = P5;
    }
    h += (uint32_t)n;
    while (end - p >= 4) { h = rotl(h + rd32(p) * P3, 17) * P4; p += 4; }
    while (p < end) { h = rotl(h + uint32_t(*p) * P5, 11) * P1; p++; }
    h ^= h >> 15; h *= P2;
    h ^= h >> 13; h *= P3;
    h ^= h >> 16;
    return h;
}

// BlockDescriptor::block_maxsize — header.rs:72-80.  0 on success, else LZF_P_*.
LZF_HD inline int bd_block_maxsize(uint8_t bd, uint64_t* size) {
    const unsigned s = (bd >> 4) & 7u;
    if (s >= 4 && s < 8) { *size = uint64_t(1) << (s * 2 + 8); return 0; }
    return LZF_P_UNIMPLEMENTED_BLOCKSIZE;
}

// BlockDescriptor::new — header.rs:53-62.  0 = Some(bd), 1 = None, 2 = panic (unwrap at :55).
inline int bd_new(uint64_t block_maxsize, uint8_t* bd) {
    unsigned tz = 64;
    if (block_maxsize) { tz = 0; while (!((block_maxsize >> tz) & 1)) tz++; }
    const unsigned maybe = ((tz > 8 ? tz - 8 : 0) / 2) & 0xff;
    const uint8_t b = (uint8_t)(maybe << 4);
    if (b & 0x8f) return 2;
    uint64_t sz;
    if (bd_block_maxsize(b, &sz) || sz != block_maxsize) return 1;
    *bd = b;
    return 0;
}

// LZ4FrameReader::new — decompress.rs:101-161.  Returns LZF_F_*; *detail = ParseError kind.
LZF_HD inline int parse_frame_header(const uint8_t* in, size_t n, lzf_frame_info* info, int32_t* detail) {
    size_t p = 0;
    *detail = 0;
    info->flags = 0; info->block_maxsize = 0; info->has_content_size = 0; info->content_size = 0;
    info->has_dictionary_id = 0; info->dictionary_id = 0; info->header_len = 0;
    if (n - p < 4) return LZF_F_INPUT_ERROR;                                   // :103
    const uint32_t magic = uint32_t(in[0]) | (uint32_t(in[1]) << 8) | (uint32_t(in[2]) << 16) | (uint32_t(in[3]) << 24);
    p = 4;
    if (magic != LZF_MAGIC) return LZF_F_WRONG_MAGIC;                          // :104-106
    if (n - p < 1) return LZF_F_INPUT_ERROR;
    const uint8_t flags_byte = in[p++];                                        // :108
    if ((flags_byte >> 6) != 1) { *detail = LZF_P_UNSUPPORTED_VERSION; return LZF_F_HEADER_PARSE_ERROR; }   // header.rs:33-36
    if (flags_byte & 2) { *detail = LZF_P_RESERVED_FLAG_BITS; return LZF_F_HEADER_PARSE_ERROR; }           // header.rs:37-39
    const uint8_t flags = flags_byte & (kFlagIndependent | kFlagBlockChecksums | kFlagContentSize |
                                        kFlagContentChecksum | kFlagDictionaryId);
    if (n - p < 1) return LZF_F_INPUT_ERROR;
    const uint8_t bd = in[p++];                                                // :110
    if (bd & 0x8f) { *detail = LZF_P_RESERVED_BD_BITS; return LZF_F_HEADER_PARSE_ERROR; }                  // header.rs:65-68
    const size_t hashed_from = 4;
    if (flags & kFlagContentSize) {                                            // :116-122
        if (n - p < 8) return LZF_F_INPUT_ERROR;
        uint64_t v = 0;
        for (int i = 7; i >= 0; i--) v = (v << 8) | in[p + i];
        info->content_size = v;
        info->has_content_size = 1;
        p += 8;
    }
    if (flags & kFlagDictionaryId) {                                           // :124-130
        if (n - p < 4) return LZF_F_INPUT_ERROR;
        info->dictionary_id = uint32_t(in[p]) | (uint32_t(in[p + 1]) << 8) | (uint32_t(in[p + 2]) << 16) | (uint32_t(in[p + 3]) << 24);
        info->has_dictionary_id = 1;
        p += 4;
    }
    if (n - p < 1) return LZF_F_INPUT_ERROR;
    const uint8_t want = in[p];                                                // :132
    const uint8_t have = (uint8_t)(xxh32_scalar(in + hashed_from, p - hashed_from) >> 8);   // :133
    p++;
    if (want != have) return LZF_F_HEADER_CHECKSUM_FAIL;                       // :134-136
    uint64_t bms = 0;
    const int pe = bd_block_maxsize(bd, &bms);                                 // :153
    if (pe) { *detail = pe; return LZF_F_HEADER_PARSE_ERROR; }
    info->flags = flags;
    info->block_maxsize = bms;
    info->header_len = p;
    return LZF_F_OK;
}

}  // namespace lzf
