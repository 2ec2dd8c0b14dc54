// lzf_compress.cu — batched LZ4 block compress for sm_100a, one warp per independent block.
//
// Behavioural contract: raw::compress2 of the reference (src/raw/compress/mod.rs:165-238) with a
// fresh zeroed EncoderTable and cursor 0, writing through NoPartialWrites(out[..cap])
// (src/framed/compress.rs:242,294-308).  The emitted bytes are IDENTICAL to the reference's.
//
// The reference parse is one serial chain per block (hash -> table swap -> candidate compare).
// We keep its exact semantics but evaluate it 32 probes at a time ("speculate 32 probes, commit
// the prefix"): within a literal run the k-th future probe position is a closed-form function of
// the run start (the step/step_counter recurrence of :174-175,225-231), so lane k hashes probe k,
// reads its table slot, and — when a lower lane of the same batch hits the same slot — takes that
// lane's position as its candidate, exactly what the serial mem::swap (:68) would have left
// there.  A ballot picks the first lane whose candidate is a real >= 4-byte match (or that hits
// the end-of-block rule :178); lanes up to the winner commit their table writes in lane order,
// later lanes are discarded.  Forward/backward match extension (:117-145, :211-214) and the
// sequence emit (:150-163, :239-260) are warp-parallel.  The per-warp hash table lives in shared
// memory (16 KiB for the reference's 4096 x u32; 8 KiB when every position fits u16).
#include "lzf_kernels.cuh"

namespace lzf {


// offset of the j-th probe of a literal run from the run start: the closed form of
//   cursor += step; step = step_counter >> 6; if literal_start + 1 != cursor { step_counter += 1 }
// (src/raw/compress/mod.rs:174-175,225-231) with step_counter starting at 64 and step at 1.
__device__ __forceinline__ uint64_t probe_offset(uint32_t j) {
    if (j < 2) return j;
    const uint64_t t = 62ull + j;
    const uint64_t q = t >> 6, r = t & 63;
    return 2 + 32 * q * (q - 1) + r * q;
}

// hash_for_u32, 64-bit little-endian branch (:40-51): ((v << 24) * 889523592379) >> (64 - hashlog)
__device__ __forceinline__ uint32_t hash5(uint64_t v, uint32_t hashlog) {
    return (uint32_t)(((v << 24) * 889523592379ull) >> (64 - hashlog));
}
// hash_for_u16 (:58-61): one more bit than hashlog because the u16 table has twice the slots
__device__ __forceinline__ uint32_t hash4(uint32_t v, uint32_t hashlog) {
    return (v * 2654435761u) >> (32 - hashlog - 1);
}

// bytes write_lsic_tail (:243-260) emits for `value`
__device__ __forceinline__ uint64_t lsic_len(uint64_t value) {
    return value < 15 ? 0 : (value - 15) / 255 + 1;
}
// warp-parallel write_lsic_tail
__device__ __forceinline__ void write_lsic(uint8_t* dst, uint64_t value) {
    if (value < 15) return;
    const uint64_t nbytes = (value - 15) / 255 + 1;
    const uint8_t last = (uint8_t)((value - 15) % 255);
    for (uint64_t i = lane_id(); i < nbytes; i += 32) dst[i] = (i == nbytes - 1) ? last : 0xff;
}

constexpr int kEncodeWarpsPerCta = 4;
constexpr int kGlobalTableCtasPerSm = 4;   // bounds the global table scratch

template <typename Slot, bool kHash4>
__global__ void __launch_bounds__(kEncodeWarpsPerCta * 32)
encode_blocks_kernel(EncodeArgs a, uint32_t nslots, int smem_tables) {
    LZF_DYN_SMEM(smem_raw);
    const unsigned lane = lane_id();
    const unsigned warp_in_cta = threadIdx.x >> 5;
    Slot* table;
    if (smem_tables) {
        table = reinterpret_cast<Slot*>(smem_raw) + (size_t)warp_in_cta * nslots;
    } else {
        const size_t gw = (size_t)blockIdx.x * kEncodeWarpsPerCta + warp_in_cta;
        table = reinterpret_cast<Slot*>(a.global_tables) + gw * nslots;
    }
    const uint32_t hashlog = a.hashlog;

    for (;;) {
        uint32_t b = 0;
        if (lane == 0) b = atomicAdd(a.work_counter, 1u);
        b = __shfl_sync(LZF_FULL_MASK, b, 0);
        if (b >= a.nblocks) break;

        const uint64_t len = a.in_len[b];
        const uint8_t* in = a.in + a.in_off[b];
        const uint8_t* in_end = in + len;
        uint8_t* out = a.out + a.out_off[b];
        const uint64_t cap = a.out_cap ? (uint64_t)a.out_cap[b] : len;

        int status = LZF_OK;
        uint64_t opos = 0;

        // assert!(input.len() <= T::payload_size_limit())  :167 ; Slot width must hold every position
        const bool too_big = (kHash4 && len > 0xffffull) || (sizeof(Slot) == 2 && len > 0x10000ull) ||
                             (a.max_block_len && len > a.max_block_len);
        if (too_big) {
            status = LZF_PANIC;
        } else {
            // fresh zeroed table (U32Table::default :32-36 / template_table.clone() compress.rs:270)
            {
                uint4* t4 = reinterpret_cast<uint4*>(table);
                const uint32_t nvec = nslots * (uint32_t)sizeof(Slot) / 16;
                for (uint32_t i = lane; i < nvec; i += 32) t4[i] = make_uint4(0, 0, 0, 0);
            }
            __syncwarp();

            uint64_t cursor = 0;
            while (cursor < len) {                                            // :171
                const uint64_t literal_start = cursor;
                uint32_t j0 = 0;
                bool finished = false;
                uint64_t cur = 0, cnd = 0;
                for (;;) {                                                    // :177, 32 probes per trip
                    const uint64_t p = literal_start + probe_offset(j0 + lane);
                    const bool is_end = (p >= len) || (len - p < 12);         // :178
                    uint64_t v = 0;
                    uint32_t h = 0xffff0000u | lane;                          // unique key for idle lanes
                    if (!is_end) {
                        v = ld_u64_unaligned(in + p, in_end);
                        h = kHash4 ? hash4((uint32_t)v, hashlog) : hash5(v, hashlog);
                    }
                    const uint32_t same = __match_any_sync(LZF_FULL_MASK, h);
                    const uint32_t lower = same & ((1u << lane) - 1u);
                    const uint32_t src_lane = lower ? (31u - __clz(lower)) : lane;
                    const uint32_t p_lo = (uint32_t)p;
                    const uint32_t in_batch = __shfl_sync(LZF_FULL_MASK, p_lo, src_lane);
                    uint64_t cand = 0;
                    bool ok = false;
                    if (!is_end) {
                        cand = lower ? (uint64_t)in_batch : (uint64_t)table[h];   // table.replace :196 (read half)
                        ok = (p != 0) && (p - cand <= 0xffffull) &&               // :200-201
                             ld_u32_unaligned(in + cand) == (uint32_t)v;           // >= MINMATCH bytes :206
                    }
                    const uint32_t trig = __ballot_sync(LZF_FULL_MASK, ok || is_end);
                    if (trig == 0) {
                        if ((same >> lane) == 1u) table[h] = (Slot)p;         // last writer of each slot wins
                        __syncwarp();
                        j0 += 32;
                        continue;
                    }
                    const uint32_t w = __ffs(trig) - 1;
                    const bool w_is_end = __shfl_sync(LZF_FULL_MASK, (int)is_end, w) != 0;
                    // probes before the winner (and the winner itself when it is a match) did replace()
                    const uint32_t commit = w_is_end ? ((1u << w) - 1u) : ((2u << w) - 1u);
                    if (((commit >> lane) & 1u) && ((same & commit) >> lane) == 1u) table[h] = (Slot)p;
                    __syncwarp();
                    finished = w_is_end;
                    cur = literal_start + probe_offset(j0 + w);
                    cnd = __shfl_sync(LZF_FULL_MASK, (uint32_t)cand, w);
                    break;
                }

                if (finished) {
                    // final literal-only sequence  :178-190
                    const uint64_t L = len - literal_start;
                    const uint64_t total = 1 + lsic_len(L) + L;
                    if (opos + total > cap) { status = LZF_WRITER_FULL; break; }
                    if (lane == 0) out[opos] = (uint8_t)((L < 15 ? L : 15) << 4);
                    write_lsic(out + opos + 1, L);
                    warp_copy(out + opos + 1 + lsic_len(L), in + literal_start, L);
                    opos += total;
                    cursor = len;
                    break;
                }

                // ---- forward extension: count_matching_bytes(input[cur..len-5], input[cnd..])  :117-145,203-204
                const uint64_t limit = len - 5 - cur;
                uint64_t matching = 4;
                for (;;) {
                    const uint64_t idx = matching + (uint64_t)lane * 8;
                    uint32_t cnt = 0;
                    if (idx < limit) {
                        const uint64_t x = ld_u64_unaligned(in + cur + idx, in_end) ^ ld_u64_unaligned(in + cnd + idx, in_end);
                        cnt = x ? (uint32_t)(__ffsll((long long)x) - 1) >> 3 : 8u;
                        const uint64_t room = limit - idx;
                        if (cnt > room) cnt = (uint32_t)room;
                    }
                    const uint32_t stop = __ballot_sync(LZF_FULL_MASK, cnt != 8u);
                    if (stop == 0) { matching += 256; continue; }
                    const uint32_t f = __ffs(stop) - 1;
                    matching += (uint64_t)f * 8 + __shfl_sync(LZF_FULL_MASK, cnt, f);
                    break;
                }
                // ---- backtrack  :211-214
                uint64_t backtrack = 0;
                {
                    const uint64_t max_backtrack = min(cur - literal_start, cnd);
                    while (backtrack < max_backtrack) {
                        const uint64_t k = backtrack + lane;
                        const bool eq = k < max_backtrack && in[cur - 1 - k] == in[cnd - 1 - k];
                        const uint32_t stop = __ballot_sync(LZF_FULL_MASK, !eq);
                        if (stop == 0) { backtrack += 32; continue; }
                        backtrack += __ffs(stop) - 1;
                        break;
                    }
                }
                const uint64_t extra = matching - 4 + backtrack;              // :206,214
                const uint32_t offset = (uint32_t)(cur - cnd);                // :208
                cursor = cur + matching;                                      // :215
                // table.replace(input, cursor - 2)  :218
                if (lane == 0) {
                    const uint64_t q = cursor - 2;
                    uint32_t h2;
                    if (kHash4) h2 = hash4(ld_u32_unaligned(in + q), hashlog);
                    else h2 = hash5(len - q >= 8 ? ld_u64_unaligned(in + q, in_end) : 0ull, hashlog);   // :43 unwrap_or(0)
                    table[h2] = (Slot)q;
                }
                __syncwarp();

                // ---- write_group  :150-163,235-236
                const uint64_t L = cur - backtrack - literal_start;
                const uint64_t ll = lsic_len(L), ml = lsic_len(extra);
                const uint64_t total = 1 + ll + L + 2 + ml;
                if (opos + total > cap) { status = LZF_WRITER_FULL; break; }
                uint8_t* o = out + opos;
                if (lane == 0) o[0] = (uint8_t)(((L < 15 ? L : 15) << 4) | (extra < 15 ? extra : 15));
                write_lsic(o + 1, L);
                warp_copy(o + 1 + ll, in + literal_start, L);
                if (lane < 2) o[1 + ll + L + lane] = (uint8_t)(offset >> (8 * lane));
                write_lsic(o + 1 + ll + L + 2, extra);
                opos += total;
            }
        }

        __syncwarp();
        if (a.xxh_plain || a.xxh_stored) {
            const uint32_t hp = warp_xxh32(in, len);
            if (lane == 0 && a.xxh_plain) a.xxh_plain[b] = hp;
            if (a.xxh_stored) {
                const uint32_t hs = (status == LZF_OK) ? warp_xxh32(out, opos) : hp;
                if (lane == 0) a.xxh_stored[b] = hs;
            }
        }
        if (lane == 0) {
            a.out_len[b] = (status == LZF_OK) ? (uint32_t)opos : 0u;
            a.status[b] = status;
        }
    }
}

}  // namespace lzf

// Host-side launcher.  Returns a cudaError_t as int.
extern "C" int lzf_launch_encode(const lzf::EncodeArgs* args, int num_sms, cudaStream_t stream) {
    using namespace lzf;
    if (args->nblocks == 0) return 0;
    const uint32_t hashlog = args->hashlog;
    const bool hash4 = args->table_kind == LZF_TABLE_U16;
    const uint32_t nslots = hash4 ? (2u << hashlog) : (1u << hashlog);
    // u16 slots are exact whenever every position fits 16 bits
    const bool slot16 = hash4 || (args->max_block_len != 0 && args->max_block_len <= 65536u);
    const size_t table_bytes = (size_t)nslots * (slot16 ? 2 : 4);
    // per-warp tables up to 32 KiB live in shared memory; larger ones (hashlog >= 14 extension) in a
    // ctx-owned global scratch that stays L2-resident
    const bool smem_tables = table_bytes <= 32 * 1024;
    if (!smem_tables && args->global_tables == nullptr) return (int)cudaErrorInvalidValue;
    void (*kern)(EncodeArgs, uint32_t, int);
    if (hash4) kern = encode_blocks_kernel<uint16_t, true>;
    else if (slot16) kern = encode_blocks_kernel<uint16_t, false>;
    else kern = encode_blocks_kernel<uint32_t, false>;
    cudaError_t e;
    int ctas_per_sm = 1;
    const size_t dyn = smem_tables ? table_bytes * kEncodeWarpsPerCta : 0;
    if (dyn > 48 * 1024) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return (int)e;
    }
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, kEncodeWarpsPerCta * 32, dyn);
    if (e != cudaSuccess) return (int)e;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (!smem_tables && ctas_per_sm > kGlobalTableCtasPerSm) ctas_per_sm = kGlobalTableCtasPerSm;
    unsigned grid = (unsigned)(num_sms * ctas_per_sm);
    const unsigned need = (args->nblocks + kEncodeWarpsPerCta - 1) / kEncodeWarpsPerCta;
    if (grid > need) grid = need;
    LZF_LAUNCH(kern, grid, kEncodeWarpsPerCta * 32, dyn, stream, *args, nslots, smem_tables ? 1 : 0);
    return (int)cudaGetLastError();
}

// Number of warps a global-table launch may start (sizing of the scratch).
extern "C" size_t lzf_encode_global_table_warps(int num_sms) {
    return (size_t)num_sms * lzf::kGlobalTableCtasPerSm * lzf::kEncodeWarpsPerCta;
}
