// lzf_compress.cu — batched LZ4 block compress for sm_100a, one warp per independent block.
//
// Behavioural contract: raw::compress2 of the reference (src/raw/compress/mod.rs:165-238) with a
// fresh zeroed EncoderTable and cursor 0, writing through NoPartialWrites(out[..cap])
// (src/framed/compress.rs:242,294-308).  The emitted bytes are IDENTICAL to the reference's.
//
// The reference parse is one serial chain per block (hash -> table swap -> candidate compare ->
// extend -> emit).  We keep its exact semantics but evaluate it 32 positions at a time:
//
//   * a BATCH is 32 probe positions of the current literal run, one per lane.  While the run is
//     young (fewer than 67 probes, src/raw/compress/mod.rs:174-175,225-231: step == 1) they are 32
//     CONSECUTIVE bytes; afterwards they follow the closed form of the step recurrence;
//   * every lane hashes its position, reads its table slot and verifies its table candidate with
//     one 4-byte load — one L2 round trip for the whole batch instead of one per probe;
//   * the serial table semantics (mem::swap at :68) are restored with __match_any_sync: a lane
//     whose slot was overwritten by an earlier *inserted* lane of the same batch takes that lane's
//     position as its candidate;
//   * a ballot picks the first lane whose candidate is a real >= 4-byte match (or that hits the
//     end-of-block rule :178).  In a consecutive batch the parse then CONTINUES inside the same batch:
//     the match end lands on some later lane, the lanes in between were never probed (only
//     `cursor - 2` is inserted, :218) and the lanes after it are re-evaluated against the updated
//     set of inserted lanes — several sequences per batch without re-hashing anything;
//   * forward and backward match extension (:117-145, :211-214) share one round trip: 24 lanes
//     compare 96 bytes ahead, 8 lanes 32 bytes behind;
//   * table writes are committed once per batch (last inserted lane per slot wins);
//   * a whole short sequence (token, LSIC, literals, offset, LSIC) is written with one byte per
//     lane; long literal runs use 16-byte vector copies.
// The per-warp hash table lives in shared memory (16 KiB for the reference's 4096 x u32; 8 KiB
// when every position fits u16).
#include "lzf_kernels.cuh"

#include <stdlib.h>
#include <type_traits>

// encoder path counters of the CPU test harness (tests/simt): which way the batches and sequences went
#if defined(LZF_SIMT_EMU) && defined(LZF_ENC_STATS)
extern "C" { uint64_t lzf_enc_stats[32]; }
#define LZF_STAT(i) do { if (lane_id() == 0) lzf_enc_stats[i]++; } while (0)
#else
#define LZF_STAT(i) do { } while (0)
#endif
#ifndef LZF_ENC_WALK
#define LZF_ENC_WALK 1          // 0: the round-1 resolve loop only (A/B builds)
#endif


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
constexpr uint32_t kConsecutiveProbes = 67;     // probe_offset(j) == j for j < 67

// hash_for_u32, 64-bit little-endian branch (:40-51): ((v << 24) * 889523592379) >> (64 - hashlog),
// evaluated from the five low bytes of v in 32-bit arithmetic.
__device__ __forceinline__ uint32_t hash5(uint32_t v32, uint32_t b4, uint32_t hashlog) {
    const uint32_t x_lo = v32 << 24;
    const uint32_t x_hi = (v32 >> 8) | (b4 << 24);
    const uint32_t k_lo = 0x1BBCDCBBu, k_hi = 0xCFu;           // 889523592379 = 0xCF1BBCDCBB
    const uint32_t hi = __umulhi(x_lo, k_lo) + x_lo * k_hi + x_hi * k_lo;
    return hi >> (32 - hashlog);
}
// hash_for_u16 (:58-61): one more bit than hashlog because the u16 table has twice the slots
__device__ __forceinline__ uint32_t hash4(uint32_t v, uint32_t hashlog) {
    return (v * 2654435761u) >> (32 - hashlog - 1);
}

// bytes write_lsic_tail (:243-260) emits for `value`
__device__ __forceinline__ uint32_t lsic_len(uint32_t value) {
    return value < 15 ? 0 : (value - 15) / 255 + 1;
}
// warp-parallel write_lsic_tail
__device__ __forceinline__ void write_lsic(uint8_t* dst, uint32_t value) {
    if (value < 15) return;
    const uint32_t nbytes = (value - 15) / 255 + 1;
    const uint8_t last = (uint8_t)((value - 15) % 255);
    for (uint32_t i = lane_id(); i < nbytes; i += 32) dst[i] = (i == nbytes - 1) ? last : 0xff;
}

// 4 bytes at in[pos] (any alignment), pos + 4 <= block length: the aligned words read all hold at
// least one byte of the range, so nothing outside the allocation is touched
__device__ __forceinline__ uint32_t ld4(const uint8_t* in, uint32_t pos) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(in + pos);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    const unsigned sh = (unsigned)(a & 3u) * 8u;
    const uint32_t lo = __ldg(w);
    if (sh == 0) return lo;
    return __funnelshift_r(lo, __ldg(w + 1), sh);
}

// Host-buffer pipeline: blocks until bytes [0, need) of the block have arrived (EncodeArgs::progress); returns
// the arrived byte count (saturated).  Warp-uniform.
__device__ __forceinline__ uint32_t wait_arrival(const uint32_t* progress, uint32_t slice_bytes, uint32_t need) {
    for (;;) {
        const uint64_t bytes = (uint64_t)ld_acquire_sys(progress) * slice_bytes;
        if (bytes >= need) return bytes > 0xffffffffull ? 0xffffffffu : (uint32_t)bytes;
        spin_pause();
    }
}

// One LZ4 sequence (write_group :150-163, or the literal-only tail :182-189 when `final`).
// Returns false when the bounded writer would refuse it (NoPartialWrites, compress.rs:298-301).
__device__ __forceinline__ bool emit_sequence(uint8_t* out, uint32_t& opos, uint32_t cap, const uint8_t* in,
                                              uint32_t lit_start, uint32_t L, uint32_t offset, uint32_t extra, bool final,
                                              bool lits_in_regs, uint32_t v32, uint32_t first_lit_lane) {
    const unsigned lane = lane_id();
    const uint32_t ll = lsic_len(L), ml = final ? 0u : lsic_len(extra);
    const uint64_t total = 1ull + ll + L + (final ? 0u : 2u + ml);
    if ((uint64_t)opos + total > cap) return false;
    uint8_t* o = out + opos;
    const uint8_t token = (uint8_t)(((L < 15 ? L : 15) << 4) | (final ? 0u : (extra < 15 ? extra : 15)));
    if (total <= 32) {
        // one byte per lane: token | lsic(L) | literals | offset | lsic(extra)
        const uint32_t t = lane - 1 - ll;                         // literal index of this lane (when in range)
        uint32_t litb = 0;
        if (lits_in_regs) litb = __shfl_sync(LZF_FULL_MASK, v32, first_lit_lane + t) & 0xffu;   // lane of in[lit_start + t]
        if (lane < total) {
            uint8_t v;
            if (lane == 0) v = token;
            else if (lane < 1 + ll) v = (lane == ll) ? (uint8_t)((L - 15) % 255) : 0xff;
            else if (lane < 1 + ll + L) v = lits_in_regs ? (uint8_t)litb : __ldg(in + lit_start + t);
            else if (lane < 1 + ll + L + 2) v = (uint8_t)(offset >> (8 * (lane - 1 - ll - L)));
            else v = (lane == total - 1) ? (uint8_t)((extra - 15) % 255) : 0xff;
            o[lane] = v;
        }
    } else {
        if (lane == 0) o[0] = token;
        write_lsic(o + 1, L);
        warp_copy(o + 1 + ll, in + lit_start, L);
        if (!final) {
            if (lane < 2) o[1 + ll + L + lane] = (uint8_t)(offset >> (8 * lane));
            write_lsic(o + 1 + ll + L + 2, extra);
        }
    }
    opos += (uint32_t)total;
    return true;
}

// ---- the per-warp hash table ----------------------------------------------------------------------
// Slot = uint32_t / uint16_t: the position itself (u16 only when every position of the block fits).
// Packed17: blocks up to 16 MiB keep 17 bits per slot — a u16 array plus a bit array, 8.5 KiB instead
// of 16 KiB, so twice as many blocks are resident per SM.  A candidate only matters while it is at
// most 65 535 bytes behind the probe (:200-201), so positions are kept modulo 2^17 and every 64 KiB
// of progress a sweep re-encodes the slots that have fallen out of the window as "65 536 behind the
// sweep point" — they stay out of reach until the next sweep, exactly like the stale u32 positions
// of the reference.  Distances 1..65 535 are therefore always exact and anything else is rejected.
template <typename Slot>
struct PlainTable {
    Slot* t;
    __device__ __forceinline__ uint32_t dist(uint32_t h, uint32_t p) const { return p - (uint32_t)t[h]; }
    __device__ __forceinline__ void put(uint32_t h, uint32_t p) { t[h] = (Slot)p; }
};
struct Packed17Table {
    uint16_t* lo;
    uint32_t* hi;          // one bit per slot
    __device__ __forceinline__ uint32_t get(uint32_t h) const { return (uint32_t)lo[h] | (((hi[h >> 5] >> (h & 31u)) & 1u) << 16); }
    __device__ __forceinline__ uint32_t dist(uint32_t h, uint32_t p) const { return (p - get(h)) & 0x1ffffu; }
    __device__ __forceinline__ void put(uint32_t h, uint32_t p) {
        lo[h] = (uint16_t)p;
        if ((p >> 16) & 1u) atomicOr(&hi[h >> 5], 1u << (h & 31u));
        else atomicAnd(&hi[h >> 5], ~(1u << (h & 31u)));
    }
    // all 32 lanes: slots whose position is more than 65 535 behind `at` become "65 536 behind `at`"
    __device__ __forceinline__ void sweep(uint32_t nslots, uint32_t at) {
        const unsigned lane = lane_id();
        const uint32_t dead = (at - 65536u) & 0x1ffffu;
        for (uint32_t i = lane; i < nslots; i += 32) {
            uint32_t e = get(i);
            const uint32_t age = (at - e) & 0x1ffffu;
            if (age > 0xffffu || age == 0) { e = dead; lo[i] = (uint16_t)e; }
            const uint32_t word = __ballot_sync(LZF_FULL_MASK, (e >> 16) & 1u);
            if (lane == 0) hi[i >> 5] = word;
        }
        __syncwarp();
    }
};
constexpr uint32_t kPacked17MaxLen = 16u << 20;

#ifndef LZF_ENC_WARPS
#define LZF_ENC_WARPS 4
#endif
#ifndef LZF_ENC_BIG_WARPS
#define LZF_ENC_BIG_WARPS 28
#endif
constexpr int kEncodeWarpsPerCta = LZF_ENC_WARPS;
// The parse is one long dependent chain per warp (~8 cycles per instruction, ncu r01_ncu_encode_v5), so
// throughput is resident warps x chain speed.  Tables of <= 8.5 KiB (u16 / packed 17-bit slots) therefore
// run ONE CTA of kEncodeBigWarps warps per SM at <= 72 registers (4096 blocks of config 3 = one wave on 148
// SMs).  Half of the warps keep their table in shared memory, the other half in an L2-resident global
// scratch read through L1: measured on B200 (config 3, GiB/s) 13 shared + 15 global 32.2, all global 31.6,
// 26 shared + 2 global 29.6 (a 227 KiB carve-out leaves almost no L1 for the input and candidate loads),
// 32 warps 30.2, 16 warps (two waves) 22.8.
constexpr int kEncodeBigWarps = LZF_ENC_BIG_WARPS;
constexpr int kEncodeSmemWarpsMax = kEncodeBigWarps / 2 - 1;
constexpr int kGlobalTableCtasPerSm = 4;   // bounds the global table scratch (CTAs of kEncodeWarpsPerCta warps)
constexpr size_t kSmemPerCtaMax = 227 * 1024;   // sm_100a opt-in maximum per CTA

// kTab: 0 = u32 slots, 1 = u16 slots (positions fit 16 bits), 2 = packed 17-bit slots
template <int kTab, bool kHash4, int kWarps>
__global__ void __launch_bounds__(kWarps * 32)
encode_blocks_kernel(EncodeArgs a, uint32_t nslots, int n_smem_warps) {
    LZF_DYN_SMEM(smem_raw);
    __shared__ HashQueue hashq;
    constexpr bool kPacked = kTab == 2;
    using Slot = typename std::conditional<kTab == 0, uint32_t, uint16_t>::type;
    constexpr int kEncodeWarpsPerCta = kWarps;
    const unsigned lane = lane_id();
    const unsigned warp_in_cta = threadIdx.x >> 5;
    const size_t table_bytes = kPacked ? (size_t)nslots * 2 + nslots / 8 : (size_t)nslots * sizeof(Slot);
    uint8_t* table_mem;
    if ((int)warp_in_cta < n_smem_warps) {
        table_mem = smem_raw + (size_t)warp_in_cta * table_bytes;
    } else {
        const size_t gw = (size_t)blockIdx.x * (kEncodeWarpsPerCta - n_smem_warps) + (warp_in_cta - n_smem_warps);
        table_mem = a.global_tables + gw * table_bytes;
    }
    typename std::conditional<kPacked, Packed17Table, PlainTable<Slot>>::type table;
    if constexpr (kPacked) {
        table.lo = reinterpret_cast<uint16_t*>(table_mem);
        table.hi = reinterpret_cast<uint32_t*>(table_mem + (size_t)nslots * 2);
    } else {
        table.t = reinterpret_cast<Slot*>(table_mem);
    }
    const uint32_t hashlog = a.hashlog;
    const uint32_t lower_mask = (1u << lane) - 1u;
    hash_queue_init(&hashq);

    const uint32_t nchains = a.chain_first ? a.nchains : a.nblocks;
    for (;;) {
        uint32_t chain = 0;
        if (lane == 0) chain = atomicAdd(a.work_counter, 1u);
        chain = __shfl_sync(LZF_FULL_MASK, chain, 0);
        if (chain >= nchains) break;
        const uint32_t b_first = a.chain_first ? a.chain_first[chain] : chain;
        const uint32_t b_count = a.chain_first ? a.chain_count[chain] : 1u;
        uint32_t last_sweep = 0;        // packed tables: every slot is exact for stream positions below last_sweep + 65536
        bool carried_in = false;        // the table came from EncodeArgs::table_io and goes back there

      for (uint32_t bi = 0; bi < b_count; bi++) {
        const uint32_t b = b_first + bi;
        const uint32_t own_len = a.in_len[b];
        const uint32_t cursor0 = a.prefix_len ? a.prefix_len[b] : 0u;         // compress2's `cursor` argument
        const uint32_t ab = a.abs_base ? a.abs_base[b] : 0u;                  // table.offset (:30,65,72-74)
        const uint64_t len64 = (uint64_t)cursor0 + own_len;
        const uint32_t len = (uint32_t)len64;                                 // input.len(): history + block
        const uint8_t* in = a.in + a.in_off[b] - cursor0;
        uint8_t* out = a.out + a.out_off[b];
        const uint32_t cap = a.out_cap ? a.out_cap[b] : own_len;  // NoPartialWrites bound (compress.rs:242)

        int status = LZF_OK;
        uint32_t opos = 0;
        // plaintext still in flight over PCIe (independent blocks without history only): bytes below `arrived` are there
        const bool gated = a.progress != nullptr && cursor0 == 0;
        uint32_t arrived = gated ? 0u : 0xffffffffu;

        // assert!(input.len() <= T::payload_size_limit())  :167 / "EncoderTable contract violated" :67,92: the slot
        // width must hold every stream position that is INSERTED — probes stop 12 bytes and cursor - 2 lies 7 bytes in
        // front of the end (:178,218), so a block may reach 7 positions past the limit
        const uint64_t reach = len64 + ab - (len64 + ab > 7 ? 7 : 0);
        const bool too_big = len64 > 0xffffffffull || (kHash4 && len64 > 0xffffull) ||
                             (!a.allow_slot_wrap && (reach > 0xffffffffull || (kHash4 && reach > 0xffffull))) ||
                             (kTab == 1 && !kHash4 && len64 + ab > 0x10000ull) || (kPacked && own_len > kPacked17MaxLen) ||
                             (a.max_block_len && own_len > a.max_block_len);
        if (too_big) {
            status = LZF_PANIC;
        } else if (own_len) {
            if (bi == 0) {
                // fresh zeroed table (U32Table::default :32-36 / template_table.clone() compress.rs:270)
                if (!kPacked && a.table_io && chain == 0) {
                    // a table carried over from an earlier call (compress2's `table: &mut T`, :165-170)
                    if constexpr (!kPacked) for (uint32_t i = lane; i < nslots; i += 32) table.t[i] = (Slot)a.table_io[i];
                    carried_in = true;
                } else {
                    uint4* t4 = reinterpret_cast<uint4*>(table_mem);
                    const uint32_t nvec = (uint32_t)(table_bytes / 16);
                    for (uint32_t i = lane; i < nvec; i += 32) t4[i] = make_uint4(0, 0, 0, 0);
                }
                __syncwarp();
                last_sweep = 0;
                // dictionary priming: template_table.replace(dict, off) for off = 0, 3, 6, ... while 8 bytes
                // remain (compress.rs:204-214), 32 positions per step, the last insert of a slot wins
                const uint32_t prime = a.prime_len ? a.prime_len[b] : 0u;
                for (uint32_t o0 = 0; o0 + 8 <= prime; o0 += 96) {
                    const uint32_t o = o0 + 3 * lane;
                    const bool on = o + 8 <= prime;
                    uint32_t hk = 0xffff0000u | lane;
                    if (on) hk = kHash4 ? hash4(ld4(in, o), hashlog) : hash5(ld4(in, o), in[o + 4], hashlog);
                    const uint32_t sm_ = __match_any_sync(LZF_FULL_MASK, hk);
                    if constexpr (kPacked) {
                        if (o0 + 96 + ab - last_sweep >= 65536u) { last_sweep = o0 + ab; table.sweep(nslots, last_sweep); }
                    }
                    if (on && (sm_ >> lane) == 1u) table.put(hk, o + ab);
                    __syncwarp();
                }
            }
            uint32_t lit_start = cursor0;   // start of the current literal run
            uint32_t j = 0;             // probes already done in the current run
            bool done = false;
            while (!done) {                                                   // :171 / :177, 32 probes per trip
                // ---- the batch: one probe position per lane
                const bool consecutive = j + 32 <= kConsecutiveProbes;
                const uint64_t p64 = (uint64_t)lit_start + (consecutive ? (uint64_t)(j + lane) : probe_offset(j + lane));
                const bool is_end = p64 >= len || len - (uint32_t)p64 < 12;    // :178
                const uint32_t p = (uint32_t)p64;
                const uint32_t endmask = __ballot_sync(LZF_FULL_MASK, is_end);
                // first and last probe position of the batch (the last one clipped to the block): warp-uniform arithmetic
                const uint64_t first64 = (uint64_t)lit_start + (consecutive ? (uint64_t)j : probe_offset(j));
                const uint64_t last64 = (uint64_t)lit_start + (consecutive ? (uint64_t)(j + 31) : probe_offset(j + 31));
                const uint32_t base = (uint32_t)first64;
                const uint32_t p_top = last64 < len ? (uint32_t)last64 : len;
                if (gated) {
                    // the batch reads at most 16 bytes past its last probe position
                    const uint32_t need = len - p_top < 64 ? len : p_top + 64;
                    if (need > arrived) arrived = wait_arrival(a.progress, a.slice_bytes, need);
                }
                if constexpr (kPacked) {
                    // Every read of this batch must lie below last_sweep + 65536.  A sweep at `at` is exact only while
                    // at <= last_sweep + 65536 (older slots would alias modulo 2^17) and no inserted position lies at or
                    // beyond it — every insert so far is below last_sweep + 65536 (see the catch-up before the cursor - 2
                    // insert) — and it must not lie beyond the next read, so a jump of the parse (the skip step, a refused
                    // block in front of this one in a chain) is walked in steps of 65536.
                    while (p_top + ab - last_sweep >= 65536u) {
                        const bool reach = base + ab - last_sweep <= 65536u;
                        last_sweep = reach ? base + ab : last_sweep + 65536u;
                        table.sweep(nslots, last_sweep);
                        if (reach) break;
                    }
                }
                uint32_t v32 = 0, h = 0xffff0000u | lane;                     // unique key for idle lanes
                uint32_t tcand = 0, tdist = 0;
                uint32_t n_a1 = 0, n_a2 = 0, n_a3 = 0, n_am1 = 0;             // this lane's bytes p+4.., p+8.., p+12.., p-4..
                if (!is_end) {
                    const uintptr_t ad = reinterpret_cast<uintptr_t>(in + p);
                    const uint32_t* w = reinterpret_cast<const uint32_t*>(ad & ~uintptr_t(3));
                    const unsigned sh = (unsigned)(ad & 3u) * 8u;
                    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1);          // 12 bytes remain: both words are inside the block
                    v32 = __funnelshift_r(w0, w1, sh);
                    if (consecutive) {
                        // bytes p+4.., p+8.., p+12.. and p-4.. for the extension summary below: same cache lines as
                        // w0 / w1, fetched in the same round trip
                        const uint32_t w2 = __ldg(w + 2);                     // bytes .. p+11
                        n_a1 = __funnelshift_r(w1, w2, sh);
                        if (p + 16 <= len) {
                            const uint32_t w3 = __ldg(w + 3);
                            const uint32_t w4 = sh ? __ldg(w + 4) : 0u;       // holds byte p+15 when p is unaligned
                            n_a2 = __funnelshift_r(w2, w3, sh);
                            n_a3 = __funnelshift_r(w3, w4, sh);
                        }
                        if (p >= 4) n_am1 = __funnelshift_r(__ldg(w - 1), w0, sh);
                    }
                    h = kHash4 ? hash4(v32, hashlog) : hash5(v32, (w1 >> sh) & 0xffu, hashlog);
                    tdist = table.dist(h, p + ab);                            // table.replace :196 (read half)
                    if (ab && tdist > p) tdist = p;                           // old.saturating_sub(self.offset) :70 -> candidate 0
                    tcand = p - tdist;
                }
                const uint32_t same = __match_any_sync(LZF_FULL_MASK, h);
                // table candidate: addressable (:200-201) and >= MINMATCH equal bytes (:206)
                bool t_ok = false;
                // FAST-PATH SUMMARY (consecutive batches): each lane also fetches 16 bytes after and 4 bytes
                // before its table candidate in the same round trip, and compares them with the bytes around its
                // own position (n_a1 .. n_am1, loaded with the probe bytes above).  If the lane later
                // wins, its forward match length (up to 16) and backtrack (up to 4) are already known and the
                // whole sequence is resolved with shuffles only — no further memory round trip.
                //   fsum: bits 0..4 forward match length 4..16, bits 5..7 backtrack 0..4,
                //         bit 8 forward summary valid, bit 9 backward summary valid
                uint32_t fsum = 0;
                // addressable (:200-201): not the first position of the block (cursor != init_cursor), 1 <= distance
                // <= 0xFFFF, and inside the history that is physically there
                if (!is_end && p != cursor0 && tdist - 1u < 0xffffu && tdist <= p) {
                    if (consecutive && tcand >= 4 && tcand + 16 <= len) {
                        // 4 bytes before and 16 bytes after the candidate, one round trip
                        const uintptr_t ad = reinterpret_cast<uintptr_t>(in + tcand - 4);
                        const uint32_t* w = reinterpret_cast<const uint32_t*>(ad & ~uintptr_t(3));
                        const unsigned sh = (unsigned)(ad & 3u) * 8u;
                        uint32_t W[6];
#pragma unroll
                        for (int i = 0; i < 5; i++) W[i] = __ldg(w + i);
                        W[5] = sh ? __ldg(w + 5) : 0u;
                        uint32_t U[5];
#pragma unroll
                        for (int i = 0; i < 5; i++) U[i] = __funnelshift_r(W[i], W[i + 1], sh);
                        t_ok = U[1] == v32;
                        if (t_ok) {
                            if (p + 16 <= len) {                                 // bytes p .. p+15 are inside the block
                                const uint32_t x1 = n_a1 ^ U[2], x2 = n_a2 ^ U[3], x3 = n_a3 ^ U[4];
                                uint32_t fm;
                                if (x1) fm = 4 + ((uint32_t)(__ffs((int)x1) - 1) >> 3);
                                else if (x2) fm = 8 + ((uint32_t)(__ffs((int)x2) - 1) >> 3);
                                else if (x3) fm = 12 + ((uint32_t)(__ffs((int)x3) - 1) >> 3);
                                else fm = 16;
                                fsum |= fm | 0x100u;
                            }
                            if (p >= 4) {
                                const uint32_t xb = n_am1 ^ U[0];
                                const uint32_t nb = xb ? (uint32_t)__clz((int)xb) >> 3 : 4u;
                                fsum |= (nb << 5) | 0x200u;
                            }
                        }
                    } else {
                        t_ok = ld4(in, tcand) == v32;
                    }
                }
                uint32_t ins = 0;       // lanes whose probe (or cursor-2 insert) has happened, in order
                uint32_t s = 0;         // first lane of the current run inside this batch
                uint32_t left_q2 = 0xffffffffu;   // the match left the batch: its cursor-2 insert, applied behind the batch's own
                uint32_t late_q2 = 0xffffffffu;   // a cursor-2 insert that falls on a lane without a hash (last 11 bytes)
                // Batches in which no two lanes share a table slot (the common case) need no per-sequence
                // candidate resolution at all: who matches is one ballot per batch, and each further sequence
                // of the batch is found with scalar bit arithmetic.
                const bool no_dups = __ballot_sync(LZF_FULL_MASK, (same & ~(1u << lane)) != 0) == 0;
                const uint32_t okmask = __ballot_sync(LZF_FULL_MASK, t_ok);
                // winner's candidate distance and extension summary travel in one word
                const uint32_t packed_w = tdist | (fsum << 16);
#if LZF_ENC_WALK
                // TIGHT WALK (consecutive batches).  A lane is GOOD when its table candidate matches, no earlier lane of
                // the batch shares its table slot (so the candidate cannot change with the parse) and its extension
                // summary is complete and short: everything such a lane would emit is then known from its own
                // registers.  The walk goes from run start to winner to match end with scalar (warp-uniform)
                // arithmetic, one shuffle to fetch the winner's summary, one to gather the literal bytes and one
                // store per sequence.  It stops in front of anything else (an end-of-block lane, a lane whose slot
                // an earlier lane shares, length extensions, literals carried in from earlier batches — letting the walk take
                // those too measured 34.9 against 35.4 GiB/s): the serial
                // step below resolves that one sequence and the walk resumes behind it.
                //   pinfo: bits 0..5 lane where the match ends (4..47), 6..7 backtrack summary (0..3), 16..31 candidate distance
                uint32_t pinfo = 0;
                {
                    const bool clean = (same & lower_mask) == 0;
                    if (consecutive && t_ok && clean && (fsum & 0x300u) == 0x300u) {
                        const uint32_t fm = fsum & 31u, limit = len - 5 - p, nb = (fsum >> 5) & 7u;
                        // length extensions (token nibble 15) and backtracks beyond the summary take the serial step
                        if ((fm < 16 || limit <= 16) && nb < 4 && min(fm, limit) + nb < 19) pinfo = (lane + min(fm, limit)) | (nb << 6) | (tdist << 16);
                    }
                }
                const uint32_t goodmask = __ballot_sync(LZF_FULL_MASK, pinfo != 0);
                // lanes the walk cannot pass without looking: matches, the end of the block, shared slots
                const uint32_t trigmask = okmask | endmask | (no_dups ? 0u : __ballot_sync(LZF_FULL_MASK, (same & lower_mask) != 0));
#endif
                for (;;) {
#if LZF_ENC_WALK
                    if (consecutive && lit_start == base + s && s < 32 && cap >= opos && cap - opos >= 128) {
                        int end = 0;            // 0: stopped in front of a lane (serial step), 1: no trigger left, 2: match left the batch
                        bool wrote = false;     // at least one sequence was written: the run now starts where the last match ended
                        uint32_t e = s;
                        for (;;) {
                            const uint32_t t = trigmask & ~((1u << s) - 1u);
                            if (t == 0) { end = 1; break; }
                            const uint32_t w = (uint32_t)__ffs((int)t) - 1u;
                            if (!((goodmask >> w) & 1u)) break;
                            const uint32_t info = __shfl_sync(LZF_FULL_MASK, pinfo, w);
                            const uint32_t lraw = w - s;
                            const uint32_t bt = min((info >> 6) & 3u, lraw);     // :211-214 (the candidate has >= 4 bytes in front of it)
                            const uint32_t L = lraw - bt;
                            if (L >= 15) break;                                   // length bytes: serial step
                            e = info & 63u;
                            // write_group :150-163, one byte per lane: token | literals | offset
                            const uint32_t tl = lane - 1u;
                            const uint32_t litb = __shfl_sync(LZF_FULL_MASK, v32, s + tl);
                            const uint32_t v = lane == 0 ? ((L << 4) | (e - w - 4u + bt)) : (tl < L ? litb : (info >> 16) >> (8u * (tl - L)));
                            if (lane < L + 3u) out[opos + lane] = (uint8_t)v;
                            opos += L + 3u;
                            ins |= ((2u << w) - 1u) & ~((1u << s) - 1u);         // probes s..w happened
                            wrote = true;
                            LZF_STAT(4);
                            if (e >= 32) { end = 2; break; }
                            const uint32_t l2 = e - 2;                            // table.replace(cursor - 2) :218
                            if (!((endmask >> l2) & 1u)) ins |= 1u << l2;
                            else late_q2 = base + l2;
                            s = e;
                        }
                        if (wrote) { lit_start = base + e; j = 0; }
                        if (end == 1) {
                            ins |= ~((1u << s) - 1u);                         // every lane from s on probed and missed
                            j += 32 - s;
                            LZF_STAT(5);
                            break;
                        }
                        if (end == 2) { left_q2 = base + e - 2; LZF_STAT(6); break; }
                        LZF_STAT(7);
                    }
#endif
                    LZF_STAT(8);
                    uint32_t trig, w, cur, cnd, wsum;
                    if (no_dups) {
                        trig = (okmask | endmask) & ~((1u << s) - 1u);
                        if (trig == 0) {
                            ins |= ~((1u << s) - 1u);                         // every lane from s on probed and missed
                            j += 32 - s;
                            break;
                        }
                        w = __ffs(trig) - 1;
                        if ((endmask >> w) & 1u) {
                            // final literal-only sequence  :178-190.  The probes before the end lane did their
                            // table.replace: a later block of the same chain (dependent blocks) sees them.
                            if (a.fin_pos && lane == 0) { a.fin_pos[b] = opos; a.fin_lit[b] = len - lit_start; }
                            if (!emit_sequence(out, opos, cap, in, lit_start, len - lit_start, 0, 0, true, false, 0, 0)) status = LZF_WRITER_FULL;
                            ins |= ((1u << w) - 1u) & ~((1u << s) - 1u);
                            done = true;
                            break;
                        }
                        ins |= ((2u << w) - 1u) & ~((1u << s) - 1u);         // probes s..w happened
                        const uint32_t pw = __shfl_sync(LZF_FULL_MASK, packed_w, w);
                        cur = consecutive ? base + w : __shfl_sync(LZF_FULL_MASK, p, w);
                        cnd = cur - (pw & 0xffffu);
                        wsum = pw >> 16;
                    } else {
                        // ---- candidates as the serial algorithm would see them: lane k of the current run
                        // comes after every insert recorded so far AND after the probes s..k-1 of its own run
                        const uint32_t inb = same & (ins | ~((1u << s) - 1u)) & lower_mask;
                        const uint32_t src = inb ? (31u - __clz(inb)) : lane;
                        const uint32_t pj = __shfl_sync(LZF_FULL_MASK, p, src);
                        const uint32_t vj = __shfl_sync(LZF_FULL_MASK, v32, src);
                        const uint32_t cand = inb ? pj : tcand;
                        const bool ok = !is_end && lane >= s && (inb ? (vj == v32 && p - pj <= 0xffffu) : t_ok);
                        trig = __ballot_sync(LZF_FULL_MASK, (ok || is_end) && lane >= s);
                        if (trig == 0) {
                            ins |= ~((1u << s) - 1u);
                            j += 32 - s;
                            break;
                        }
                        w = __ffs(trig) - 1;
                        if ((endmask >> w) & 1u) {
                            if (a.fin_pos && lane == 0) { a.fin_pos[b] = opos; a.fin_lit[b] = len - lit_start; }
                            if (!emit_sequence(out, opos, cap, in, lit_start, len - lit_start, 0, 0, true, false, 0, 0)) status = LZF_WRITER_FULL;
                            ins |= ((1u << w) - 1u) & ~((1u << s) - 1u);
                            done = true;
                            break;
                        }
                        ins |= ((2u << w) - 1u) & ~((1u << s) - 1u);
                        cur = __shfl_sync(LZF_FULL_MASK, p, w);
                        cnd = __shfl_sync(LZF_FULL_MASK, cand, w);
                        wsum = __shfl_sync(LZF_FULL_MASK, inb ? 0u : fsum, w);   // in-batch candidates take the general extension
                    }

                    const uint32_t limit = len - 5 - cur;                     // bytes of current_batch :195
                    const uint32_t max_back = min(cur - lit_start, cnd);
                    uint32_t matching = 0, backtrack = 0;
                    // ---- fast extension: the winner's own summary, unless it came from an in-batch candidate
                    // or the match / backtrack runs past what the summary covers
                    bool fast = false;
                    {
                        const uint32_t fm = wsum & 31u, nb = (wsum >> 5) & 7u;
                        const bool f_ok = (wsum & 0x100u) && (fm < 16 || limit <= 16);
                        const bool b_ok = max_back == 0 || ((wsum & 0x200u) && (nb < 4 || max_back <= 4));
                        if (f_ok && b_ok) {
                            fast = true;
                            matching = min(fm, limit);
                            backtrack = min(nb, max_back);
                        }
                    }
                    bool emitted = false;
                    if (fast) {
                        // ---- HOT PATH: a short sequence whose literals start inside this batch — token, <= 14
                        // literals (from registers) and the offset are written with one byte per lane
                        const uint32_t L = cur - backtrack - lit_start;
                        const uint32_t extra = matching - 4 + backtrack;
                        const uint32_t cursor = cur + matching;
                        if (L < 15 && extra < 15 && lit_start >= base && cap - opos >= 17 && cap >= opos) {
                            const uint32_t litb = __shfl_sync(LZF_FULL_MASK, v32, lit_start - base + lane - 1) & 0xffu;
                            const uint32_t offset = cur - cnd;
                            const uint32_t v = lane == 0 ? ((L << 4) | extra) : (lane <= L ? litb : (offset >> (8 * (lane - 1 - L))));
                            if (lane < L + 3) out[opos + lane] = (uint8_t)v;
                            opos += L + 3;
                            if (cursor - base < 32) {
                                // ... and the match ends inside the batch too: the parse goes on with its later lanes
                                lit_start = cursor;
                                j = 0;
                                const uint32_t l2 = cursor - 2 - base;            // table.replace(cursor - 2) :218
                                if (!((endmask >> l2) & 1u)) ins |= 1u << l2;
                                else late_q2 = cursor - 2;
                                s = cursor - base;
                                continue;
                            }
                            emitted = true;                                       // the match leaves the batch: commit below
                        }
                    }
                    if (!fast) {
                    // ---- general extension: lanes 0..23 look 96 bytes ahead (count_matching_bytes :117-145 over
                    // input[cur..len-5]), lanes 24..31 look 32 bytes behind (backtrack :211-214)
                    if (gated) {
                        const uint32_t need = len - cur < 128 ? len : cur + 128;
                        if (need > arrived) arrived = wait_arrival(a.progress, a.slice_bytes, need);
                    }
                    uint32_t cnt = 4;                                         // bytes this lane's word contributes
                    if (lane < 24) {
                        const uint32_t f = 4 + 4 * lane;
                        if (f < limit) {
                            const uint32_t x = ld4(in, cur + f) ^ ld4(in, cnd + f);   // cur + f + 4 <= len - 1
                            const uint32_t c = x ? (uint32_t)(__ffs((int)x) - 1) >> 3 : 4u;
                            cnt = min(c, limit - f);
                        } else cnt = 0;
                    } else {
                        const uint32_t t = lane - 24;                         // word t covers bytes [4t, 4t+4) behind
                        if (4 * t < max_back) {
                            const uint32_t avail = min(4u, max_back - 4 * t);
                            uint32_t c = 0;
                            if (avail == 4) {
                                const uint32_t x = ld4(in, cur - 4 * t - 4) ^ ld4(in, cnd - 4 * t - 4);
                                c = x ? (uint32_t)__clz((int)x) >> 3 : 4u;    // equal bytes counted from the top (nearest to cur)
                            } else {
                                while (c < avail && in[cur - 4 * t - 1 - c] == in[cnd - 4 * t - 1 - c]) c++;
                            }
                            cnt = c;
                        } else cnt = 0;
                    }
                    const uint32_t stop = __ballot_sync(LZF_FULL_MASK, cnt != 4u);
                    {
                        const uint32_t fstop = stop & 0x00ffffffu;
                        if (fstop) {
                            const uint32_t fl = __ffs(fstop) - 1;
                            matching = 4 + 4 * fl + __shfl_sync(LZF_FULL_MASK, cnt, fl);
                        } else {
                            // long match: stream on, 8 bytes per lane and step
                            matching = 100;
                            for (;;) {
                                if (gated) {
                                    const uint32_t top = cur + matching;          // this pass reads [top, top + 264)
                                    const uint32_t need = len - top < 264 ? len : top + 264;
                                    if (need > arrived) arrived = wait_arrival(a.progress, a.slice_bytes, need);
                                }
                                const uint32_t idx = matching + lane * 8;
                                uint32_t c = 0;
                                if (idx < limit) {
                                    const uint32_t room = limit - idx;
                                    if (room >= 8) {
                                        const uint32_t x0 = ld4(in, cur + idx) ^ ld4(in, cnd + idx);
                                        const uint32_t x1 = ld4(in, cur + idx + 4) ^ ld4(in, cnd + idx + 4);
                                        c = x0 ? (uint32_t)(__ffs((int)x0) - 1) >> 3 : (x1 ? 4u + ((uint32_t)(__ffs((int)x1) - 1) >> 3) : 8u);
                                    } else {
                                        while (c < room && in[cur + idx + c] == in[cnd + idx + c]) c++;
                                    }
                                }
                                const uint32_t st2 = __ballot_sync(LZF_FULL_MASK, c != 8u);
                                if (st2 == 0) { matching += 256; continue; }
                                const uint32_t f2 = __ffs(st2) - 1;
                                matching += f2 * 8 + __shfl_sync(LZF_FULL_MASK, c, f2);
                                break;
                            }
                        }
                    }
                    {
                        const uint32_t bstop = stop >> 24;
                        if (bstop) {
                            const uint32_t bl = __ffs(bstop) - 1;
                            backtrack = 4 * bl + __shfl_sync(LZF_FULL_MASK, cnt, 24 + bl);
                        } else {
                            backtrack = 32;
                            while (backtrack < max_back) {
                                const uint32_t k = backtrack + lane;
                                const bool eq = k < max_back && in[cur - 1 - k] == in[cnd - 1 - k];
                                const uint32_t st2 = __ballot_sync(LZF_FULL_MASK, !eq);
                                if (st2 == 0) { backtrack += 32; continue; }
                                backtrack += __ffs(st2) - 1;
                                break;
                            }
                        }
                    }
                    }   // !fast
                    const uint32_t extra = matching - 4 + backtrack;          // :206,214
                    const uint32_t cursor = cur + matching;                   // :215

                    // ---- write_group  :150-163,235-236
                    // literal bytes of a run that started inside this batch are the low bytes of the lanes' v32
                    const bool lits_in_regs = consecutive && lit_start >= base;
                    if (!emitted && !emit_sequence(out, opos, cap, in, lit_start, cur - backtrack - lit_start, cur - cnd, extra, false,
                                                   lits_in_regs, v32, lit_start - base)) {
                        // the writer refused (:150-163 via NoPartialWrites): compress2 returns here, AFTER the
                        // table.replace(cursor - 2) of this match (:218) — a later block of the chain sees that state
                        status = LZF_WRITER_FULL;
                        done = true;
                    }
                    lit_start = cursor;
                    j = 0;

                    // ---- table.replace(input, cursor - 2)  :218, and where the parse resumes
                    const uint32_t q2 = cursor - 2;
                    if (consecutive && cursor - base < 32 && !done) {
                        const uint32_t l2 = q2 - base;
                        // a lane past the 12-byte rule has no hash: its insert is applied after the batch's commit
                        // (only a later block of the same chain can ever read it)
                        if (!((endmask >> l2) & 1u)) ins |= 1u << l2;
                        else late_q2 = q2;
                        s = cursor - base;
                        continue;                                             // next sequence of the same batch
                    }
                    // the match left the batch: its cursor - 2 insert goes behind everything the batch commits
                    left_q2 = q2;
                    break;
                }
                {
                    // commit: last inserted lane of each slot wins (mem::swap order, :64-71)
                    const uint32_t later = same & ins & ~lower_mask & ~(1u << lane);
                    if (((ins >> lane) & 1u) && later == 0) table.put(h, p + ab);
                    // ... then the cursor - 2 insert of a match that left the batch, or of one that ended on a lane past
                    // the 12-byte rule (such a lane has no hash; the two never happen in the same batch)
                    const uint32_t q2 = left_q2 != 0xffffffffu ? left_q2 : late_q2;
                    if (q2 != 0xffffffffu) {
                        __syncwarp();
                        if constexpr (kPacked) {
                            // a long match: never let more than 65536 positions pass between sweeps (the table does
                            // not change during a match, so the intermediate sweeps are exact), and keep the insert
                            // below last_sweep + 65536
                            while (q2 + ab - last_sweep >= 65536u) {
                                last_sweep += 65536u;
                                table.sweep(nslots, last_sweep);
                            }
                        }
                        if (lane == 0) {
                            uint32_t h2;
                            if (kHash4) h2 = hash4(ld4(in, q2), hashlog);
                            else if (len - q2 >= 8) h2 = hash5(ld4(in, q2), in[q2 + 4], hashlog);
                            else h2 = 0;                                      // :43 unwrap_or(0) -> hash of 0
                            table.put(h2, q2 + ab);
                        }
                    }
                }
                __syncwarp();
            }
        }

        __syncwarp();
        if (lane == 0) {
            a.out_len[b] = (status == LZF_OK) ? opos : 0u;
            a.status[b] = status;
        }
        if (gated && len > arrived) arrived = wait_arrival(a.progress, a.slice_bytes, len);   // a refused block stops early
        // XXH32 of the plaintext / of the stored bytes (compressed, or the plaintext when stored raw):
        // queued so that 8 blocks are hashed per warp pass
        if (a.xxh_plain) hash_queue_push(&hashq, in + cursor0, own_len, a.xxh_plain + b);
        if (a.xxh_stored) {
            if (status == LZF_OK) hash_queue_push(&hashq, out, opos, a.xxh_stored + b);
            else hash_queue_push(&hashq, in + cursor0, own_len, a.xxh_stored + b);
        }
      }   // blocks of the chain
        if constexpr (!kPacked) {
            if (carried_in) {
                __syncwarp();
                for (uint32_t i = lane; i < nslots; i += 32) a.table_io[i] = (uint32_t)table.t[i];
            }
        }
    }
    hash_queue_finish(&hashq, kEncodeWarpsPerCta);
}

}  // namespace lzf

// Host-side launcher.  Returns a cudaError_t as int.
namespace lzf {
// Which instantiation a batch runs on, and where its tables live.
struct EncodePlan {
    int variant;            // 0: u32 slots, 1: u16 slots, 2: packed 17-bit slots, 3: hash4 (U16Table)
    uint32_t nslots;
    size_t table_bytes;     // per warp
    int warps;              // per CTA
    int n_smem_warps;       // warps of a CTA whose table is in shared memory
    int global_warps_per_sm;   // upper bound of warps per SM that keep their table in the global scratch
};

static EncodePlan plan_encode(const EncodeArgs* args) {
    EncodePlan p;
    const uint32_t hashlog = args->hashlog;
    const bool hash4 = args->table_kind == LZF_TABLE_U16;
    p.nslots = hash4 ? (2u << hashlog) : (1u << hashlog);
    // u16 slots are exact whenever every position fits 16 bits; blocks up to 16 MiB use packed 17-bit slots
    const uint64_t span = args->max_pos ? args->max_pos : args->max_block_len;
    const bool slot16 = hash4 || (span != 0 && span <= 65536u);
    const bool packed = !slot16 && args->max_block_len != 0 && args->max_block_len <= kPacked17MaxLen;
    p.table_bytes = slot16 ? (size_t)p.nslots * 2 : packed ? (size_t)p.nslots * 2 + p.nslots / 8 : (size_t)p.nslots * 4;
    p.variant = hash4 ? 3 : slot16 ? 1 : packed ? 2 : 0;
    bool big = p.variant == 1 || p.variant == 2;
    if (p.variant == 2 && args->tune_u32_slots) {                     // tuning knob: plain u32 slots in the global scratch
        p.variant = 0;
        p.table_bytes = (size_t)p.nslots * 4;
    }
    const size_t smem_budget = kSmemPerCtaMax - sizeof(HashQueue) - 256;
    if (big) {
        // one big CTA per SM; whatever does not fit shared memory goes to the global scratch
        p.warps = kEncodeBigWarps;
        int fit = (int)(smem_budget / p.table_bytes);
        if (fit > kEncodeSmemWarpsMax) fit = kEncodeSmemWarpsMax;
        if (args->tune_smem_warps_p1) {                               // test / tuning knob: fewer shared-memory tables
            const int v = (int)args->tune_smem_warps_p1 - 1;
            if (v < fit) fit = v;
        }
        p.n_smem_warps = fit < p.warps ? fit : p.warps;
        p.global_warps_per_sm = 2 * (p.warps - p.n_smem_warps);       // 2: small tables could make two CTAs resident
    } else {
        // 16 KiB and larger tables (u32 slots, U16Table): CTAs of 4 warps, all in shared memory or all in the scratch
        p.warps = kEncodeWarpsPerCta;
        const bool smem_tables = p.table_bytes <= 32 * 1024 && p.table_bytes * p.warps <= 200 * 1024;
        p.n_smem_warps = smem_tables ? p.warps : 0;
        p.global_warps_per_sm = smem_tables ? 0 : kGlobalTableCtasPerSm * p.warps;
    }
    return p;
}

template <int kTab, bool kHash4, int kWarps>
static int launch_encode_variant(const EncodeArgs* args, int num_sms, const EncodePlan& p, cudaStream_t stream) {
    auto kern = encode_blocks_kernel<kTab, kHash4, kWarps>;
    cudaError_t e;
    int ctas_per_sm = 1;
    const size_t dyn = p.table_bytes * p.n_smem_warps;
    if (dyn > 48 * 1024) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return (int)e;
    }
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, kWarps * 32, dyn);
    if (e != cudaSuccess) return (int)e;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const int gw = kWarps - p.n_smem_warps;                           // global-table warps per CTA
    if (gw > 0 && ctas_per_sm * gw > p.global_warps_per_sm) ctas_per_sm = p.global_warps_per_sm / gw;
    unsigned grid = (unsigned)(num_sms * ctas_per_sm);
    const uint32_t nwork = args->chain_first ? args->nchains : args->nblocks;
    const unsigned need = (nwork + kWarps - 1) / kWarps;
    if (grid > need) grid = need;
    LZF_LAUNCH(kern, grid, kWarps * 32, dyn, stream, *args, p.nslots, p.n_smem_warps);
    return (int)cudaGetLastError();
}
}  // namespace lzf

extern "C" int lzf_launch_encode(const lzf::EncodeArgs* args, int num_sms, cudaStream_t stream) {
    using namespace lzf;
    if (args->nblocks == 0) return 0;
    const EncodePlan p = plan_encode(args);
    if (p.global_warps_per_sm > 0 && args->global_tables == nullptr) return (int)cudaErrorInvalidValue;
    switch (p.variant) {
        case 3: return launch_encode_variant<1, true, kEncodeWarpsPerCta>(args, num_sms, p, stream);
        case 1: return launch_encode_variant<1, false, kEncodeBigWarps>(args, num_sms, p, stream);
        case 2: return launch_encode_variant<2, false, kEncodeBigWarps>(args, num_sms, p, stream);
        default:
            if (p.warps == kEncodeBigWarps) return launch_encode_variant<0, false, kEncodeBigWarps>(args, num_sms, p, stream);
            return launch_encode_variant<0, false, kEncodeWarpsPerCta>(args, num_sms, p, stream);
    }
}

// Bytes of global table scratch a launch with these arguments may touch (0: every table is in shared memory).
extern "C" size_t lzf_encode_global_table_bytes(const lzf::EncodeArgs* args, int num_sms) {
    const lzf::EncodePlan p = lzf::plan_encode(args);
    return (size_t)num_sms * p.global_warps_per_sm * p.table_bytes;
}
