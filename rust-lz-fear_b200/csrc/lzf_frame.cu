// lzf_frame.cu — device side of the frame glue around the block kernels:
//   * XXH32 over byte ranges (block checksums of stored payloads, content checksums of frames)
//   * compress: per-frame layout scan + parallel assembly of `u32 len | payload | [u32 xxh]`
//     (src/framed/compress.rs:244-263,277-281)
//   * decompress: walking the block-length words of device-resident frames
//     (src/framed/decompress.rs:205-235)
#include "lzf_kernels.cuh"
#include "lzf_frame.cuh"

namespace lzf {

// ------------------------------------------------------------------------------------------
// XXH32 of ranges: 8 ranges per warp (4 lanes = the 4 accumulator chains of one range)
// ------------------------------------------------------------------------------------------
// One warp per CTA at 32 registers: such a CTA still fits on an SM whose register file is otherwise taken by the
// 28-warp block-encode CTA (28 x 32 x 72 of 65 536 registers), so frame content checksums overlap the encode kernel.
//
// The four accumulator chains of one range are serial (~13 cycles per 16-byte stripe), so the only way to hash a
// long range fast is to keep its bytes from ever being waited for: each range streams through a double-buffered
// 2 KiB shared-memory window filled by TMA bulk copies (cp.async.bulk + mbarrier) while the previous window is
// being hashed.  Loading straight from global memory instead (8 stripes in flight per lane) measured 5.2 ms per
// MiB on B200 — every group of 8 stripes paid a full memory round trip.
// Measured on B200, one 64 MiB range (ms): 2 KiB windows x 2 in flight 43.7 (58 before the rolling register window),
// x 4 in flight 43.8, with an L2 prefetch 16 windows ahead 45.0, 4 KiB x 2 38.8, 8 KiB x 2 36.3 — the chain, not memory,
// is the limit (IMAD, SHF, IMAD ~ 15 cycles per stripe), so long ranges take larger windows (fewer window switches) and
// batches of short ranges keep 2 KiB ones (more CTAs per SM: 4096 x 256 KiB in 0.36 ms instead of 0.83).
constexpr uint32_t kHashDepth = 2;                    // windows of one range in flight
template <uint32_t kHashWin>
struct __align__(16) RangeHashSmem {
    uint8_t buf[8][kHashDepth][kHashWin];
    uint64_t bar[8][kHashDepth];
};

// The stripe phase of one range per lane group (j = group, a = accumulator of the lane) through the group's double-
// buffered shared-memory windows; every lane of the warp calls it (nwin_max = the longest range of the warp).
template <uint32_t kHashWin>
__device__ __forceinline__ uint32_t hash_windows(RangeHashSmem<kHashWin>& sm, unsigned j, unsigned a, const uint8_t* p, uint64_t nstripes,
                                                 uint64_t nwin, uint64_t nwin_max, uint32_t acc) {
    if (a == 0) for (uint32_t d = 0; d < kHashDepth; d++) mbar_init(&sm.bar[j][d], 1);
    __syncwarp();
    auto issue = [&](uint64_t w) {                                                // leader lane of the range
        const uint64_t o = w * kHashWin;
        const uint64_t left = nstripes * 16 - o;
        bulk_load(sm.buf[j][w % kHashDepth], p + o, (uint32_t)(left < kHashWin ? left : kHashWin), &sm.bar[j][w % kHashDepth]);
    };
    if (a == 0) for (uint64_t w = 0; w < kHashDepth && w < nwin; w++) issue(w);
    __syncwarp();
    for (uint64_t w = 0; w < nwin_max; w++) {
        if (w < nwin) {
#ifndef LZF_SIMT_EMU   // (the CPU test harness copies synchronously, and its wait is a warp collective)
            mbar_wait(&sm.bar[j][w % kHashDepth], (uint32_t)((w / kHashDepth) & 1));
#endif
            const uint64_t left = nstripes - w * (kHashWin / 16);
            const uint32_t ns = (uint32_t)(left < kHashWin / 16 ? left : kHashWin / 16);
            const uint32_t* q = reinterpret_cast<const uint32_t*>(sm.buf[j][w % kHashDepth]) + a;
            // the four chains are serial (IMAD, SHF, IMAD per stripe): the words of the next 8 stripes are loaded
            // while the current 8 are hashed, so the chain never waits for shared memory either
            uint32_t s = 0;
            if (ns >= 8) {
                uint32_t x[8];                                                    // a rolling window: each word is reloaded right after its round
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = q[i * 4];
                for (; s + 16 <= ns; s += 8) {
#pragma unroll
                    for (int i = 0; i < 8; i++) { acc = xxh_round(acc, x[i]); x[i] = q[(s + 8 + i) * 4]; }
                }
#pragma unroll
                for (int i = 0; i < 8; i++) acc = xxh_round(acc, x[i]);
                s += 8;
            }
            for (; s < ns; s++) acc = xxh_round(acc, q[s * 4]);
        }
        __syncwarp();                                                             // the window is free again
        if (a == 0 && w + kHashDepth < nwin) issue(w + kHashDepth);
    }
    return acc;
}

template <uint32_t kHashWin>
#ifdef LZF_SIMT_EMU
__global__ void
#else
__global__ void __maxnreg__(32)
#endif
xxh32_ranges_kernel(const uint8_t* data, const uint64_t* off, const uint64_t* len, uint32_t nranges, uint32_t* hash) {
    LZF_DYN_SMEM(smem_raw);
    RangeHashSmem<kHashWin>& sm = *reinterpret_cast<RangeHashSmem<kHashWin>*>(smem_raw);
    const uint32_t warp = blockIdx.x;
    const unsigned lane = lane_id();
    const unsigned j = lane >> 2, a = lane & 3u;
    const uint32_t r = warp * 8 + j;
    const bool valid = r < nranges;
    const uint8_t* p = valid ? data + off[r] : nullptr;
    const uint64_t n = valid ? len[r] : 0;
    // TMA wants 16-byte aligned sources; anything else (and short ranges) takes the plain path
    const bool streamable = (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
    if (__ballot_sync(LZF_FULL_MASK, !streamable) != 0 || warp_max_u32((uint32_t)(n > 0xffffffffull ? 0xffffffffu : n)) < 4 * kHashWin) {
        const uint32_t h = warp_xxh32_x8(p, n);
        if (valid && a == 0) hash[r] = h;
        return;
    }
    const uint64_t nstripes = n >> 4;
    const uint64_t nwin = (nstripes * 16 + kHashWin - 1) / kHashWin;              // windows of this range
    uint64_t nwin_max = nwin;
    for (int sft = 16; sft; sft >>= 1) {                                          // 64-bit max over the warp
        const uint64_t o = __shfl_xor_sync(LZF_FULL_MASK, nwin_max, sft);
        nwin_max = o > nwin_max ? o : nwin_max;
    }
    const uint32_t acc = hash_windows<kHashWin>(sm, j, a, p, nstripes, nwin, nwin_max, xxh32_seed_acc(a));
    const uint32_t h = warp_xxh32_x8_finish(acc, p, n);
    if (valid && a == 0) hash[r] = h;
}

// Stripe phase only, one warp, carrying a running state: acc[0..4) in/out (streaming content checksums of the
// Read/Write-style host API, src/framed/compress.rs:233-235).  Long inputs stream through the same TMA windows as
// xxh32_ranges_kernel (loading the stripes straight from global memory cost 5.2 ms per MiB).
template <uint32_t kHashWin>
#ifdef LZF_SIMT_EMU
__global__ void
#else
__global__ void __maxnreg__(32)
#endif
xxh32_stream_kernel(const uint8_t* data, uint64_t nstripes, uint32_t* acc_io) {
    LZF_DYN_SMEM(smem_raw);
    RangeHashSmem<kHashWin>& sm = *reinterpret_cast<RangeHashSmem<kHashWin>*>(smem_raw);
    const unsigned lane = lane_id();
    const unsigned j = lane >> 2, a = lane & 3u;
    uint32_t acc = lane < 4 ? acc_io[lane] : 0u;
    if ((reinterpret_cast<uintptr_t>(data) & 15u) != 0 || nstripes * 16 < 4 * kHashWin) {
        acc = warp_xxh32_stripes(data, nstripes, acc);
    } else {
        const uint64_t nwin = j == 0 ? (nstripes * 16 + kHashWin - 1) / kHashWin : 0;
        const uint64_t nwin_max = (nstripes * 16 + kHashWin - 1) / kHashWin;
        acc = hash_windows<kHashWin>(sm, j, a, data, j == 0 ? nstripes : 0, nwin, nwin_max, acc);
    }
    if (lane < 4) acc_io[lane] = acc;
}

constexpr uint32_t kSliceBytes = 64 * 1024;

// compress with a dictionary: block k's input becomes dst[dst_off[k]] = dictionary ++ block, so that the
// history compress2 sees in front of the block (in_buffer = block_initializer ++ block, compress.rs:218,268)
// is physically contiguous.  grid = (blocks, 64 KiB slices of the block).
__global__ void __launch_bounds__(256) stage_dict_kernel(StageArgs a) {
    const uint32_t k = blockIdx.x;
    uint8_t* dst = a.dst + a.dst_off[k];
    const unsigned warp = threadIdx.x >> 5;
    if (blockIdx.y == 0) {
        for (uint64_t w0 = (uint64_t)warp * 8192; w0 < a.dlen; w0 += 8 * 8192)
            warp_copy(dst + w0, a.dict + w0, min((uint64_t)8192, (uint64_t)a.dlen - w0));
    }
    const uint64_t s0 = (uint64_t)blockIdx.y * kSliceBytes + (uint64_t)warp * 8192;
    if (s0 < a.len[k]) warp_copy(dst + a.dlen + s0, a.in + a.src_off[k] + s0, min((uint64_t)8192, (uint64_t)a.len[k] - s0));
}

// ------------------------------------------------------------------------------------------
// compress: layout.  One warp per frame scans its blocks' stored sizes, decides the position of
// every block inside the frame, and writes header, EndMark and content checksum.
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(128) frame_layout_kernel(LayoutArgs a) {
    const unsigned lane = lane_id();
    const uint32_t f = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (f >= a.nframes) return;
    const uint32_t b0 = a.first_block[f], nb = a.nblocks[f];
    const uint32_t hlen = a.headers[(size_t)f * 20];
    const uint64_t per_block_extra = 4 + (a.block_checksums ? 4 : 0);

    // pass 1: total size + panic detection
    uint64_t total = 0;
    int bad = 0;
    for (uint32_t i = lane; i < nb; i += 32) {
        const int st = a.blk_status[b0 + i];
        const uint64_t stored = (st == LZF_OK) ? a.blk_comp_len[b0 + i] : a.blk_in_len[b0 + i];
        if (st != LZF_OK && st != LZF_WRITER_FULL) bad = 1;
        total += stored + per_block_extra;
    }
    for (int s = 16; s; s >>= 1) total += __shfl_xor_sync(LZF_FULL_MASK, total, s);
    bad = __any_sync(LZF_FULL_MASK, bad);
    const uint64_t flen = hlen + total + 4 + (a.content_checksum ? 4 : 0);
    int status = LZF_F_OK;
    if (bad) status = LZF_F_PANIC;
    else if (flen > a.out_cap[f]) status = LZF_F_WRITE_ERROR;

    // pass 2: exclusive scan -> block positions
    uint64_t run = a.out_off[f] + hlen;
    for (uint32_t i0 = 0; i0 < nb; i0 += 32) {
        const uint32_t i = i0 + lane;
        uint64_t sz = 0;
        if (i < nb) {
            const int st = a.blk_status[b0 + i];
            sz = ((st == LZF_OK) ? a.blk_comp_len[b0 + i] : a.blk_in_len[b0 + i]) + per_block_extra;
        }
        uint64_t inc = sz;
        for (int s = 1; s < 32; s <<= 1) {
            const uint64_t t = __shfl_up_sync(LZF_FULL_MASK, inc, s);
            if (lane >= (unsigned)s) inc += t;
        }
        if (i < nb) a.blk_dst[b0 + i] = (status == LZF_F_OK) ? run + inc - sz : ~0ull;
        run += __shfl_sync(LZF_FULL_MASK, inc, 31);
    }
    if (status == LZF_F_OK) {
        uint8_t* o = a.out + a.out_off[f];
        if (lane < hlen) o[lane] = a.headers[(size_t)f * 20 + 1 + lane];        // compress.rs:200
        uint8_t* t = a.out + run;
        if (lane < 4) t[lane] = 0;                                              // EndMark compress.rs:277
        if (a.content_checksum && lane < 4) t[4 + lane] = (uint8_t)(a.content_hash[f] >> (8 * lane));   // :279-281
    }
    if (lane == 0) {
        a.frame_len[f] = (status == LZF_F_OK) ? flen : 0;
        a.frame_status[f] = status;
    }
}

// ------------------------------------------------------------------------------------------
// compress: assembly.  grid = (blocks, slices); every CTA moves one 64 KiB slice of one block's
// stored payload into the frame; slice 0 also writes the length word and the block checksum.
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) frame_assemble_kernel(AssembleArgs a) {
    const uint32_t b = blockIdx.x;
    const uint64_t dst = a.blk_dst[b];
    if (dst == ~0ull) return;
    const int st = a.blk_status[b];
    const bool compressed = st == LZF_OK;
    const uint32_t stored = compressed ? a.blk_comp_len[b] : a.blk_in_len[b];
    const uint8_t* src = compressed ? a.comp + a.blk_comp_off[b] : a.in + a.blk_in_off[b];
    uint8_t* o = a.out + dst;
    const uint64_t s0 = (uint64_t)blockIdx.y * kSliceBytes;
    if (blockIdx.y == 0 && threadIdx.x < 4) {
        const uint32_t word = compressed ? stored : (stored | LZF_INCOMPRESSIBLE);   // compress.rs:247,253
        o[threadIdx.x] = (uint8_t)(word >> (8 * threadIdx.x));
        if (a.blk_xxh_stored) o[4 + (uint64_t)stored + threadIdx.x] = (uint8_t)(a.blk_xxh_stored[b] >> (8 * threadIdx.x));   // :259-263
    }
    if (s0 >= stored) return;
    const uint64_t slice = min((uint64_t)kSliceBytes, (uint64_t)stored - s0);
    // 8 warps x 8 KiB
    const unsigned warp = threadIdx.x >> 5;
    const uint64_t w0 = (uint64_t)warp * 8192;
    if (w0 < slice) warp_copy(o + 4 + s0 + w0, src + s0 + w0, min((uint64_t)8192, slice - w0));
}

// ------------------------------------------------------------------------------------------
// decompress: frame walk.  One thread per frame parses the header and chases the block-length
// words (decompress.rs:205-235).  mode 0 = count blocks, mode 1 = also fill block descriptors.
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t rd32_bytes(const uint8_t* p) {
    return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24);
}

__global__ void __launch_bounds__(64) frame_walk_kernel(WalkArgs a) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.nframes) return;
    const uint8_t* in = a.in + a.in_off[f];
    const uint64_t n = a.in_len[f];
    WalkFrame w;
    lzf_frame_info info;
    int32_t detail = 0;
    w.header_status = parse_frame_header(in, n, &info, &detail);
    w.header_detail = detail;
    w.flags = info.flags; w.block_maxsize = info.block_maxsize; w.nblocks = 0;
    w.term_status = LZF_F_INPUT_ERROR; w.content_checksum = 0; w.consumed = 0;
    w.content_size = info.content_size; w.dictionary_id = info.dictionary_id;
    w.has_fields = (info.has_content_size ? 1u : 0u) | (info.has_dictionary_id ? 2u : 0u);
    if (w.header_status == LZF_F_OK) {
        uint64_t p = info.header_len;
        const uint64_t bms = info.block_maxsize;
        const bool bc = (info.flags & kFlagBlockChecksums) != 0;
        const uint32_t b0 = a.mode ? a.first_block[f] : 0;
        uint32_t i = 0;
        for (;;) {
            if (n - p < 4) { w.term_status = LZF_F_INPUT_ERROR; break; }             // :205
            const uint32_t word = rd32_bytes(in + p);
            p += 4;
            if (word == 0) {                                                         // :206-215
                if (info.flags & kFlagContentChecksum) {
                    if (n - p < 4) { w.term_status = LZF_F_INPUT_ERROR; break; }
                    w.content_checksum = rd32_bytes(in + p);
                    p += 4;
                }
                w.term_status = LZF_F_OK;
                break;
            }
            const uint32_t blen = word & ~LZF_INCOMPRESSIBLE;                        // :217-218
            if (blen > (uint32_t)bms) { w.term_status = LZF_F_BLOCK_SIZE_OVERFLOW; break; }   // :220-222
            if (n - p < blen) { w.term_status = LZF_F_INPUT_ERROR; break; }          // :226
            const uint64_t payload = p;
            p += blen;
            uint32_t cks = 0;
            if (bc) {                                                                // :228-229
                if (n - p < 4) { w.term_status = LZF_F_INPUT_ERROR; break; }
                cks = rd32_bytes(in + p);
                p += 4;
            }
            if (a.mode) {
                const uint32_t b = b0 + i;
                a.blk_in_off[b] = a.in_off[f] + payload;
                a.blk_len_word[b] = word;
                a.blk_payload_len[b] = blen;
                a.blk_checksum[b] = cks;
                a.blk_end[b] = p;
                const uint64_t rel = (uint64_t)i * bms;
                const uint64_t capf = a.out_cap[f];
                a.blk_out_off[b] = a.out_off[f] + (rel < capf ? rel : capf);
                a.blk_out_cap[b] = (uint32_t)(rel < capf ? min(bms, capf - rel) : 0);
                a.blk_out_limit[b] = (uint32_t)bms;                                  // :248 output_limit = block_maxsize
            }
            i++;
        }
        w.nblocks = i;
        w.consumed = p;
    }
    if (a.mode == 0) a.frames[f] = w;
}

// ------------------------------------------------------------------------------------------
// segmented parse: descriptors of the segments, and stitching their sequence streams into one block
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) segment_plan_kernel(SegmentPlanArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.nblocks * a.nseg) return;
    const uint32_t b = i / a.nseg, k = i % a.nseg;
    const uint64_t start = (uint64_t)k * a.seg_len;
    const uint32_t n = a.in_len[b];
    a.seg_in_off[i] = a.in_off[b] + (start < n ? start : n);
    a.seg_in_len[i] = start < n ? (uint32_t)min((uint64_t)a.seg_len, (uint64_t)n - start) : 0u;
    a.seg_prefix[i] = start < n ? (uint32_t)min(start, (uint64_t)LZF_WINDOW_SIZE) : 0u;   // the window in front of the segment primes its table
    a.seg_out_off[i] = (uint64_t)i * a.seg_cap;
    a.seg_out_cap[i] = a.seg_cap;
    a.seg_chain_first[i] = i; a.seg_chain_count[i] = 1; a.seg_abs[i] = 0;
}

__device__ __forceinline__ uint32_t lsic_bytes(uint32_t v) { return v < 15 ? 0u : (v - 15u) / 255u + 1u; }
__device__ __forceinline__ uint8_t* put_lsic(uint8_t* o, uint32_t v) {          // write_lsic_tail, compress/mod.rs:243-260
    if (v < 15) return o;
    v -= 15;
    while (v >= 255) { *o++ = 0xff; v -= 255; }
    *o++ = (uint8_t)v;
    return o;
}

// One thread per block walks its S segment streams in order.  Every stream is a complete LZ4 block of its own: it ends
// in a literal-only sequence.  Those closing literals become the head of the next stream's first sequence, whose token
// and length bytes are rewritten for the joint count (a segment without any match just passes its bytes on); the block
// ends with one closing sequence.  Literal bytes are always taken from the plaintext: they ARE the input bytes.
// Pass 0 sizes everything (NoPartialWrites: a block that does not fit its capacity is refused as a whole, compress.rs:
// 242-256,298-301), pass 1 writes the headers and the copy jobs.
__global__ void __launch_bounds__(64) stitch_plan_kernel(StitchArgs a) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.nblocks) return;
    const uint32_t n = a.in_len[b];
    const uint32_t cap = a.out_cap ? a.out_cap[b] : n;
    uint8_t* out = a.out + a.out_off[b];
    StitchJob* jobs = a.jobs + (size_t)b * (a.nseg + 1);
    const uint64_t in0 = a.in_off[b];
    int status = LZF_OK;
    uint64_t total = 0;
    for (int pass = 0; pass < 2; pass++) {
        uint64_t o = 0;            // output position
        uint32_t carry = 0;        // literals waiting for the next match: the input bytes right in front of `pos`
        uint64_t pos = 0;          // input position where the current segment starts
        for (uint32_t k = 0; k <= a.nseg; k++) {
            StitchJob j{0, 0, 0, 0, 0, 0};
            const uint32_t i = b * a.nseg + k;
            const uint64_t left = pos < n ? (uint64_t)n - pos : 0;
            const uint32_t slen = k < a.nseg ? (uint32_t)(left < a.seg_len ? left : a.seg_len) : 0u;
            if (k < a.nseg && slen) {
                if (a.seg_status[i] != LZF_OK) status = LZF_PANIC;       // (scratch is sized by the worst case: cannot happen)
                const uint8_t* x = a.seg + a.seg_out_off[i];
                const uint32_t f = a.fin_pos[i];
                if (f == 0) {
                    carry += slen;                                       // no match in this segment: all of it is literals
                } else {
                    // first sequence: token, literal length A (read_lsic), header bytes h
                    const uint32_t tok = x[0];
                    uint32_t A = tok >> 4, h = 1;
                    if (A == 15) { uint32_t e; do { e = x[h++]; A += e; } while (e == 255); }
                    const uint32_t L = carry + A;
                    const uint64_t hdr = 1 + lsic_bytes(L);
                    if (pass) {
                        uint8_t* q = out + o;
                        *q++ = (uint8_t)(((L < 15 ? L : 15) << 4) | (tok & 15u));
                        put_lsic(q, L);
                        j.lit_src = in0 + pos - carry; j.lit_dst = a.out_off[b] + o + hdr; j.lit_len = carry;
                        j.body_src = a.seg_out_off[i] + h; j.body_dst = a.out_off[b] + o + hdr + carry; j.body_len = f - h;
                    }
                    o += hdr + carry + (f - h);
                    carry = a.fin_lit[i];
                }
                pos += slen;
            } else if (k == a.nseg && n) {
                // the block's closing literal-only sequence (:178-190)
                const uint64_t hdr = 1 + lsic_bytes(carry);
                if (pass) {
                    uint8_t* q = out + o;
                    *q++ = (uint8_t)((carry < 15 ? carry : 15) << 4);
                    put_lsic(q, carry);
                    j.lit_src = in0 + n - carry; j.lit_dst = a.out_off[b] + o + hdr; j.lit_len = carry;
                }
                o += hdr + carry;
            }
            if (pass) jobs[k] = j;
        }
        if (pass == 0) {
            total = o;
            if (status == LZF_OK && total > cap) status = LZF_WRITER_FULL;
            if (status != LZF_OK) {
                for (uint32_t k = 0; k <= a.nseg; k++) jobs[k] = StitchJob{0, 0, 0, 0, 0, 0};
                break;
            }
        }
    }
    a.out_len[b] = status == LZF_OK ? (uint32_t)total : 0u;
    a.status[b] = status;
    if (a.hash_off) {        // what the frame stores for this block: the compressed bytes, or the plaintext (compress.rs:259-263)
        a.hash_off[b] = status == LZF_OK ? (uint64_t)(uintptr_t)(a.out + a.out_off[b]) : (uint64_t)(uintptr_t)(a.in + in0);
        a.hash_len[b] = status == LZF_OK ? total : n;
    }
    if (a.plain_off) { a.plain_off[b] = (uint64_t)(uintptr_t)(a.in + in0); a.plain_len[b] = n; }
}

// grid = (blocks * (S + 1) jobs, 64 KiB slices): the carried literals (from the plaintext) and the body (from the segment's stream)
__global__ void __launch_bounds__(256) stitch_copy_kernel(StitchArgs a) {
    const StitchJob j = a.jobs[blockIdx.x];
    const unsigned warp = threadIdx.x >> 5;
    const uint32_t longest = j.lit_len > j.body_len ? j.lit_len : j.body_len;
    for (uint64_t s0 = (uint64_t)blockIdx.y * kSliceBytes + (uint64_t)warp * 8192; s0 < longest; s0 += (uint64_t)gridDim.y * kSliceBytes) {
        if (s0 < j.lit_len) warp_copy(a.out + j.lit_dst + s0, a.in + j.lit_src + s0, min((uint64_t)8192, (uint64_t)j.lit_len - s0));
        if (s0 < j.body_len) warp_copy(a.out + j.body_dst + s0, a.seg + j.body_src + s0, min((uint64_t)8192, (uint64_t)j.body_len - s0));
    }
}

}  // namespace lzf

extern "C" int lzf_launch_segment_plan(const lzf::SegmentPlanArgs* a, cudaStream_t s) {
    const uint32_t n = a->nblocks * a->nseg;
    if (!n) return 0;
    LZF_LAUNCH(lzf::segment_plan_kernel, (n + 127) / 128, 128, 0, s, *a);
    return (int)cudaGetLastError();
}
extern "C" int lzf_launch_stitch(const lzf::StitchArgs* a, cudaStream_t s) {
    if (!a->nblocks) return 0;
    LZF_LAUNCH(lzf::stitch_plan_kernel, (a->nblocks + 63) / 64, 64, 0, s, *a);
    // a job usually moves one segment's stream; literals that several all-literal segments passed on take more trips
    uint32_t slices = (uint32_t)(((uint64_t)a->seg_len + lzf::kSliceBytes - 1) / lzf::kSliceBytes);
    if (slices == 0) slices = 1;
    dim3 grid(a->nblocks * (a->nseg + 1), slices);
    LZF_LAUNCH(lzf::stitch_copy_kernel, grid, 256, 0, s, *a);
    return (int)cudaGetLastError();
}

namespace lzf {
template <uint32_t kHashWin>
static int launch_xxh32_ranges(const uint8_t* data, const uint64_t* off, const uint64_t* len, uint32_t nranges, uint32_t* hash, cudaStream_t s) {
    const size_t dyn = sizeof(RangeHashSmem<kHashWin>);
    if (dyn > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(xxh32_ranges_kernel<kHashWin>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return (int)e;
    }
    LZF_LAUNCH(xxh32_ranges_kernel<kHashWin>, (nranges + 7) / 8, 32, dyn, s, data, off, len, nranges, hash);
    return (int)cudaGetLastError();
}
}  // namespace lzf

// long_ranges: the caller knows that the ranges are few and long (whole frames of megabytes): 4 KiB windows — 64 KiB
// of shared memory per CTA still fits beside the 28-warp encode CTA, 8 KiB windows (36.3 ms per 64 MiB) would not
extern "C" int lzf_launch_xxh32_ranges(const uint8_t* data, const uint64_t* off, const uint64_t* len,
                                       uint32_t nranges, uint32_t* hash, cudaStream_t s, int long_ranges) {
    if (!nranges) return 0;
    return long_ranges ? lzf::launch_xxh32_ranges<4096>(data, off, len, nranges, hash, s)
                       : lzf::launch_xxh32_ranges<2048>(data, off, len, nranges, hash, s);
}
extern "C" int lzf_launch_xxh32_stripes(const uint8_t* data, uint64_t nstripes, uint32_t* acc, cudaStream_t s) {
    const size_t dyn = sizeof(lzf::RangeHashSmem<4096>);
    const cudaError_t e = cudaFuncSetAttribute(lzf::xxh32_stream_kernel<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return (int)e;
    LZF_LAUNCH(lzf::xxh32_stream_kernel<4096>, 1, 32, dyn, s, data, nstripes, acc);
    return (int)cudaGetLastError();
}
extern "C" int lzf_launch_stage_dict(const lzf::StageArgs* a, uint32_t max_block_len, cudaStream_t s) {
    if (!a->n) return 0;
    uint32_t slices = (max_block_len + lzf::kSliceBytes - 1) / lzf::kSliceBytes;   // one CTA per 64 KiB of the longest block
    if (slices == 0) slices = 1;
    dim3 grid(a->n, slices);
    LZF_LAUNCH(lzf::stage_dict_kernel, grid, 256, 0, s, *a);
    return (int)cudaGetLastError();
}
extern "C" int lzf_launch_layout(const lzf::LayoutArgs* a, cudaStream_t s) {
    if (!a->nframes) return 0;
    LZF_LAUNCH(lzf::frame_layout_kernel, (a->nframes + 3) / 4, 128, 0, s, *a);
    return (int)cudaGetLastError();
}
extern "C" int lzf_launch_assemble(const lzf::AssembleArgs* a, uint32_t max_block_len, cudaStream_t s) {
    if (!a->nblocks) return 0;
    uint32_t slices = (max_block_len + lzf::kSliceBytes - 1) / lzf::kSliceBytes;
    if (slices == 0) slices = 1;
    dim3 grid(a->nblocks, slices);
    LZF_LAUNCH(lzf::frame_assemble_kernel, grid, 256, 0, s, *a);
    return (int)cudaGetLastError();
}
extern "C" int lzf_launch_walk(const lzf::WalkArgs* a, cudaStream_t s) {
    if (!a->nframes) return 0;
    LZF_LAUNCH(lzf::frame_walk_kernel, (a->nframes + 63) / 64, 64, 0, s, *a);
    return (int)cudaGetLastError();
}
