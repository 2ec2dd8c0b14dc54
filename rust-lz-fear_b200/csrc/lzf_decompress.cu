// lzf_decompress.cu — batched LZ4 block decode for sm_100a, one warp per independent block.
//
// Behavioural contract: raw::decompress_raw of the reference (src/raw/decompress.rs:58-138),
// including its error precedence, with an initially empty output Vec and an
// optional `prefix`; stored blocks (bit 31 of the length word, src/framed/decompress.rs:217,
// 249-251) are copied verbatim.  XXH32 of the decoded bytes is fused into the block epilogue.
//
// Mapping (persistent warps, dynamic block queue):
//   * the compressed stream is staged into a per-warp shared-memory window with TMA bulk copies
//     (cp.async.bulk + mbarrier), the following window is prefetched into L2;
//   * FAST PATH, up to 32 sequences per step: the warp walks the token chain in shared memory
//     (the only serial part: one byte load + a few integer ops per sequence), lane k keeps the
//     k-th sequence; a warp scan turns (literal length + match length) into output positions;
//     every lane then copies its own literals and its own match.  Matches that read bytes another
//     match of the same step still has to produce wait for it (round loop ordered by a ballot), so
//     the sequential semantics of copy_overlapping (src/raw/decompress.rs:80-138) are preserved;
//   * the step's output is assembled in a shared-memory staging ring and written to HBM with
//     16-byte vector stores;
//   * SLOW PATH, one sequence per step, warp-cooperative: anything with an LSIC length extension
//     (long literal runs / long matches become 16-byte vectorised warp copies), sequences that
//     touch the end of the block, every error and the out-of-capacity "dry" mode.
#include "lzf_kernels.cuh"

#include <stddef.h>
#include <stdlib.h>

namespace lzf {

#ifndef LZF_DEC_MINCTAS
#define LZF_DEC_MINCTAS 4                  // resident CTAs per SM the register budget is tuned for (64 regs; 5 x 48 spills)
#endif
#ifndef LZF_DEC_WIN
#define LZF_DEC_WIN 1024
#endif
#ifndef LZF_DEC_WARPS
#define LZF_DEC_WARPS 8
#endif
#ifndef LZF_DEC_EARLY_GATHER
#define LZF_DEC_EARLY_GATHER 1
#endif
#ifndef LZF_DEC_GATHER_MAX
#define LZF_DEC_GATHER_MAX 12          // 12 (4 words) or 20 (6 words): longest match fetched ahead of the literal copies
#endif
constexpr int kDecodeWarpsPerCta = LZF_DEC_WARPS;
constexpr uint32_t kWin = LZF_DEC_WIN;     // staged bytes of compressed stream per refill
constexpr uint32_t kStage = 2048;          // output staging ring (power of two, > 32 * 32 + 16)
constexpr uint32_t kFastSeqMax = 3 + 14;   // token + 14 literals + offset: no LSIC byte anywhere

struct __align__(16) DecodeWarpSmem {
    uint8_t win[kWin];            // staged compressed bytes
    uint8_t step[kWin + 32];      // step[p] = encoded size of the LSIC-free sequence whose token is win[p]; 0 = not fast
    uint8_t stage[kStage];        // output staging ring
    uint32_t plist[40];           // token positions of the current step (a group of 8 may overshoot the 32 kept)
    uint64_t mbar;
    uint64_t pad;
};

constexpr int kStepOff = (int)kWin;          // offsetof(DecodeWarpSmem, step)
static_assert(offsetof(DecodeWarpSmem, step) == kStepOff, "step[] follows win[]");

// step[] for the freshly staged window: 4 positions per lane and pass, SIMD-in-a-word.
//   step = 3 + (token >> 4) for a sequence without length extensions that ends inside the window;
//   step = 0 otherwise: either nibble is 15 (the walk works the size out itself if it ever lands there — most
//          such bytes are literals, not tokens, so nothing is spent on them here), or the sequence would end
//          beyond `wend` (window / block end -> refill or slow path).  Adding a 0 step leaves the walk in place.
__device__ __forceinline__ void build_steps(DecodeWarpSmem& sm, uint32_t wlen, uint32_t wend) {
    const unsigned lane = lane_id();
    const uint32_t* w4 = reinterpret_cast<const uint32_t*>(sm.win);
    uint32_t* s4 = reinterpret_cast<uint32_t*>(sm.step);
    const uint32_t nwords = (wlen + 3) >> 2;
    for (uint32_t i = lane; i < nwords; i += 32) {
        const uint32_t w = w4[i];
        const uint32_t special = __vcmpeq4(w & 0xf0f0f0f0u, 0xf0f0f0f0u) | __vcmpeq4(w & 0x0f0f0f0fu, 0x0f0f0f0fu);
        s4[i] = (((w >> 4) & 0x0f0f0f0fu) + 0x03030303u) & ~special;
    }
    __syncwarp();
    // the last kFastSeqMax positions may describe sequences that cross wend; wend itself terminates a walk
    const uint32_t p = wend - min(wend, kFastSeqMax + 1u) + lane;
    if (p <= wend && (p == wend || p + sm.step[p] > wend)) sm.step[p] = 0;
    __syncwarp();
}

// Encoded size of the sequence whose token sits at win[p] when it carries single-byte length extensions
// (literal run 15..269, match 19..273); 0 when an extension continues (0xFF), when the sequence does not
// fit the window, or when it touches the end of the block (all of those: slow path).  Warp-uniform.
__device__ __forceinline__ uint32_t medium_size(const DecodeWarpSmem& sm, uint32_t p, uint32_t wend) {
    if (p + 4 > wend) return 0;
    const uint32_t tok = sm.win[p];
    uint32_t lit = tok >> 4, q = p + 1;
    if (lit == 15u) {
        const uint32_t e = sm.win[q];
        if (e == 255u) return 0;
        lit += e; q++;
    }
    q += lit + 2;                                   // literals + offset
    if (q > wend) return 0;
    if ((tok & 15u) == 15u) {
        if (q + 1 > wend || sm.win[q] == 255u) return 0;
        q++;
    }
    return q - p;
}
constexpr uint32_t kStageBudget = 1400;             // output bytes one step may stage (ring is 2048)
constexpr uint32_t kLaneCopyMax = 20;               // longer literal runs / matches are copied by the whole warp

// 128-byte register window over the compressed stream (slow path).
struct Window {
    const uint8_t* base;   // block start
    const uint8_t* end;    // block end
    uintptr_t wa;          // 4-aligned absolute address of the window start
    uint32_t reg;          // this lane's word: bytes [wa + 4*lane, +4)

    __device__ __forceinline__ void load(uint64_t pos) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(base + pos);
        wa = a & ~uintptr_t(3);
        const uint8_t* w = reinterpret_cast<const uint8_t*>(wa) + 4 * lane_id();
        reg = (w < end) ? __ldg(reinterpret_cast<const uint32_t*>(w)) : 0u;
    }
    __device__ __forceinline__ bool covers(uint64_t pos, uint32_t need) const {
        const uintptr_t a = reinterpret_cast<uintptr_t>(base + pos);
        return a >= wa && a + need <= wa + 128;
    }
    __device__ __forceinline__ uint32_t u32(uint64_t pos) const {
        const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(base + pos) - wa);
        const uint32_t lo = __shfl_sync(LZF_FULL_MASK, reg, a >> 2);
        const uint32_t hi = __shfl_sync(LZF_FULL_MASK, reg, (a >> 2) + 1);   // wraps mod 32: only used when covered
        return __funnelshift_r(lo, hi, (a & 3u) * 8u);
    }
};

// byte of the history `prefix ++ out` at signed index i relative to out[0]
__device__ __forceinline__ uint8_t hist_byte(const uint8_t* out, const uint8_t* prefix_end, int64_t i) {
    return i >= 0 ? out[i] : prefix_end[i];
}

struct BlockState {
    const uint8_t* in; uint64_t n;
    uint8_t* out; uint64_t cap, limit, plen; const uint8_t* prefix_end;
    uint64_t pos, olen;
    int status; bool dry; bool finished;
};

// One sequence, warp-cooperative, exactly the reference's order of checks (decompress.rs:61-75).
__device__ __forceinline__ void slow_sequence(BlockState& s) {
    const unsigned lane = lane_id();
    const uint8_t* in = s.in;
    const uint64_t n = s.n;
    uint8_t* out = s.out;
    Window win;
    win.base = in; win.end = in + n;
    win.load(s.pos);
    uint64_t pos = s.pos, olen = s.olen;
    const uint32_t t4 = win.u32(pos);
    const uint32_t token = t4 & 0xffu;
    pos += 1;
    uint64_t lit = token >> 4;
    if (lit == 15) {                                                    // read_lsic :30-43
        for (;;) {
            if (pos >= n) { s.status = LZF_UNEXPECTED_END; s.pos = pos; return; }
            const uint32_t more = __ldg(in + pos);
            pos += 1;
            lit += more;
            if (more != 0xffu) break;
        }
    }
    if (n - pos < lit) { s.status = LZF_UNEXPECTED_END; s.pos = pos; return; }   // :67 read_exact
    if (lit) {
        if (!s.dry && olen + lit > s.cap) s.dry = true;
        if (!s.dry) warp_copy(out + olen, in + pos, lit);
        olen += lit;
        pos += lit;
    }
    s.olen = olen;
    if (n - pos < 2) { s.pos = n; s.finished = true; return; }          // :70 (recent-std EOF behaviour)
    if (!win.covers(pos, 4)) win.load(pos);
    const uint32_t offset = win.u32(pos) & 0xffffu;
    pos += 2;
    uint64_t mlen = token & 0xfu;
    if (mlen == 15) {                                                   // :71 read_lsic
        for (;;) {
            if (pos >= n) { s.status = LZF_UNEXPECTED_END; s.pos = pos; return; }
            const uint32_t more = __ldg(in + pos);
            pos += 1;
            mlen += more;
            if (more != 0xffu) break;
        }
    }
    mlen += 4;
    s.pos = pos;
    if (olen + mlen > s.limit) { s.status = LZF_MEMORY_LIMIT_EXCEEDED; return; }          // :72-74
    if (offset == 0) { s.status = LZF_ZERO_DEDUP_OFFSET; return; }                        // :83
    if (offset > olen && offset - olen > s.plen) { s.status = LZF_INVALID_DEDUP_OFFSET; return; }   // :84-89
    if (!s.dry && olen + mlen > s.cap) s.dry = true;
    if (!s.dry) {
        __syncwarp();   // literal bytes just stored by other lanes are match history
        uint8_t* dst = out + olen;
        const int64_t src0 = (int64_t)olen - (int64_t)offset;
        if (offset >= 32) {
            if (offset >= mlen && src0 >= 0 && mlen >= 64) {
                warp_copy(dst, out + src0, mlen);                       // non-overlapping: vectorised
            } else {
                // each 32-byte step only reads bytes at least 32 behind its own writes
                for (uint64_t k0 = 0; k0 < mlen; k0 += 32) {
                    const uint64_t k = k0 + lane;
                    if (k < mlen) dst[k] = hist_byte(out, s.prefix_end, src0 + (int64_t)k);
                    __syncwarp();
                }
            }
        } else {
            // overlapping run: out[olen+k] = hist[olen-offset + (k mod offset)]
            for (uint64_t k0 = 0; k0 < mlen; k0 += 32) {
                const uint64_t k = k0 + lane;
                if (k < mlen) dst[k] = hist_byte(out, s.prefix_end, src0 + (int64_t)(k % offset));
            }
        }
    }
    s.olen = olen + mlen;
    __syncwarp();
}

// Writes staged output [from, upto) to global memory.  Output position x lives at stage[x - sbias];
// sbias keeps (x - sbias) congruent to the destination address modulo 16, so aligned 16-byte chunks
// of the stage are aligned 16-byte chunks of the output.  `whole` = false: upto is rounded down to a
// 16-byte boundary of the destination address (the remainder stays staged); returns the new flushed
// position.
__device__ __forceinline__ uint32_t flush_stage(const uint8_t* stage, uint8_t* out, uint32_t sbias, uint32_t from,
                                                uint32_t upto, bool whole) {
    const unsigned lane = lane_id();
    const uintptr_t oa = reinterpret_cast<uintptr_t>(out);
    if (!whole) {
        const uint32_t r = (uint32_t)((oa + upto) & 15u);
        if (upto - from < r) return from;
        upto -= r;
    }
    if (upto <= from) return from;
    uint32_t a = from;
    uint32_t head = (uint32_t)((16u - ((oa + a) & 15u)) & 15u);
    if (head > upto - a) head = upto - a;
    if (lane < head) out[a + lane] = stage[a + lane - sbias];
    a += head;
    const uint32_t nvec = (upto - a) >> 4;
    for (uint32_t v = lane; v < nvec; v += 32) {
        const uint32_t p = a + 16 * v;
        *reinterpret_cast<uint4*>(out + p) = *reinterpret_cast<const uint4*>(stage + (p - sbias));
    }
    a += nvec * 16;
    const uint32_t tail = upto - a;     // only when `whole`
    if (lane < tail) out[a + lane] = stage[a + lane - sbias];
    return upto;
}

// After a flush: re-base the stage so that the (< 16-byte) carry [flushed, olen) starts at the front.
__device__ __forceinline__ uint32_t rebase_stage(uint8_t* stage, const uint8_t* out, uint32_t sbias, uint32_t flushed, uint32_t olen) {
    const unsigned lane = lane_id();
    const uint32_t nb = flushed - (uint32_t)((reinterpret_cast<uintptr_t>(out) + flushed) & 15u);
    if (nb != sbias) {
        const uint32_t carry = olen - flushed;      // < 32
        uint8_t v = 0;
        if (lane < carry) v = stage[flushed + lane - sbias];
        __syncwarp();
        if (lane < carry) stage[flushed + lane - nb] = v;
    }
    return nb;
}

__global__ void __launch_bounds__(kDecodeWarpsPerCta * 32, LZF_DEC_MINCTAS)
decode_blocks_kernel(DecodeArgs a) {
    LZF_DYN_SMEM(smem_raw);
    __shared__ HashQueue hashq;
    const unsigned lane = lane_id();
    DecodeWarpSmem& sm = reinterpret_cast<DecodeWarpSmem*>(smem_raw)[threadIdx.x >> 5];
    if (lane == 0) mbar_init(&sm.mbar, 1);
    hash_queue_init(&hashq);
    uint32_t phase = 0;

    for (;;) {
        uint32_t b = 0;
        if (lane == 0) b = atomicAdd(a.work_counter, 1u);
        b = __shfl_sync(LZF_FULL_MASK, b, 0);
        if (b >= a.nblocks) break;

        const uint32_t len_word = a.in_len[b];
        BlockState s;
        s.n = len_word & ~LZF_INCOMPRESSIBLE;
        s.in = a.in + a.in_off[b];
        s.out = a.out + a.out_off[b];
        s.cap = a.out_cap[b];
        s.limit = a.out_limit[b];
        const bool has_prefix = a.prefix != nullptr || a.prefix_abs;
        s.plen = has_prefix ? a.prefix_len[b] : 0;
        s.prefix_end = !has_prefix ? nullptr
                       : a.prefix_abs ? reinterpret_cast<const uint8_t*>((uintptr_t)a.prefix_off[b]) + s.plen
                                      : a.prefix + a.prefix_off[b] + s.plen;
        if (a.wait_for) {
            // dependent block (src/framed/decompress.rs:238-269): its window is the output of the block
            // before it; the dynamic queue hands blocks out in ascending order, so the predecessor is
            // already running (or done) on some warp
            const int32_t w = a.wait_for[b];
            if (w >= 0) {
                if (lane == 0) while (ld_acquire_gpu(a.done + w) == 0) spin_pause();
                __syncwarp();
                __threadfence();
            }
        }
        s.pos = 0; s.olen = 0; s.status = LZF_OK; s.dry = false; s.finished = false;

        if (len_word & LZF_INCOMPRESSIBLE) {
            // stored block: output.extend_from_slice(buf)   src/framed/decompress.rs:249-251
            s.olen = s.n;
            if (s.n <= s.cap) warp_copy(s.out, s.in, s.n);
        } else {
            // positions q are relative to the 16-byte aligned address at or below the block start
            const uintptr_t in_addr = reinterpret_cast<uintptr_t>(s.in);
            const uintptr_t a0 = in_addr & ~uintptr_t(15);
            const uint32_t q0 = (uint32_t)(in_addr - a0);
            const uint64_t qn = q0 + s.n;
            const uint64_t qn16 = (qn + 15) & ~uint64_t(15);
            uint64_t wq = 0;          // window start (multiple of 16), relative to a0
            uint32_t wlen = 0;        // staged bytes
            uint32_t flushed = 0;     // output below this position is in global memory; [flushed, olen) is staged
            uint32_t sbias = 0u - (uint32_t)(reinterpret_cast<uintptr_t>(s.out) & 15u);   // stage index of position x is x - sbias
            // the fast path works in 32-bit output positions and never exceeds this bound
            const uint64_t bound64 = s.cap < s.limit ? s.cap : s.limit;
            const uint32_t bound = bound64 > 0xfffff000ull ? 0xfffff000u : (uint32_t)bound64;

            while (s.pos < s.n && s.status == LZF_OK && !s.finished) {
                const uint64_t q = q0 + s.pos;
                // ---- (re)fill the staged window
                if (q < wq || q + kFastSeqMax > wq + wlen) {
                    const uint64_t want = q & ~uint64_t(15);
                    const uint64_t avail = qn16 - want;
                    const uint32_t nbytes = avail < kWin ? (uint32_t)avail : kWin;
                    if (want != wq || nbytes != wlen) {
                        __syncwarp();
                        if (lane == 0) {
                            bulk_load(sm.win, reinterpret_cast<const uint8_t*>(a0) + want, nbytes, &sm.mbar);
                            if (avail > kWin) {
                                const uint64_t more = avail - kWin;
                                bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(a0) + want + kWin, more < kWin ? (uint32_t)more : kWin);
                            }
                        }
                        mbar_wait(&sm.mbar, phase);
                        phase ^= 1u;
                        wq = want;
                        wlen = nbytes;
                        build_steps(sm, wlen, (uint32_t)(((wq + wlen) < qn ? (wq + wlen) : qn) - wq));
                    }
                }
                // window-relative positions from here on
                uint32_t p = (uint32_t)(q - wq);
                const uint32_t wend = (uint32_t)(((wq + wlen) < qn ? (wq + wlen) : qn) - wq);

                // ---- walk: up to 32 sequences that lie completely inside the window (plain ones, and ones whose
                // length extensions are single bytes).  This is the only serial part of the decoder: one
                // shared-memory byte per plain sequence.
                uint32_t cnt = 0;
                {
                    // Three instructions per plain sequence (LDS step, STS slot, IADD): positions are kept as offsets
                    // into the CTA's shared memory so that one register addresses the step byte, and a 0 step (not a
                    // plain sequence) simply leaves the walk where it is — no test inside a group of 8.
                    const uint32_t qbase = smem_off(smem_raw, &sm);             // slot value = qbase + window position
                    uint32_t q = qbase + p, d = 1, nk = 0;
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        if (d != 0) {
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                d = lds_u8<kStepOff>(smem_raw, q);
                                sm.plist[c * 8 + i] = q;
                                q += d;
                            }
                            nk = c * 8 + 8;
                        }
                    }
                    __syncwarp();
                    const uint32_t dm = lane < nk ? lds_u8<kStepOff>(smem_raw, sm.plist[lane]) : 0u;
                    cnt = __popc(__ballot_sync(LZF_FULL_MASK, dm != 0));        // a prefix of the lanes
                    p = q - qbase;
                    if (cnt < 32) {
                        // stuck: a sequence with single-byte length extensions is sized by medium_size and the walk
                        // goes on behind it with counted slots (text: ~6 % of the sequences); anything else ends the step
                        uint32_t* kp = sm.plist + cnt;
                        for (;;) {
                            const uint32_t m = medium_size(sm, p, wend);
                            if (m == 0) break;
                            *kp++ = qbase + p;
                            p += m;
                            if (kp >= sm.plist + 32) break;
                            q = qbase + p;
                            for (;;) {
#pragma unroll
                                for (int i = 0; i < 8; i++) {
                                    d = lds_u8<kStepOff>(smem_raw, q);
                                    *kp = q;
                                    q += d;
                                    kp += d != 0;
                                }
                                if (d == 0 || kp >= sm.plist + 32) break;
                            }
                            p = q - qbase;
                            if (kp >= sm.plist + 32) break;
                        }
                        cnt = (uint32_t)(kp - sm.plist);
                        if (cnt > 32) { cnt = 32; p = sm.plist[32] - qbase; }     // a group overshot: sequence 32 starts the next step
                    }
                    __syncwarp();
                    if (lane < cnt) sm.plist[lane] -= qbase;                      // back to window positions
                }
                __syncwarp();
                const uint32_t my_p = sm.plist[lane];
                // ---- per-lane decode of the sequence headers, output positions by warp scan
                uint32_t lit = 0, ml = 0, off = 1, tot = 0, lit_src = 0;
                if (lane < cnt) {
                    const uint32_t tok = sm.win[my_p];
                    lit = tok >> 4;
                    lit_src = my_p + 1;
                    if (lit == 15u) { lit += sm.win[lit_src]; lit_src++; }
                    const uint32_t op = lit_src + lit;
                    off = (uint32_t)sm.win[op] | ((uint32_t)sm.win[op + 1] << 8);
                    ml = tok & 15u;
                    if (ml == 15u) ml += sm.win[op + 2];
                    ml += 4u;
                    tot = lit + ml;
                }
                uint32_t inc = tot;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(LZF_FULL_MASK, inc, d);
                    if (lane >= (unsigned)d) inc += t;
                }
                const uint32_t olen0 = (uint32_t)s.olen;
                const uint32_t o_k = olen0 + inc - tot;                       // output position of my literals
                const uint32_t dstp = o_k + lit;                               // ... and of my match
                // A step stops before: a sequence whose offset is zero or reaches outside prefix ++ output
                // (decompress.rs:83-89) or that would pass the output limit / capacity — the slow path replays it
                // and reports the error exactly like the reference — and before the staging budget runs out.
                const bool bad = lane < cnt && (off == 0u || (uint64_t)off > (uint64_t)dstp + s.plen ||
                                                s.olen + inc > bound || (inc > kStageBudget && lane > 0));
                const uint32_t fb = __ballot_sync(LZF_FULL_MASK, bad);
                uint32_t in_end = p;
                if (fb) {
                    const uint32_t keep = (uint32_t)(__ffs(fb) - 1);
                    if (keep < cnt) {                                           // the next sequence starts where the kept ones end
                        cnt = keep;
                        in_end = sm.plist[keep];
                    }
                }
                if (cnt == 0) {
                    flushed = flush_stage(sm.stage, s.out, sbias, flushed, (uint32_t)s.olen, true);
                    __syncwarp();
                    slow_sequence(s);
                    flushed = (uint32_t)s.olen;
                    sbias = flushed - (uint32_t)((reinterpret_cast<uintptr_t>(s.out) + flushed) & 15u);
                    continue;
                }
                const bool act = lane < cnt;
                const uint32_t out_end = __shfl_sync(LZF_FULL_MASK, o_k + tot, cnt - 1);

                const uint32_t maxml = warp_max_u32((act && ml <= kLaneCopyMax) ? ml : 0u);   // longest lane-copied match of the step
#if LZF_DEC_EARLY_GATHER
                // ---- matches whose source lies completely in the flushed history (global memory) and is short: their
                // bytes are asked for NOW, as aligned words, so that the round trip runs under the literal copies below
                // instead of in the dependency rounds.  g_mis: 4 = not taken; else the source's misalignment (0..3)
                const int64_t srcp_e = (int64_t)dstp - (int64_t)off;
                // (up to 12 bytes: measured on B200, config 2 435 GiB/s against 421 with 16 and 375 without)
                uint32_t gw0 = 0, gw1 = 0, gw2 = 0, gw3 = 0, g_mis = 4;
#if LZF_DEC_GATHER_MAX > 12
                uint32_t gw4 = 0, gw5 = 0;
#endif
                // Only in steps whose lane-copied matches ALL fit the words fetched here: a step that also has longer ones
                // would run this path and the byte-wise one below (text: 151 instead of 166 GiB/s when it did).
                if (maxml <= (uint32_t)LZF_DEC_GATHER_MAX && act && ml <= (uint32_t)LZF_DEC_GATHER_MAX && srcp_e >= 0 &&
                    srcp_e + (int64_t)ml <= (int64_t)flushed) {
                    const uintptr_t ga = reinterpret_cast<uintptr_t>(s.out + srcp_e);
                    const uint32_t* gp = reinterpret_cast<const uint32_t*>(ga & ~uintptr_t(3));
                    g_mis = (uint32_t)(ga & 3u);
                    const uint32_t span = g_mis + ml;                         // bytes from the first aligned word on
                    gw0 = gp[0];
                    if (span > 4) gw1 = gp[1];
                    if (span > 8) gw2 = gp[2];
                    if (span > 12) gw3 = gp[3];
#if LZF_DEC_GATHER_MAX > 12
                    if (span > 16) gw4 = gp[4];
                    if (span > 20) gw5 = gp[5];
#endif
                }
#endif
                // ---- literals: every lane copies its own run into the staging ring
                {
                    // trip counts follow the longest run / match of the step (warp-uniform), four bytes per check
                    const uint32_t mylit = (act && lit <= kLaneCopyMax) ? lit : 0u;
                    const uint32_t maxlit = warp_max_u32(mylit);
                    const uint8_t* src = sm.win + lit_src;
                    uint8_t* d = sm.stage + (o_k - sbias);
#pragma unroll
                    for (uint32_t c = 0; c < kLaneCopyMax; c += 4) {
                        if (c >= maxlit) break;
#pragma unroll
                        for (uint32_t i = c; i < c + 4; i++) if (i < mylit) d[i] = src[i];
                    }
                }
                for (uint32_t lm = __ballot_sync(LZF_FULL_MASK, act && lit > kLaneCopyMax); lm; lm &= lm - 1) {
                    const uint32_t k = __ffs(lm) - 1;                           // a long run: the whole warp copies it
                    const uint32_t n = __shfl_sync(LZF_FULL_MASK, lit, k);
                    const uint8_t* src = sm.win + __shfl_sync(LZF_FULL_MASK, lit_src, k);
                    uint8_t* d = sm.stage + (__shfl_sync(LZF_FULL_MASK, o_k, k) - sbias);
                    for (uint32_t i = lane; i < n; i += 32) d[i] = src[i];
                }
                __syncwarp();

                // ---- matches: a lane may go once everything its source range needs is complete
                const int64_t srcp = (int64_t)dstp - (int64_t)off;
                uint32_t pending = __ballot_sync(LZF_FULL_MASK, act);
                while (pending) {
                    const uint32_t first = __ffs(pending) - 1;
                    const uint32_t dst_first = __shfl_sync(LZF_FULL_MASK, dstp, first);
                    const uint32_t ml_first = __shfl_sync(LZF_FULL_MASK, ml, first);
                    if (ml_first > kLaneCopyMax) {
                        // a long match at the head of the queue: the whole warp copies it, 32 bytes per pass when
                        // the offset allows, as the periodic pattern it is when source and destination overlap closely
                        const uint32_t off_f = __shfl_sync(LZF_FULL_MASK, off, first);
                        const int64_t src_f = (int64_t)dst_first - (int64_t)off_f;
                        uint8_t* d = sm.stage + (dst_first - sbias);
                        for (uint32_t k0 = 0; k0 < ml_first; k0 += 32) {
                            const uint32_t k = k0 + lane;
                            if (k < ml_first) {
                                const int64_t x = src_f + (off_f >= 32 ? (int64_t)k : (int64_t)(k % off_f));
                                d[k] = x >= (int64_t)flushed ? sm.stage[(uint32_t)x - sbias] : hist_byte(s.out, s.prefix_end, x);
                            }
                            if (off_f >= 32) __syncwarp();
                        }
                        __syncwarp();
                        pending &= ~(1u << first);
                        continue;
                    }
                    const bool ready = ((pending >> lane) & 1u) && ml <= kLaneCopyMax &&
                                       (lane == first || srcp + (int64_t)ml <= (int64_t)dst_first);
                    if (ready) {
                        uint8_t* d = sm.stage + (dstp - sbias);
                        if (srcp >= (int64_t)flushed) {
                            const uint8_t* sp = sm.stage + ((uint32_t)srcp - sbias);   // staged -> staged (in order: overlap-safe)
                            for (uint32_t i = 0; i < ml; i++) d[i] = sp[i];
#if LZF_DEC_EARLY_GATHER
                        } else if (g_mis < 4) {
                            // the words fetched before the literal copies: up to 12 source bytes realigned in registers
                            const uint32_t sh = g_mis * 8u;
                            const uint32_t u0 = __funnelshift_r(gw0, gw1, sh), u1 = __funnelshift_r(gw1, gw2, sh), u2 = __funnelshift_r(gw2, gw3, sh);
#if LZF_DEC_GATHER_MAX > 12
                            const uint32_t u3 = __funnelshift_r(gw3, gw4, sh), u4 = __funnelshift_r(gw4, gw5, sh);
#else
                            const uint32_t u3 = 0, u4 = 0;
#endif
#pragma unroll
                            for (uint32_t c = 0; c < (uint32_t)LZF_DEC_GATHER_MAX; c += 4) {
                                if (c >= maxml) break;
                                const uint32_t u = c == 0 ? u0 : c == 4 ? u1 : c == 8 ? u2 : c == 12 ? u3 : u4;
#pragma unroll
                                for (uint32_t i = 0; i < 4; i++) if (c + i < ml) d[c + i] = (uint8_t)(u >> (8u * i));
                            }
#endif
                        } else if (srcp >= 0 && srcp + (int64_t)ml <= (int64_t)flushed) {
                            const uint8_t* g = s.out + srcp;                      // flushed history -> staged
                            // loads first, then stores: one round trip for matches up to 12 bytes, two beyond
                            uint8_t v[12];
#pragma unroll
                            for (uint32_t c = 0; c < 12; c += 4) {
                                if (c >= maxml) break;
#pragma unroll
                                for (uint32_t i = c; i < c + 4; i++) if (i < ml) v[i] = g[i];
                            }
#pragma unroll
                            for (uint32_t c = 0; c < 12; c += 4) {
                                if (c >= maxml) break;
#pragma unroll
                                for (uint32_t i = c; i < c + 4; i++) if (i < ml) d[i] = v[i];
                            }
                            if (maxml > 12) {
#pragma unroll
                                for (uint32_t i = 12; i < kLaneCopyMax; i++) if (i < ml) d[i] = g[i];
                            }
                        } else {
                            for (uint32_t i = 0; i < ml; i++) {                   // straddles the flush point or the prefix
                                const int64_t x = srcp + i;
                                d[i] = x >= (int64_t)flushed ? sm.stage[(uint32_t)x - sbias] : hist_byte(s.out, s.prefix_end, x);
                            }
                        }
                    }
                    __syncwarp();
                    pending &= ~__ballot_sync(LZF_FULL_MASK, ready);
                }
                s.olen = out_end;
                s.pos += in_end - (uint32_t)(q - wq);
                flushed = flush_stage(sm.stage, s.out, sbias, flushed, out_end, false);
                __syncwarp();
                sbias = rebase_stage(sm.stage, s.out, sbias, flushed, out_end);
                __syncwarp();
            }
            flushed = flush_stage(sm.stage, s.out, sbias, flushed, (uint32_t)s.olen, true);
        }
        if (s.status == LZF_OK && s.olen > s.cap) s.status = LZF_OUTPUT_CAP;
        __syncwarp();
        if (lane == 0) {
            a.out_len[b] = (uint32_t)(s.olen > 0xffffffffull ? 0xffffffffull : s.olen);
            a.status[b] = s.status;
        }
        if (a.done) {
            __threadfence();
            __syncwarp();
            if (lane == 0) st_release_gpu(a.done + b, 1u);
        }
        if (a.xxh_plain) {
            // XXH32 of the decoded bytes: queued so that 8 blocks are hashed per warp pass
            if (s.status == LZF_OK) hash_queue_push(&hashq, s.out, s.olen, a.xxh_plain + b);
            else if (lane == 0) a.xxh_plain[b] = 0;
        }
    }
    hash_queue_finish(&hashq, kDecodeWarpsPerCta);
}

}  // namespace lzf

extern "C" int lzf_launch_decode(const lzf::DecodeArgs* args, int num_sms, cudaStream_t stream) {
    using namespace lzf;
    if (args->nblocks == 0) return 0;
    const size_t dyn = sizeof(DecodeWarpSmem) * kDecodeWarpsPerCta;
    cudaError_t e = cudaFuncSetAttribute(decode_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return (int)e;
    int ctas_per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, decode_blocks_kernel, kDecodeWarpsPerCta * 32, dyn);
    if (e != cudaSuccess) return (int)e;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (args->tune_ctas_per_sm >= 1 && (int)args->tune_ctas_per_sm < ctas_per_sm) ctas_per_sm = (int)args->tune_ctas_per_sm;   // tuning knob
    unsigned grid = (unsigned)(num_sms * ctas_per_sm);
    const unsigned need = (args->nblocks + kDecodeWarpsPerCta - 1) / kDecodeWarpsPerCta;
    if (grid > need) grid = need;
    LZF_LAUNCH(decode_blocks_kernel, grid, kDecodeWarpsPerCta * 32, dyn, stream, *args);
    return (int)cudaGetLastError();
}
