// lzf_decompress.cu — batched LZ4 block decode for sm_100a, one warp per independent block.
//
// Behavioural contract: raw::decompress_raw of the reference (src/raw/decompress.rs:58-138),
// including its error precedence, with an initially empty output Vec and an
// optional `prefix`; stored blocks (bit 31 of the length word, src/framed/decompress.rs:217,
// 249-251) are copied verbatim.  XXH32 of the decoded bytes is fused into the block epilogue.
//
// Mapping: a warp walks the sequence chain of its block.  The compressed stream is held in a
// 128-byte register window (one aligned 32-bit word per lane, refilled with one coalesced load),
// so tokens, offsets and short literal runs are served by warp shuffles instead of dependent
// memory loads; long literal runs and stored blocks use 16-byte vector copies; matches are copied
// lane-parallel with the sequential (overlapping) semantics of copy_overlapping
// (src/raw/decompress.rs:80-138) preserved through modular source indexing.
#include "lzf_kernels.cuh"

namespace lzf {


// 128-byte register window over the compressed stream.
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
    // true when bytes [pos, pos+need) are inside the window
    __device__ __forceinline__ bool covers(uint64_t pos, uint32_t need) const {
        const uintptr_t a = reinterpret_cast<uintptr_t>(base + pos);
        return a >= wa && a + need <= wa + 128;
    }
    // up to 4 bytes at pos (caller guarantees covers(pos, 4) or that the excess is ignored)
    __device__ __forceinline__ uint32_t u32(uint64_t pos) const {
        const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(base + pos) - wa);
        const uint32_t lo = __shfl_sync(LZF_FULL_MASK, reg, a >> 2);
        const uint32_t hi = __shfl_sync(LZF_FULL_MASK, reg, (a >> 2) + 1);   // wraps mod 32: only used when covered
        return __funnelshift_r(lo, hi, (a & 3u) * 8u);
    }
    __device__ __forceinline__ uint32_t u8(uint64_t pos) const {
        const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(base + pos) - wa);
        const uint32_t w = __shfl_sync(LZF_FULL_MASK, reg, a >> 2);
        return (w >> ((a & 3u) * 8u)) & 0xffu;
    }
};

// byte of the history `prefix ++ out` at signed index i relative to out[0]
__device__ __forceinline__ uint8_t hist_byte(const uint8_t* out, const uint8_t* prefix_end, int64_t i) {
    return i >= 0 ? out[i] : prefix_end[i];
}

constexpr int kDecodeWarpsPerCta = 4;

__global__ void __launch_bounds__(kDecodeWarpsPerCta * 32)
decode_blocks_kernel(DecodeArgs a) {
    const unsigned lane = lane_id();
    const uint32_t b = blockIdx.x * kDecodeWarpsPerCta + (threadIdx.x >> 5);
    if (b >= a.nblocks) return;

    const uint32_t len_word = a.in_len[b];
    const uint64_t n = len_word & ~LZF_INCOMPRESSIBLE;
    const uint8_t* in = a.in + a.in_off[b];
    uint8_t* out = a.out + a.out_off[b];
    const uint64_t cap = a.out_cap[b];
    const uint64_t limit = a.out_limit[b];
    const uint64_t plen = a.prefix ? a.prefix_len[b] : 0;
    const uint8_t* prefix_end = a.prefix ? a.prefix + a.prefix_off[b] + plen : nullptr;

    int status = LZF_OK;
    uint64_t olen = 0;

    if (len_word & LZF_INCOMPRESSIBLE) {
        // stored block: output.extend_from_slice(buf)   src/framed/decompress.rs:249-251
        olen = n;
        if (n <= cap) warp_copy(out, in, n);
    } else {
        Window win;
        win.base = in;
        win.end = in + n;
        win.wa = 0;
        win.reg = 0;
        if (n) win.load(0);
        uint64_t pos = 0;
        bool dry = false;   // physical cap exceeded: keep parsing for exact error reporting, stop writing

        while (pos < n) {                                                   // :61
            if (!win.covers(pos, 4)) win.load(pos);
            const uint32_t t4 = win.u32(pos);
            const uint32_t token = t4 & 0xffu;
            pos += 1;
            uint64_t lit = token >> 4;
            if (lit == 15) {                                                // read_lsic :30-43
                for (;;) {
                    if (pos >= n) { status = LZF_UNEXPECTED_END; break; }
                    const uint32_t more = __ldg(in + pos);
                    pos += 1;
                    lit += more;
                    if (more != 0xffu) break;
                }
                if (status) break;
            }
            if (n - pos < lit) { status = LZF_UNEXPECTED_END; break; }      // :67 read_exact
            if (lit) {
                if (!dry && olen + lit > cap) dry = true;
                if (!dry) {
                    if (lit <= 32 && win.covers(pos, (uint32_t)lit)) {
                        // literals straight out of the register window
                        const uint32_t aoff = (uint32_t)(reinterpret_cast<uintptr_t>(in + pos) - win.wa) + lane;
                        const uint32_t w = __shfl_sync(LZF_FULL_MASK, win.reg, aoff >> 2);
                        if (lane < lit) out[olen + lane] = (uint8_t)(w >> ((aoff & 3u) * 8u));
                    } else {
                        warp_copy(out + olen, in + pos, lit);
                    }
                }
                olen += lit;
                pos += lit;
            }
            if (n - pos < 2) { pos = n; break; }                            // :70 (recent-std EOF behaviour)
            if (!win.covers(pos, 4)) win.load(pos);
            const uint32_t o4 = win.u32(pos);
            const uint32_t offset = o4 & 0xffffu;
            pos += 2;
            uint64_t mlen = token & 0xfu;
            if (mlen == 15) {                                               // :71 read_lsic
                for (;;) {
                    if (pos >= n) { status = LZF_UNEXPECTED_END; break; }
                    const uint32_t more = __ldg(in + pos);
                    pos += 1;
                    mlen += more;
                    if (more != 0xffu) break;
                }
                if (status) break;
            }
            mlen += 4;
            if (olen + mlen > limit) { status = LZF_MEMORY_LIMIT_EXCEEDED; break; }      // :72-74
            if (offset == 0) { status = LZF_ZERO_DEDUP_OFFSET; break; }                  // :83
            if (offset > olen && offset - olen > plen) { status = LZF_INVALID_DEDUP_OFFSET; break; }   // :84-89
            if (!dry && olen + mlen > cap) dry = true;
            if (!dry) {
                __syncwarp();   // literal bytes just stored by other lanes are match history
                uint8_t* dst = out + olen;
                const int64_t src0 = (int64_t)olen - (int64_t)offset;
                if (offset >= 32) {
                    // each 32-byte step only reads bytes at least 32 behind its own writes
                    for (uint64_t k0 = 0; k0 < mlen; k0 += 32) {
                        const uint64_t k = k0 + lane;
                        if (k < mlen) dst[k] = hist_byte(out, prefix_end, src0 + (int64_t)k);
                        __syncwarp();
                    }
                } else {
                    // overlapping run: out[olen+k] = hist[olen-offset + (k mod offset)]
                    for (uint64_t k0 = 0; k0 < mlen; k0 += 32) {
                        const uint64_t k = k0 + lane;
                        if (k < mlen) dst[k] = hist_byte(out, prefix_end, src0 + (int64_t)(k % offset));
                    }
                }
            }
            olen += mlen;
            __syncwarp();
        }
    }
    if (status == LZF_OK && olen > cap) status = LZF_OUTPUT_CAP;
    __syncwarp();
    if (a.xxh_plain) {
        uint32_t h = 0;
        if (status == LZF_OK) h = warp_xxh32(out, olen);
        if (lane == 0) a.xxh_plain[b] = h;
    }
    if (lane == 0) {
        a.out_len[b] = (uint32_t)(olen > 0xffffffffull ? 0xffffffffull : olen);
        a.status[b] = status;
    }
}

}  // namespace lzf

extern "C" int lzf_launch_decode(const lzf::DecodeArgs* args, int num_sms, cudaStream_t stream) {
    (void)num_sms;
    if (args->nblocks == 0) return 0;
    const unsigned grid = (args->nblocks + lzf::kDecodeWarpsPerCta - 1) / lzf::kDecodeWarpsPerCta;
    LZF_LAUNCH(lzf::decode_blocks_kernel, grid, lzf::kDecodeWarpsPerCta * 32, 0, stream, *args);
    return (int)cudaGetLastError();
}
