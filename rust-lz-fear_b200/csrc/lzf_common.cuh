// lzf_common.cuh — device helpers shared by the sm_100a kernels.
//
// All helpers are warp-synchronous: every lane of a full 32-lane warp calls them with the same
// (warp-uniform) arguments unless stated otherwise.
#pragma once

#include <stdint.h>
#ifndef LZF_SIMT_EMU   // tests/simt/simt_emu.h pre-defines the CUDA vocabulary for the CPU SIMT test harness
#include <cuda_runtime.h>
#define LZF_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define LZF_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#endif

#include "../../include/lzfear_b200.h"

#define LZF_FULL_MASK 0xffffffffu

namespace lzf {

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// ---- unaligned little-endian loads built from aligned 32-bit loads -------------------------
// An aligned word that contains at least one byte of an allocation lies entirely inside it
// (cudaMalloc / caching-allocator granularity >= 256 B), so the only over-read these helpers can
// perform is within the first/last word of the range, never past the allocation.

// 32 bits at byte address p (any alignment). Reads the word holding p and, if needed, the next.
__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    const unsigned sh = (unsigned)(a & 3u) * 8u;
    const uint32_t lo = w[0];
    if (sh == 0) return lo;
    return __funnelshift_r(lo, w[1], sh);
}

// 64 bits at byte address p (any alignment). `end` = one past the last readable byte of the
// range; the third word is only touched when it still holds a byte below `end`.
__device__ __forceinline__ uint64_t ld_u64_unaligned(const uint8_t* p, const uint8_t* end) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    const unsigned sh = (unsigned)(a & 3u) * 8u;
    const uint32_t w0 = w[0];
    const uint32_t w1 = w[1];
    uint32_t w2 = 0;
    if (sh != 0 && reinterpret_cast<const uint8_t*>(w + 2) < end) w2 = w[2];
    const uint32_t lo = __funnelshift_r(w0, w1, sh);
    const uint32_t hi = __funnelshift_r(w1, w2, sh);
    return (uint64_t(hi) << 32) | lo;
}

// ---- XXH32 (public algorithm; twox-hash XxHash32 in the reference) --------------------------
constexpr uint32_t XP1 = 2654435761u, XP2 = 2246822519u, XP3 = 3266489917u, XP4 = 668265263u, XP5 = 374761393u;

__device__ __forceinline__ uint32_t rotl32(uint32_t x, unsigned r) { return __funnelshift_l(x, x, r); }
__device__ __forceinline__ uint32_t xxh_round(uint32_t acc, uint32_t x) { return rotl32(acc + x * XP2, 13) * XP1; }

// Stripe phase of XXH32 by one warp: lanes 0..3 each own one accumulator of the 16-byte stripes.
// `acc` holds the lane's incoming accumulator (lanes >= 4: ignored); returns the updated one.
__device__ __forceinline__ uint32_t warp_xxh32_stripes(const uint8_t* p, size_t nstripes, uint32_t acc) {
    const unsigned lane = lane_id();
    if (lane < 4) {
        const uint8_t* q = p + 4 * lane;
        const bool aligned = (reinterpret_cast<uintptr_t>(p) & 3u) == 0;
        size_t s = 0;
        if (aligned) {
            const uint32_t* qw = reinterpret_cast<const uint32_t*>(q);
            for (; s + 8 <= nstripes; s += 8) {
                uint32_t x[8];
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = qw[(s + i) * 4];
#pragma unroll
                for (int i = 0; i < 8; i++) acc = xxh_round(acc, x[i]);
            }
            for (; s < nstripes; s++) acc = xxh_round(acc, qw[s * 4]);
        } else {
            for (; s + 4 <= nstripes; s += 4) {
                uint32_t x[4];
#pragma unroll
                for (int i = 0; i < 4; i++) x[i] = ld_u32_unaligned(q + (s + i) * 16);
#pragma unroll
                for (int i = 0; i < 4; i++) acc = xxh_round(acc, x[i]);
            }
            for (; s < nstripes; s++) acc = xxh_round(acc, ld_u32_unaligned(q + s * 16));
        }
    }
    return acc;
}

__device__ __forceinline__ uint32_t xxh32_seed_acc(unsigned lane) {
    return lane == 0 ? XP1 + XP2 : lane == 1 ? XP2 : lane == 2 ? 0u : lane == 3 ? 0u - XP1 : 0u;
}

// XXH32(seed 0) of p[0..n) computed by one warp; the tail and avalanche run on lane 0.
// Returns the hash in every lane.
__device__ __forceinline__ uint32_t warp_xxh32(const uint8_t* p, size_t n) {
    const unsigned lane = lane_id();
    const size_t nstripes = n >> 4;
    const uint32_t acc = warp_xxh32_stripes(p, nstripes, xxh32_seed_acc(lane));
    const uint32_t a0 = __shfl_sync(LZF_FULL_MASK, acc, 0);
    const uint32_t a1 = __shfl_sync(LZF_FULL_MASK, acc, 1);
    const uint32_t a2 = __shfl_sync(LZF_FULL_MASK, acc, 2);
    const uint32_t a3 = __shfl_sync(LZF_FULL_MASK, acc, 3);
    uint32_t h = 0;
    if (lane == 0) {
        if (n >= 16) h = rotl32(a0, 1) + rotl32(a1, 7) + rotl32(a2, 12) + rotl32(a3, 18);
        else h = XP5;
        h += (uint32_t)n;
        const uint8_t* t = p + (nstripes << 4);
        unsigned rem = (unsigned)(n & 15);
        while (rem >= 4) {
            // byte-wise gather keeps us inside [p, p+n)
            uint32_t x = uint32_t(t[0]) | (uint32_t(t[1]) << 8) | (uint32_t(t[2]) << 16) | (uint32_t(t[3]) << 24);
            h = rotl32(h + x * XP3, 17) * XP4;
            t += 4; rem -= 4;
        }
        while (rem) { h = rotl32(h + uint32_t(*t) * XP5, 11) * XP1; t++; rem--; }
        h ^= h >> 15; h *= XP2;
        h ^= h >> 13; h *= XP3;
        h ^= h >> 16;
    }
    return __shfl_sync(LZF_FULL_MASK, h, 0);
}

// XXH32(seed 0) of EIGHT byte ranges at once: lane = 4 * j + a hashes accumulator a of range j
// (pointer/length are per lane group; a group with len 0 still yields XXH32("")).  The four
// accumulator chains of one range are inherently serial, so a single range can only ever keep
// 4 lanes busy; packing 8 independent ranges (blocks / frames) into one warp fills all 32.
// Returns the hash of range j in all four lanes of group j.
// Tail and avalanche of warp_xxh32_x8: `acc` = accumulator a of range j after all 16-byte stripes.
__device__ __forceinline__ uint32_t warp_xxh32_x8_finish(uint32_t acc, const uint8_t* p, uint64_t n) {
    const unsigned lane = lane_id();
    const unsigned a = lane & 3u;
    const uint64_t nstripes = n >> 4;
    const unsigned g = lane & ~3u;
    const uint32_t a0 = __shfl_sync(LZF_FULL_MASK, acc, g);
    const uint32_t a1 = __shfl_sync(LZF_FULL_MASK, acc, g + 1);
    const uint32_t a2 = __shfl_sync(LZF_FULL_MASK, acc, g + 2);
    const uint32_t a3 = __shfl_sync(LZF_FULL_MASK, acc, g + 3);
    uint32_t h = 0;
    if (a == 0) {
        if (n >= 16) h = rotl32(a0, 1) + rotl32(a1, 7) + rotl32(a2, 12) + rotl32(a3, 18);
        else h = XP5;
        h += (uint32_t)n;
        const uint8_t* t = p + (nstripes << 4);
        unsigned rem = (unsigned)(n & 15);
        while (rem >= 4) {
            uint32_t x = uint32_t(t[0]) | (uint32_t(t[1]) << 8) | (uint32_t(t[2]) << 16) | (uint32_t(t[3]) << 24);
            h = rotl32(h + x * XP3, 17) * XP4;
            t += 4; rem -= 4;
        }
        while (rem) { h = rotl32(h + uint32_t(*t) * XP5, 11) * XP1; t++; rem--; }
        h ^= h >> 15; h *= XP2;
        h ^= h >> 13; h *= XP3;
        h ^= h >> 16;
    }
    return __shfl_sync(LZF_FULL_MASK, h, g);
}

__device__ __forceinline__ uint32_t warp_xxh32_x8(const uint8_t* p, uint64_t n) {
    const unsigned lane = lane_id();
    const unsigned a = lane & 3u;
    const uint64_t nstripes = n >> 4;
    uint32_t acc = xxh32_seed_acc(a);
    {
        const uint8_t* q = p + 4 * a;
        uint64_t s = 0;
        if ((reinterpret_cast<uintptr_t>(p) & 3u) == 0) {
            // the loads of group g + 1 are in flight while group g is hashed (the chains are serial, ~13 cycles per
            // stripe).  What the fused checksum costs the decode kernel is its memory traffic, not this loop: config 2
            // runs at 421 GiB/s without it and 373 with it, with or without an L2 prefetch ahead of the loads.
            const uint32_t* qw = reinterpret_cast<const uint32_t*>(q);
            uint32_t x[8], y[8];
            if (nstripes >= 8) {
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = qw[i * 4];
                for (s = 8; s + 16 <= nstripes; s += 16) {
#pragma unroll
                    for (int i = 0; i < 8; i++) y[i] = qw[(s + i) * 4];
#pragma unroll
                    for (int i = 0; i < 8; i++) acc = xxh_round(acc, x[i]);
#pragma unroll
                    for (int i = 0; i < 8; i++) x[i] = qw[(s + 8 + i) * 4];
#pragma unroll
                    for (int i = 0; i < 8; i++) acc = xxh_round(acc, y[i]);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) acc = xxh_round(acc, x[i]);
            }
            for (; s < nstripes; s++) acc = xxh_round(acc, qw[s * 4]);
        } else {
            for (; s < nstripes; s++) acc = xxh_round(acc, ld_u32_unaligned(q + s * 16));
        }
    }
    return warp_xxh32_x8_finish(acc, p, n);
}

// ---- per-CTA queue of finished blocks whose XXH32 is still owed -------------------------------
// A warp that finishes a block pushes (pointer, length, destination slot); the warp that pushes
// the 8th entry of a group hashes the whole group with warp_xxh32_x8; the last warp to leave the
// CTA hashes the remainder.  Keeps the checksum fused in the block kernel at 1/8 of the issue cost.
#ifdef LZF_SIMT_EMU
__device__ __forceinline__ void spin_pause() { simt::yield(); }
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) { return *(const volatile uint32_t*)p; }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) { return *(const volatile uint32_t*)p; }
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) { *(volatile uint32_t*)p = v; }
#else
__device__ __forceinline__ void spin_pause() { __nanosleep(32); }
// gpu-scope acquire / release on a flag in global memory (the acquire also drops stale L1 lines)
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// system-scope acquire: the writer is the copy engine (a host-issued cudaMemcpyAsync on another stream)
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
#endif

constexpr uint32_t kHashRing = 64;     // entries; a power of two, multiple of 8
struct HashEntry { const uint8_t* p; uint64_t n; uint32_t* dst; uint32_t pad; };
struct HashQueue {
    uint32_t count;                    // tickets handed out
    uint32_t warps_done;
    volatile uint32_t ready[kHashRing];   // ticket + 1 once the entry is written
    HashEntry e[kHashRing];
};

__device__ __forceinline__ void hash_queue_init(HashQueue* q) {
    if (threadIdx.x == 0) { q->count = 0; q->warps_done = 0; }
    for (uint32_t i = threadIdx.x; i < kHashRing; i += blockDim.x) q->ready[i] = 0;
    __syncthreads();
}

// hashes tickets [first, first + cnt), cnt <= 8
__device__ __forceinline__ void hash_queue_run(HashQueue* q, uint32_t first, uint32_t cnt) {
    const unsigned lane = lane_id();
    const uint32_t j = lane >> 2;
    const uint8_t* p = nullptr; uint64_t n = 0; uint32_t* dst = nullptr;
    if (j < cnt) {
        const uint32_t slot = (first + j) & (kHashRing - 1);
        while (q->ready[slot] != first + j + 1) spin_pause();
        __threadfence_block();
        p = q->e[slot].p; n = q->e[slot].n; dst = q->e[slot].dst;
    }
    __syncwarp();                                          // every lane has seen its flag and copied its entry
    if (j < cnt && (lane & 3u) == 0) q->ready[(first + j) & (kHashRing - 1)] = 0;   // slot may be reused
    const uint32_t h = warp_xxh32_x8(p, n);
    if (j < cnt && (lane & 3u) == 0 && dst) *dst = h;
}

// warp-uniform call; `dst` == nullptr entries are skipped by the caller
__device__ __forceinline__ void hash_queue_push(HashQueue* q, const uint8_t* p, uint64_t n, uint32_t* dst) {
    const unsigned lane = lane_id();
    uint32_t t = 0;
    if (lane == 0) {
        t = atomicAdd(&q->count, 1u);
        const uint32_t slot = t & (kHashRing - 1);
        // a slot is reused only after its previous occupant (ticket t - kHashRing) has been consumed
        while (q->ready[slot] != 0) spin_pause();
        q->e[slot].p = p; q->e[slot].n = n; q->e[slot].dst = dst;
        __threadfence_block();
        q->ready[slot] = t + 1;
    }
    t = __shfl_sync(LZF_FULL_MASK, t, 0);
    if ((t & 7u) == 7u) hash_queue_run(q, t - 7u, 8u);
}

// called once per warp when it leaves the kernel's work loop
__device__ __forceinline__ void hash_queue_finish(HashQueue* q, uint32_t nwarps) {
    const unsigned lane = lane_id();
    uint32_t d = 0, total = 0;
    __syncwarp();
    if (lane == 0) { __threadfence_block(); d = atomicAdd(&q->warps_done, 1u); total = *(volatile uint32_t*)&q->count; }
    d = __shfl_sync(LZF_FULL_MASK, d, 0);
    total = __shfl_sync(LZF_FULL_MASK, total, 0);
    if (d == nwarps - 1 && (total & 7u)) hash_queue_run(q, total & ~7u, total & 7u);
}

// ---- warp-wide byte copy, non-overlapping, any alignment -----------------------------------
// Copies n bytes src -> dst.  Short runs go byte-per-lane; long runs align the destination to
// 16 bytes and move one uint4 per lane per step, re-aligning the source with funnel shifts.
__device__ __forceinline__ void warp_copy(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, size_t n) {
    const unsigned lane = lane_id();
    if (n < 128) {
        for (size_t i = lane; i < n; i += 32) dst[i] = src[i];
        return;
    }
    // head: bring dst to 16-byte alignment
    const unsigned head = (unsigned)((16u - (reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u);
    if (lane < head) dst[lane] = src[lane];
    dst += head; src += head; n -= head;
    const size_t nvec = n >> 4;
    const uintptr_t sa = reinterpret_cast<uintptr_t>(src);
    const unsigned sh = (unsigned)(sa & 3u) * 8u;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(sa & ~uintptr_t(3));
    uint4* dv = reinterpret_cast<uint4*>(dst);
    if (sh == 0 && (sa & 15u) == 0) {
        const uint4* sv = reinterpret_cast<const uint4*>(src);
        for (size_t v = lane; v < nvec; v += 32) dv[v] = sv[v];
    } else if (sh == 0) {
        for (size_t v = lane; v < nvec; v += 32) {
            const uint32_t* s4 = sw + v * 4;
            dv[v] = make_uint4(s4[0], s4[1], s4[2], s4[3]);
        }
    } else {
        // the 5th word of the last vector holds at least one byte < src+n only if a tail exists;
        // guard it so we never touch a word fully past the range
        const uint8_t* send = src + n;
        for (size_t v = lane; v < nvec; v += 32) {
            const uint32_t* s4 = sw + v * 4;
            const uint32_t w0 = s4[0], w1 = s4[1], w2 = s4[2], w3 = s4[3];
            uint32_t w4 = 0;
            if (reinterpret_cast<const uint8_t*>(s4 + 4) < send) w4 = s4[4];
            dv[v] = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh),
                               __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
        }
    }
    const size_t done = nvec << 4;
    const unsigned tail = (unsigned)(n - done);
    if (lane < tail) dst[done + lane] = src[done + lane];
}

// ---- shared memory addressed by its 32-bit offset ---------------------------------------------
// A chain of dependent shared-memory loads (the decoder's token walk) wants ONE register that is both the loop
// variable and the load address.  `base` = the CTA's dynamic shared memory; offsets are relative to it in the
// CPU test harness and are shared-window addresses on the device.
#ifndef LZF_SIMT_EMU
__device__ __forceinline__ uint32_t smem_off(const uint8_t*, const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int kImm>
__device__ __forceinline__ uint32_t lds_u8(const uint8_t*, uint32_t off) {      // byte at off + kImm
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(off), "n"(kImm) : "memory");
    return v;
}
#else
__device__ __forceinline__ uint32_t smem_off(const uint8_t* base, const void* p) { return (uint32_t)((const uint8_t*)p - base); }
template <int kImm>
__device__ __forceinline__ uint32_t lds_u8(const uint8_t* base, uint32_t off) { return base[off + kImm]; }
#endif

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers ------------------------------
// One elected lane arms the barrier with the byte count and issues the copy; every lane of the
// warp then waits on the barrier's phase parity.  Sizes and both addresses are multiples of 16.
#ifndef LZF_SIMT_EMU
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    // order earlier generic-proxy accesses of the destination before the async-proxy write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ uint32_t warp_max_u32(uint32_t v) { return __reduce_max_sync(LZF_FULL_MASK, v); }
#else
// CPU SIMT test harness: the copy is synchronous, the barrier is always complete
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t) { *bar = 0; }
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t*) { memcpy(smem_dst, gmem_src, bytes); }
__device__ __forceinline__ void bulk_prefetch_l2(const void*, uint32_t) {}
__device__ __forceinline__ void mbar_wait(uint64_t*, uint32_t) { __syncwarp(); }   // the issuing lane has copied by then
__device__ __forceinline__ uint32_t warp_max_u32(uint32_t v) {
    for (int s = 16; s; s >>= 1) { const uint32_t o = __shfl_xor_sync(LZF_FULL_MASK, v, s); v = o > v ? o : v; }
    return v;
}
#endif

}  // namespace lzf
