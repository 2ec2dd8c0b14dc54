// simt_emu.h — TEST INFRASTRUCTURE ONLY.
//
// A tiny CPU SIMT emulator: it lets the *same* kernel sources that nvcc compiles for sm_100a
// (rust-lz-fear_b200/csrc/*.cu) be compiled by g++ and executed on the host, one ucontext fibre
// per CUDA thread, so that the warp-level logic (ballot / shfl / match_any speculation, the
// overlapping-copy arithmetic, the frame layout scans) can be checked against the oracle in the
// `-m "not gpu"` test tier of a box that has no GPU.  It is NOT a fallback: nothing under
// rust-lz-fear_b200/ includes, links or loads it, and the shipped library fails loudly without a
// CUDA device.  Fibres of one CTA are scheduled round-robin and only switch at warp/CTA
// collectives, so data races that need a __syncwarp() are not detected here (compute-sanitizer
// racecheck on the GPU box covers that).
#pragma once

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
#include <execinfo.h>

#include <algorithm>
#include <functional>
#include <vector>

#define LZF_SIMT_EMU 1

// ---- CUDA keywords ---------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct __attribute__((aligned(16))) uint4 { uint32_t x, y, z, w; };
struct __attribute__((aligned(8))) uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

// ---- the sliver of the CUDA runtime the C-ABI layer uses: "device" memory is host memory, streams
// are synchronous.  SIMT_NUM_SMS (env) sets the reported SM count (default 2).
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaDevAttrMultiProcessorCount = 16 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
template <typename F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "simt-emu error"; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) {
    const char* e = getenv("SIMT_NUM_SMS"); *v = e ? atoi(e) : 2; if (*v < 1) *v = 1; return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < h; r++) memmove((uint8_t*)d + r * dp, (const uint8_t*)s + r * sp, w);
    return cudaSuccess;
}

namespace simt {

constexpr int kWarp = 32;
constexpr size_t kStack = 256 * 1024;

struct Fibre {
    ucontext_t ctx;
    uint8_t* stack = nullptr;
    bool done = false;
    uint3 tid{0, 0, 0};
};

struct WarpSlot {           // one in-flight collective per warp
    uint64_t in[kWarp];
    uint64_t out[kWarp];
    uint32_t arg[kWarp];
    int op = -1;
    uint32_t arrived = 0;   // lanes that have deposited an input
    uint32_t generation = 0;
};

struct State {
    std::vector<Fibre> fibres;
    std::vector<WarpSlot> warps;
    ucontext_t sched;
    int cur = -1;
    uint3 bid{0, 0, 0};
    dim3 bdim, gdim;
    uint8_t* dyn_smem = nullptr;
    std::function<void()> body;
    // CTA barrier
    uint32_t bar_arrived = 0, bar_generation = 0;
    uint64_t collectives = 0;
};
extern State g;

inline const uint3& thread_idx() { return g.fibres[g.cur].tid; }
inline void yield() { int me = g.cur; swapcontext(&g.fibres[me].ctx, &g.sched); g.cur = me; }

enum Op { OP_BALLOT, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_MATCH_ANY, OP_ANY, OP_ALL, OP_SYNC, OP_REDUCE_OR };

// All 32 lanes of a warp must call with the full mask (our kernels only ever use full masks).
inline uint64_t collective(Op op, uint64_t value, uint32_t arg) {
    const int me = g.cur;
    const int lane = me % kWarp;
    WarpSlot& w = g.warps[me / kWarp];
    const int base = (me / kWarp) * kWarp;
    int live = 0;
    for (int l = 0; l < kWarp; l++) if (base + l < (int)g.fibres.size() && !g.fibres[base + l].done) live++;
    if (live != kWarp) { fprintf(stderr, "simt: collective with %d live lanes (partial warps unsupported)\n", live); abort(); }
    static const bool trace = getenv("SIMT_TRACE") != nullptr;
    if (trace) { void* bt[4]; backtrace(bt, 4); fprintf(stderr, "T warp %d lane %d op %d val %llu\n", me / kWarp, lane, (int)op, (unsigned long long)value); }
    if (w.arrived == 0) w.op = (int)op;
    else if (w.op != (int)op) {
        fprintf(stderr, "simt: divergent collectives in one warp (%d vs %d) at lane %d\n", w.op, (int)op, lane);
        void* bt[16]; const int nbt = backtrace(bt, 16); backtrace_symbols_fd(bt, nbt, 2);
        abort();
    }
    w.in[lane] = value;
    w.arg[lane] = arg;   // per-lane argument (shfl source lane etc.)
    const uint32_t gen = w.generation;
    w.arrived++;
    if (w.arrived == (uint32_t)kWarp) {
        g.collectives++;
        for (int l = 0; l < kWarp; l++) {
            const uint32_t a = w.arg[l];
            uint64_t r = 0;
            switch (op) {
                case OP_BALLOT: for (int k = 0; k < kWarp; k++) if (w.in[k]) r |= (1ull << k); break;
                case OP_ANY: for (int k = 0; k < kWarp; k++) if (w.in[k]) r = 1; break;
                case OP_ALL: r = 1; for (int k = 0; k < kWarp; k++) if (!w.in[k]) r = 0; break;
                case OP_SHFL: r = w.in[a & 31u]; break;
                case OP_SHFL_UP: r = (l >= (int)a) ? w.in[l - a] : w.in[l]; break;
                case OP_SHFL_DOWN: r = (l + (int)a < kWarp) ? w.in[l + a] : w.in[l]; break;
                case OP_SHFL_XOR: r = w.in[(l ^ a) & 31]; break;
                case OP_MATCH_ANY: for (int k = 0; k < kWarp; k++) if (w.in[k] == w.in[l]) r |= (1ull << k); break;
                case OP_SYNC: break;
                case OP_REDUCE_OR: for (int k = 0; k < kWarp; k++) r |= w.in[k]; break;
            }
            w.out[l] = r;
        }
        w.arrived = 0;
        w.generation++;
    } else {
        while (w.generation == gen) yield();
    }
    return w.out[lane];
}

inline void cta_barrier() {
    const uint32_t gen = g.bar_generation;
    uint32_t live = 0;
    for (auto& f : g.fibres) if (!f.done) live++;
    g.bar_arrived++;
    if (g.bar_arrived == live) { g.bar_arrived = 0; g.bar_generation++; }
    else while (g.bar_generation == gen) yield();
}

void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, std::function<void()> body);

}  // namespace simt

#define threadIdx (simt::thread_idx())
#define blockIdx (simt::g.bid)
#define blockDim (simt::g.bdim)
#define gridDim (simt::g.gdim)

// ---- warp / CTA collectives ------------------------------------------------------------------
static inline void __syncwarp(unsigned = 0xffffffffu) { simt::collective(simt::OP_SYNC, 0, 0); }
static inline void __syncthreads() { simt::cta_barrier(); }
static inline unsigned __ballot_sync(unsigned, int pred) { return (unsigned)simt::collective(simt::OP_BALLOT, pred != 0, 0); }
static inline int __any_sync(unsigned, int pred) { return (int)simt::collective(simt::OP_ANY, pred != 0, 0); }
static inline int __all_sync(unsigned, int pred) { return (int)simt::collective(simt::OP_ALL, pred != 0, 0); }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) { return (unsigned)simt::collective(simt::OP_REDUCE_OR, v, 0); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) {
    static_assert(sizeof(T) <= 8, "shfl width");
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    raw = simt::collective(simt::OP_SHFL, raw, (uint32_t)src);
    T r; memcpy(&r, &raw, sizeof(T)); return r;
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    raw = simt::collective(simt::OP_SHFL_UP, raw, d);
    T r; memcpy(&r, &raw, sizeof(T)); return r;
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) {
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    raw = simt::collective(simt::OP_SHFL_DOWN, raw, d);
    T r; memcpy(&r, &raw, sizeof(T)); return r;
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) {
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    raw = simt::collective(simt::OP_SHFL_XOR, raw, (uint32_t)m);
    T r; memcpy(&r, &raw, sizeof(T)); return r;
}
template <typename T> static inline unsigned __match_any_sync(unsigned, T v) {
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    return (unsigned)simt::collective(simt::OP_MATCH_ANY, raw, 0);
}
static inline unsigned __activemask() { return 0xffffffffu; }

// ---- scalar intrinsics -----------------------------------------------------------------------
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
    s &= 31u; return s ? (uint32_t)((((uint64_t)hi << 32) | lo) >> s) : lo;
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) {
    s &= 31u; return s ? (uint32_t)(((((uint64_t)hi << 32) | lo) << s) >> 32) : hi;
}
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t s = (sel >> (4 * i)) & 0xf;
        uint32_t byte = (uint32_t)(v >> (8 * (s & 7))) & 0xff;
        if (s & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline uint32_t __vcmpeq4(uint32_t a, uint32_t b) {
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) if (((a >> (8 * i)) & 0xff) == ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i);
    return r;
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline uint32_t __brev(uint32_t x) { uint32_t r = 0; for (int i = 0; i < 32; i++) if (x & (1u << i)) r |= 1u << (31 - i); return r; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __nanosleep(unsigned) {}
using std::max;
using std::min;
static inline uint64_t min(uint64_t a, unsigned long long b) { return a < b ? a : (uint64_t)b; }
static inline uint64_t max(uint64_t a, unsigned long long b) { return a > b ? a : (uint64_t)b; }

// ---- launch / dynamic shared memory ------------------------------------------------------------
#define LZF_LAUNCH(kernel, grid, block, smem, stream, ...) \
    simt::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define LZF_DYN_SMEM(name) uint8_t* name = simt::g.dyn_smem
