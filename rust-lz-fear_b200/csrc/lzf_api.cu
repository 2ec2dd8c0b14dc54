// lzf_api.cu — the C ABI of include/lzfear_b200.h: context, batched block calls, single-block
// host conveniences and the frame layer (a restatement of the glue in
// src/framed/compress.rs:159-282 and src/framed/decompress.rs:101-161,197-288 around the
// sm_100a block kernels).  There is no CPU code path in here: every entry point needs a CUDA
// device and fails with LZF_ERR_NO_DEVICE / LZF_ERR_CUDA otherwise.
#include "lzf_common.cuh"
#include "lzf_frame.cuh"
#include "lzf_kernels.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
namespace lzf {

struct Buf {            // grow-only allocation
    void* p = nullptr;
    size_t cap = 0;
};

// A bump arena carved out of one Buf (device) mirrored by one pinned host Buf: descriptors are
// written on the host side and shipped with a single H2D copy.
struct Arena {
    size_t used = 0;
    size_t take(size_t bytes, size_t align = 256) {
        used = (used + align - 1) / align * align;
        const size_t o = used;
        used += bytes;
        return o;
    }
};

}  // namespace lzf
using lzf::Arena;
using lzf::Buf;

// One pipeline slot: a stream with everything a call in flight on it needs.  Host-buffer batches
// rotate over the slots so that the H2D copy of chunk i+1, the kernels of chunk i and the D2H copy
// of chunk i-1 overlap; device-pointer calls use slot 0.
struct lzf_slot {
    cudaStream_t stream = nullptr;      // copies + kernels of this slot
    cudaStream_t side = nullptr;        // content-checksum chain, overlapped with the block kernels
    cudaStream_t copy = nullptr;        // sliced H2D feed of the host-buffer compress pipeline
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_feed0 = nullptr, ev_feed1 = nullptr;
    uint32_t* h_seq = nullptr;          // pinned 1, 2, 3, ...: sources of the progress-word copies
    uint32_t* d_counter = nullptr;      // dynamic work counters (encode, decode)
    // The block kernels of a slot share its work counters and table scratch, so two batched calls of one slot
    // never run side by side: a launch on another stream than the previous one first waits for it (ev_last).
    std::mutex launch_mu;
    cudaEvent_t ev_last = nullptr;
    cudaStream_t last_stream = nullptr;
    bool has_last = false;
    Buf d_tables;                       // per-warp global hash tables (hashlog >= 14)
    Buf d_desc, h_desc;                 // descriptor arenas (device / pinned host)
    Buf d_res, h_res;                   // result arenas (device / pinned host)
    Buf d_comp;                         // compressed-block scratch (frame compress)
    Buf d_io_in, d_io_out;              // staging of host-buffer calls
    Buf d_dict, d_aux;                  // dictionary copy / dependent-block descriptors and window scratch
    Buf d_seg;                          // segmented parse: segment descriptors, results and sequence streams
};
constexpr int kSlots = 4;
constexpr uint32_t kMaxSlices = 256;

struct lzf_ctx {
    int device = 0;
    int num_sms = 0;
    std::atomic<uint64_t> launches{0};
    std::mutex err_mu;
    std::string err;
    lzf_slot slots[kSlots];
    // payload per pipeline chunk of the host-buffer frame calls.  A chunk is one kernel launch and one
    // warp owns one block, so the chunks in flight together must hold enough blocks to fill the GPU
    // (148 SMs x tens of warps): up to kSlots chunks run concurrently, one host thread + stream each.
    // decompress: compressed + plaintext bytes.  Measured on B200 (config 2 e2e, GiB/s): 256 MiB 41.6, 512 MiB 40.3,
    // 1 GiB 39.2 once the host-consumed results no longer queue behind the bulk D2H copies
    uint64_t chunk_bytes = 256ull << 20;
    // compress: plaintext bytes.  0 = one full wave of the block kernel (one warp per block, 28 warps per SM;
    // the parse is latency-bound, so a chunk with fewer blocks takes just as long), at most 24 GiB
    uint64_t compress_chunk_bytes = 0;
    // LZF_OPT_SEGMENT_BYTES: 0 = every block is parsed exactly like the reference, by one warp (default).  > 0: a compress
    // launch that cannot fill the GPU cuts its blocks into segments of at least this many bytes (see SegmentPlanArgs)
    std::atomic<uint64_t> segment_bytes{0};
    // tuning / test knobs.  The environment is read ONCE, in lzf_create (LZF_B200_*); nothing on a call path calls getenv.
    struct Tuning {
        bool trace = false;                     // LZF_B200_TRACE: phase times of the host-buffer compress chunks on stderr
        uint64_t feed_slice = 64 << 10;         // LZF_B200_FEED_SLICE: slice of the H2D feed under the encode kernel (0 = off)
        uint64_t feed_min_blocks = 64;          // LZF_B200_FEED_MIN_BLOCKS
        uint32_t enc_u32_slots = 0;             // LZF_B200_ENC_U32
        uint32_t enc_smem_warps_p1 = 0;         // LZF_B200_ENC_SMEM_WARPS + 1
        uint32_t dec_ctas_per_sm = 0;           // LZF_B200_DEC_CTAS_PER_SM
        uint64_t pos_limit = 0xffffffffull;     // LZF_B200_TEST_POS_LIMIT (tests only): last stream position a u32 slot holds
    } tune;
};

// slot of the calling thread: worker threads of the host-buffer pipelines bind their own, every other
// call works in slot 0
static thread_local lzf_slot* tls_slot = nullptr;
static inline lzf_slot* cur_slot(lzf_ctx* c) { return tls_slot ? tls_slot : &c->slots[0]; }

namespace {

int fail(lzf_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess) {
    if (c) {
        std::lock_guard<std::mutex> lock(c->err_mu);
        c->err = what;
        if (e != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(e); }
    }
    return code;
}
#define LZF_CU(c, call)                                                         \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) return fail((c), LZF_ERR_CUDA, #call, e__);     \
    } while (0)
#define LZF_LAUNCHED(c, rc, n)                                                  \
    do {                                                                        \
        const int rc__ = (rc);                                                  \
        if (rc__ != 0) return fail((c), LZF_ERR_CUDA, "kernel launch", (cudaError_t)rc__); \
        (c)->launches += (n);                                                   \
    } while (0)

int ensure_dev(lzf_ctx* c, Buf& b, size_t bytes) {
    if (bytes <= b.cap) return LZF_SUCCESS;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) { want = bytes; e = cudaMalloc(&b.p, want); }
    if (e != cudaSuccess) { b.p = nullptr; return fail(c, LZF_ERR_OOM, "cudaMalloc", e); }
    b.cap = want;
    return LZF_SUCCESS;
}
int ensure_host(lzf_ctx* c, Buf& b, size_t bytes) {
    if (bytes <= b.cap) return LZF_SUCCESS;
    if (b.p) { cudaFreeHost(b.p); b.p = nullptr; b.cap = 0; }
    const size_t want = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaMallocHost(&b.p, want);
    if (e != cudaSuccess) { b.p = nullptr; return fail(c, LZF_ERR_OOM, "cudaMallocHost", e); }
    b.cap = want;
    return LZF_SUCCESS;
}

inline void wr32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
inline void wr64(uint8_t* p, uint64_t v) { wr32(p, (uint32_t)v); wr32(p + 4, (uint32_t)(v >> 32)); }

}  // namespace

extern "C" int lzf_abi_version(void) { return LZF_ABI_VERSION; }

extern "C" int lzf_create(int device, lzf_ctx** out) {
    if (!out) return LZF_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return LZF_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return LZF_ERR_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return LZF_ERR_CUDA;
    lzf_ctx* c = new (std::nothrow) lzf_ctx();
    if (!c) return LZF_ERR_OOM;
    c->device = device;
    if (cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || c->num_sms <= 0) {
        delete c;
        return LZF_ERR_CUDA;
    }
    bool ok = true;
    for (int i = 0; i < kSlots && ok; i++) {
        lzf_slot& sl = c->slots[i];
        ok = cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking) == cudaSuccess &&
             cudaStreamCreateWithFlags(&sl.side, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&sl.ev_fork, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&sl.ev_join, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&sl.ev_last, cudaEventDisableTiming) == cudaSuccess &&
             cudaStreamCreateWithFlags(&sl.copy, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&sl.ev_feed0, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&sl.ev_feed1, cudaEventDisableTiming) == cudaSuccess &&
             cudaMallocHost((void**)&sl.h_seq, kMaxSlices * sizeof(uint32_t)) == cudaSuccess &&
             cudaMalloc((void**)&sl.d_counter, 256) == cudaSuccess;
    }
    for (int i = 0; i < kSlots && ok; i++)
        for (uint32_t k = 0; k < kMaxSlices; k++) c->slots[i].h_seq[k] = k + 1;
    if (const char* e = getenv("LZF_B200_CHUNK_BYTES")) {       // tuning / test knobs, all read here and only here
        const unsigned long long v = strtoull(e, nullptr, 10);
        if (v) c->chunk_bytes = c->compress_chunk_bytes = v;
    }
    c->tune.trace = getenv("LZF_B200_TRACE") != nullptr;
    if (const char* e = getenv("LZF_B200_FEED_SLICE")) c->tune.feed_slice = strtoull(e, nullptr, 10);
    if (const char* e = getenv("LZF_B200_FEED_MIN_BLOCKS")) c->tune.feed_min_blocks = strtoull(e, nullptr, 10);
    c->tune.enc_u32_slots = getenv("LZF_B200_ENC_U32") != nullptr;
    if (const char* e = getenv("LZF_B200_ENC_SMEM_WARPS")) { const int v = atoi(e); if (v >= 0) c->tune.enc_smem_warps_p1 = (uint32_t)v + 1; }
    if (const char* e = getenv("LZF_B200_DEC_CTAS_PER_SM")) { const int v = atoi(e); if (v >= 1) c->tune.dec_ctas_per_sm = (uint32_t)v; }
    if (const char* e = getenv("LZF_B200_TEST_POS_LIMIT")) { const unsigned long long v = strtoull(e, nullptr, 10); if (v && v < 0xffffffffull) c->tune.pos_limit = v; }
    if (!ok) { lzf_destroy(c); return LZF_ERR_CUDA; }
    *out = c;
    return LZF_SUCCESS;
}

extern "C" void lzf_destroy(lzf_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (int i = 0; i < kSlots; i++) {
        lzf_slot& sl = c->slots[i];
        if (sl.stream) cudaStreamSynchronize(sl.stream);
        if (sl.side) cudaStreamSynchronize(sl.side);
        if (sl.copy) { cudaStreamSynchronize(sl.copy); cudaStreamDestroy(sl.copy); }
        if (sl.ev_feed0) cudaEventDestroy(sl.ev_feed0);
        if (sl.ev_feed1) cudaEventDestroy(sl.ev_feed1);
        if (sl.h_seq) cudaFreeHost(sl.h_seq);
        Buf* dev[] = {&sl.d_tables, &sl.d_desc, &sl.d_res, &sl.d_comp, &sl.d_io_in, &sl.d_io_out, &sl.d_dict, &sl.d_aux, &sl.d_seg};
        for (Buf* b : dev) if (b->p) cudaFree(b->p);
        Buf* host[] = {&sl.h_desc, &sl.h_res};
        for (Buf* b : host) if (b->p) cudaFreeHost(b->p);
        if (sl.d_counter) cudaFree(sl.d_counter);
        if (sl.ev_fork) cudaEventDestroy(sl.ev_fork);
        if (sl.ev_join) cudaEventDestroy(sl.ev_join);
        if (sl.ev_last) cudaEventDestroy(sl.ev_last);
        if (sl.stream) cudaStreamDestroy(sl.stream);
        if (sl.side) cudaStreamDestroy(sl.side);
    }
    delete c;
}

// Gives the grow-only scratch of every pipeline slot back to the driver (device staging, descriptor arenas, segment
// streams, table scratch); the next call allocates what it needs again.  No call may be in flight on the ctx.
extern "C" int lzf_trim(lzf_ctx* c) {
    if (!c) return LZF_ERR_INVALID_ARG;
    LZF_CU(c, cudaSetDevice(c->device));
    for (int i = 0; i < kSlots; i++) {
        lzf_slot& sl = c->slots[i];
        std::lock_guard<std::mutex> lock(sl.launch_mu);
        if (sl.stream) cudaStreamSynchronize(sl.stream);
        if (sl.side) cudaStreamSynchronize(sl.side);
        if (sl.copy) cudaStreamSynchronize(sl.copy);
        Buf* dev[] = {&sl.d_tables, &sl.d_desc, &sl.d_res, &sl.d_comp, &sl.d_io_in, &sl.d_io_out, &sl.d_dict, &sl.d_aux, &sl.d_seg};
        for (Buf* b : dev) if (b->p) { cudaFree(b->p); b->p = nullptr; b->cap = 0; }
        Buf* host[] = {&sl.h_desc, &sl.h_res};
        for (Buf* b : host) if (b->p) { cudaFreeHost(b->p); b->p = nullptr; b->cap = 0; }
        sl.has_last = false;
    }
    return LZF_SUCCESS;
}

extern "C" int lzf_set_option(lzf_ctx* c, int option, uint64_t value) {
    if (!c) return LZF_ERR_INVALID_ARG;
    switch (option) {
        case LZF_OPT_SEGMENT_BYTES:
            if (value && value < 65536) return fail(c, LZF_ERR_INVALID_ARG, "segments are at least 64 KiB");
            c->segment_bytes.store(value);
            return LZF_SUCCESS;
        default: return fail(c, LZF_ERR_INVALID_ARG, "unknown option");
    }
}

extern "C" const char* lzf_last_error(const lzf_ctx* c) { return c ? c->err.c_str() : "null ctx"; }
extern "C" uint64_t lzf_launch_count(const lzf_ctx* c) { return c ? c->launches.load() : 0; }
extern "C" size_t lzf_compress_bound(size_t n) { return n + n / 255 + 16; }

// ------------------------------------------------------------------------------------------------
// batched block calls (device pointers, asynchronous on `stream`)
// ------------------------------------------------------------------------------------------------
namespace {

// Orders a block-kernel launch on stream `s` behind the previous one of the same slot (see lzf_slot::ev_last).
// Held for the whole memset + launch + record sequence: host threads sharing a ctx serialise here.
struct LaunchOrder {
    lzf_slot* sl; cudaStream_t s; std::unique_lock<std::mutex> lock;
    LaunchOrder(lzf_slot* slot, cudaStream_t stream) : sl(slot), s(stream), lock(slot->launch_mu) {}
    cudaError_t begin() {
        if (sl->has_last && sl->last_stream != s) return cudaStreamWaitEvent(s, sl->ev_last, 0);
        return cudaSuccess;
    }
    cudaError_t end() {
        sl->has_last = true; sl->last_stream = s;
        return cudaEventRecord(sl->ev_last, s);
    }
};

// Host-buffer pipeline: the plaintext is still being copied in (slice k of every block, then slice k + 1, ...) on
// another stream while the block kernel runs; see EncodeArgs::progress.
struct InputFeed {
    const uint32_t* d_progress; uint32_t slice_bytes;
    cudaEvent_t ready;      // recorded behind the last slice
    // queues the slice copies.  Called right before the block kernel is launched, after every other host->device
    // copy and memset of the call has been queued: the copy engine works first-in first-out across streams, so
    // anything queued behind 16 GiB of slices would hold the kernel launch back until the feed is over
    std::function<int()> start;
};

// Segments per block for a launch of `nblocks` independent blocks of at most max_block_len bytes (1 = no segmentation):
// as many as it takes to offer the block kernel one warp per resident slot, segments no shorter than the option asks.
uint32_t plan_segments(const lzf_ctx* c, uint32_t nblocks, uint32_t max_block_len, uint32_t table_kind) {
    const uint64_t seg_min = c->segment_bytes.load();
    if (!seg_min || !nblocks || table_kind != LZF_TABLE_U32 || max_block_len < 2 * seg_min || max_block_len > (16u << 20)) return 1;
    const uint64_t wave = (uint64_t)c->num_sms * 28;              // kEncodeBigWarps resident warps per SM
    if ((uint64_t)nblocks * 2 > wave) return 1;
    uint64_t S = (wave + nblocks - 1) / nblocks;
    if (S > max_block_len / seg_min) S = max_block_len / seg_min;
    if (S > 64) S = 64;
    return S < 2 ? 1u : (uint32_t)S;
}

int compress_blocks_impl(lzf_ctx* c, const uint8_t* d_in, const uint64_t* d_in_off, const uint32_t* d_in_len,
                         uint32_t nblocks, uint32_t hashlog, uint32_t table_kind, uint32_t max_block_len,
                         uint8_t* d_out, const uint64_t* d_out_off, const uint32_t* d_out_cap,
                         uint32_t* d_out_len, int32_t* d_status, uint32_t* d_xxh_plain, uint32_t* d_xxh_stored,
                         cudaStream_t s, const lzf::EncodeArgs* chains = nullptr, const InputFeed* feed = nullptr);

// The segmented parse of SegmentPlanArgs: plan -> one encode launch over all segments -> stitch (-> block checksums).
int compress_blocks_segmented(lzf_ctx* c, const uint8_t* d_in, const uint64_t* d_in_off, const uint32_t* d_in_len,
                              uint32_t nblocks, uint32_t hashlog, uint32_t max_block_len, uint32_t S,
                              uint8_t* d_out, const uint64_t* d_out_off, const uint32_t* d_out_cap,
                              uint32_t* d_out_len, int32_t* d_status, uint32_t* d_xxh_plain, uint32_t* d_xxh_stored,
                              cudaStream_t s) {
    const uint32_t seg_len = (uint32_t)((((uint64_t)max_block_len + S - 1) / S + 15) / 16 * 16);
    const uint32_t seg_cap = (uint32_t)((lzf_compress_bound(seg_len) + 15) / 16 * 16);
    const size_t N = (size_t)nblocks * S;
    Arena ar;
    const size_t o_in_off = ar.take(N * 8), o_out_off = ar.take(N * 8);
    const size_t o_in_len = ar.take(N * 4), o_pfx = ar.take(N * 4), o_cap = ar.take(N * 4), o_cf = ar.take(N * 4), o_cc = ar.take(N * 4);
    const size_t o_abs = ar.take(N * 4), o_olen = ar.take(N * 4), o_st = ar.take(N * 4), o_fp = ar.take(N * 4), o_fl = ar.take(N * 4);
    const size_t o_jobs = ar.take((size_t)nblocks * (S + 1) * sizeof(lzf::StitchJob));
    const size_t o_h = ar.take((size_t)nblocks * 8 * 4);
    const size_t o_streams = ar.take(N * seg_cap + 64);
    lzf_slot* sl = cur_slot(c);
    int rc;
    if ((rc = ensure_dev(c, sl->d_seg, ar.used))) return rc;
    uint8_t* g = (uint8_t*)sl->d_seg.p;
    lzf::SegmentPlanArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.nblocks = nblocks; pa.nseg = S; pa.seg_len = seg_len; pa.seg_cap = seg_cap;
    pa.in_off = d_in_off; pa.in_len = d_in_len;
    pa.seg_in_off = (uint64_t*)(g + o_in_off); pa.seg_in_len = (uint32_t*)(g + o_in_len); pa.seg_prefix = (uint32_t*)(g + o_pfx);
    pa.seg_out_off = (uint64_t*)(g + o_out_off); pa.seg_out_cap = (uint32_t*)(g + o_cap);
    pa.seg_chain_first = (uint32_t*)(g + o_cf); pa.seg_chain_count = (uint32_t*)(g + o_cc); pa.seg_abs = (uint32_t*)(g + o_abs);
    LZF_LAUNCHED(c, lzf_launch_segment_plan(&pa, s), 1);
    lzf::EncodeArgs ch;
    memset(&ch, 0, sizeof(ch));
    ch.prefix_len = pa.seg_prefix; ch.abs_base = pa.seg_abs; ch.prime_len = pa.seg_prefix;
    ch.chain_first = pa.seg_chain_first; ch.chain_count = pa.seg_chain_count; ch.nchains = (uint32_t)N;
    ch.max_pos = (uint64_t)LZF_WINDOW_SIZE + seg_len;
    ch.fin_pos = (uint32_t*)(g + o_fp); ch.fin_lit = (uint32_t*)(g + o_fl);
    rc = compress_blocks_impl(c, d_in, pa.seg_in_off, pa.seg_in_len, (uint32_t)N, hashlog, LZF_TABLE_U32, seg_len, g + o_streams,
                              pa.seg_out_off, pa.seg_out_cap, (uint32_t*)(g + o_olen), (int32_t*)(g + o_st), nullptr, nullptr, s, &ch);
    if (rc) return rc;
    lzf::StitchArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.nblocks = nblocks; sa.nseg = S; sa.seg_len = seg_len;
    sa.in = d_in; sa.in_off = d_in_off; sa.in_len = d_in_len;
    sa.seg = g + o_streams; sa.seg_out_off = pa.seg_out_off; sa.seg_out_len = (const uint32_t*)(g + o_olen);
    sa.seg_status = (const int32_t*)(g + o_st); sa.fin_pos = ch.fin_pos; sa.fin_lit = ch.fin_lit;
    sa.out = d_out; sa.out_off = d_out_off; sa.out_cap = d_out_cap; sa.out_len = d_out_len; sa.status = d_status;
    sa.jobs = (lzf::StitchJob*)(g + o_jobs);
    uint64_t* hh = (uint64_t*)(g + o_h);
    if (d_xxh_stored) { sa.hash_off = hh; sa.hash_len = hh + nblocks; }
    if (d_xxh_plain) { sa.plain_off = hh + 2 * (size_t)nblocks; sa.plain_len = hh + 3 * (size_t)nblocks; }
    LZF_LAUNCHED(c, lzf_launch_stitch(&sa, s), 2);
    if (d_xxh_stored) LZF_LAUNCHED(c, lzf_launch_xxh32_ranges(nullptr, sa.hash_off, sa.hash_len, nblocks, d_xxh_stored, s), 1);
    if (d_xxh_plain) LZF_LAUNCHED(c, lzf_launch_xxh32_ranges(nullptr, sa.plain_off, sa.plain_len, nblocks, d_xxh_plain, s), 1);
    return LZF_SUCCESS;
}

int compress_blocks_impl(lzf_ctx* c, const uint8_t* d_in, const uint64_t* d_in_off, const uint32_t* d_in_len,
                         uint32_t nblocks, uint32_t hashlog, uint32_t table_kind, uint32_t max_block_len,
                         uint8_t* d_out, const uint64_t* d_out_off, const uint32_t* d_out_cap,
                         uint32_t* d_out_len, int32_t* d_status, uint32_t* d_xxh_plain, uint32_t* d_xxh_stored,
                         cudaStream_t s, const lzf::EncodeArgs* chains, const InputFeed* feed) {
    if (hashlog == 0) hashlog = 12;
    if (hashlog < 8 || hashlog > 16) return fail(c, LZF_ERR_INVALID_ARG, "hashlog must be 0 or 8..16");
    if (table_kind != LZF_TABLE_U32 && table_kind != LZF_TABLE_U16) return fail(c, LZF_ERR_INVALID_ARG, "table_kind");
    if (nblocks == 0) return LZF_SUCCESS;
    if ((!d_in && !chains) || !d_in_off || !d_in_len || !d_out || !d_out_off || !d_out_len || !d_status)
        return fail(c, LZF_ERR_INVALID_ARG, "null pointer");
    if (!chains && !feed) {
        const uint32_t S = plan_segments(c, nblocks, max_block_len, table_kind);
        if (S > 1) return compress_blocks_segmented(c, d_in, d_in_off, d_in_len, nblocks, hashlog, max_block_len, S, d_out, d_out_off,
                                                    d_out_cap, d_out_len, d_status, d_xxh_plain, d_xxh_stored, s);
    }
    lzf::EncodeArgs a;
    memset(&a, 0, sizeof(a));
    if (chains) {
        a.prefix_len = chains->prefix_len; a.abs_base = chains->abs_base; a.prime_len = chains->prime_len;
        a.chain_first = chains->chain_first; a.chain_count = chains->chain_count; a.nchains = chains->nchains;
        a.max_pos = chains->max_pos; a.table_io = chains->table_io;
        a.fin_pos = chains->fin_pos; a.fin_lit = chains->fin_lit;
        a.allow_slot_wrap = chains->allow_slot_wrap;
    }
    a.in = d_in; a.in_off = d_in_off; a.in_len = d_in_len; a.nblocks = nblocks;
    a.hashlog = hashlog; a.table_kind = table_kind;
    a.out = d_out; a.out_off = d_out_off; a.out_cap = d_out_cap; a.out_len = d_out_len; a.status = d_status;
    a.xxh_plain = d_xxh_plain; a.xxh_stored = d_xxh_stored;
    a.work_counter = cur_slot(c)->d_counter;
    a.max_block_len = max_block_len;
    a.tune_u32_slots = c->tune.enc_u32_slots; a.tune_smem_warps_p1 = c->tune.enc_smem_warps_p1;
    if (const size_t scratch = lzf_encode_global_table_bytes(&a, c->num_sms)) {   // tables that do not fit shared memory
        const int rc = ensure_dev(c, cur_slot(c)->d_tables, scratch);
        if (rc) return rc;
        a.global_tables = (uint8_t*)cur_slot(c)->d_tables.p;
    }
    LaunchOrder order(cur_slot(c), s);
    LZF_CU(c, order.begin());
    LZF_CU(c, cudaMemsetAsync(cur_slot(c)->d_counter, 0, 4, s));
    if (feed) {
        const int rc = feed->start();
        if (rc) return rc;
        a.progress = feed->d_progress; a.slice_bytes = feed->slice_bytes;
    }
    LZF_LAUNCHED(c, lzf_launch_encode(&a, c->num_sms, s), 1);
    LZF_CU(c, order.end());
    return LZF_SUCCESS;
}

int decompress_blocks_impl(lzf_ctx* c, const uint8_t* d_in, const uint64_t* d_in_off, const uint32_t* d_in_len,
                           uint32_t nblocks, const uint8_t* d_prefix, const uint64_t* d_prefix_off,
                           const uint32_t* d_prefix_len, uint8_t* d_out, const uint64_t* d_out_off,
                           const uint32_t* d_out_cap, const uint32_t* d_out_limit, uint32_t* d_out_len,
                           int32_t* d_status, uint32_t* d_xxh_plain, cudaStream_t s,
                           bool prefix_abs = false, const int32_t* d_wait_for = nullptr, uint32_t* d_done = nullptr) {
    if (nblocks == 0) return LZF_SUCCESS;
    if (!d_in || !d_in_off || !d_in_len || !d_out || !d_out_off || !d_out_cap || !d_out_limit || !d_out_len || !d_status)
        return fail(c, LZF_ERR_INVALID_ARG, "null pointer");
    if ((d_prefix || prefix_abs) && (!d_prefix_off || !d_prefix_len)) return fail(c, LZF_ERR_INVALID_ARG, "prefix triple incomplete");
    lzf::DecodeArgs a;
    memset(&a, 0, sizeof(a));
    a.in = d_in; a.in_off = d_in_off; a.in_len = d_in_len; a.nblocks = nblocks;
    a.prefix = d_prefix; a.prefix_off = d_prefix_off; a.prefix_len = d_prefix_len;
    a.out = d_out; a.out_off = d_out_off; a.out_cap = d_out_cap; a.out_limit = d_out_limit;
    a.out_len = d_out_len; a.status = d_status; a.xxh_plain = d_xxh_plain;
    a.prefix_abs = prefix_abs ? 1 : 0; a.wait_for = d_wait_for; a.done = d_done;
    a.work_counter = cur_slot(c)->d_counter + 16;
    a.tune_ctas_per_sm = c->tune.dec_ctas_per_sm;
    LaunchOrder order(cur_slot(c), s);
    LZF_CU(c, order.begin());
    LZF_CU(c, cudaMemsetAsync(cur_slot(c)->d_counter + 16, 0, 4, s));
    LZF_LAUNCHED(c, lzf_launch_decode(&a, c->num_sms, s), 1);
    LZF_CU(c, order.end());
    return LZF_SUCCESS;
}

}  // namespace

extern "C" int lzf_compress_blocks(lzf_ctx* c, const uint8_t* d_in, const uint64_t* d_in_off, const uint32_t* d_in_len,
                                   uint32_t nblocks, uint32_t hashlog, uint32_t table_kind, uint32_t max_block_len,
                                   uint8_t* d_out, const uint64_t* d_out_off, const uint32_t* d_out_cap,
                                   uint32_t* d_out_len, int32_t* d_status,
                                   uint32_t* d_xxh_plain, uint32_t* d_xxh_stored, void* stream) {
    if (!c) return LZF_ERR_INVALID_ARG;
    LZF_CU(c, cudaSetDevice(c->device));
    return compress_blocks_impl(c, d_in, d_in_off, d_in_len, nblocks, hashlog, table_kind, max_block_len, d_out, d_out_off,
                                d_out_cap, d_out_len, d_status, d_xxh_plain, d_xxh_stored, (cudaStream_t)stream);
}

extern "C" int lzf_decompress_blocks(lzf_ctx* c, const uint8_t* d_in, const uint64_t* d_in_off, const uint32_t* d_in_len,
                                     uint32_t nblocks, const uint8_t* d_prefix, const uint64_t* d_prefix_off,
                                     const uint32_t* d_prefix_len, uint8_t* d_out, const uint64_t* d_out_off,
                                     const uint32_t* d_out_cap, const uint32_t* d_out_limit, uint32_t* d_out_len,
                                     int32_t* d_status, uint32_t* d_xxh_plain, void* stream) {
    if (!c) return LZF_ERR_INVALID_ARG;
    LZF_CU(c, cudaSetDevice(c->device));
    return decompress_blocks_impl(c, d_in, d_in_off, d_in_len, nblocks, d_prefix, d_prefix_off, d_prefix_len, d_out,
                                  d_out_off, d_out_cap, d_out_limit, d_out_len, d_status, d_xxh_plain,
                                  (cudaStream_t)stream);
}

extern "C" int lzf_xxh32_ranges(lzf_ctx* c, const uint8_t* d_data, const uint64_t* d_off, const uint64_t* d_len,
                                uint32_t nranges, uint32_t* d_hash, void* stream) {
    if (!c) return LZF_ERR_INVALID_ARG;
    if (nranges == 0) return LZF_SUCCESS;
    if (!d_off || !d_len || !d_hash) return fail(c, LZF_ERR_INVALID_ARG, "null pointer");
    LZF_CU(c, cudaSetDevice(c->device));
    LZF_LAUNCHED(c, lzf_launch_xxh32_ranges(d_data, d_off, d_len, nranges, d_hash, (cudaStream_t)stream), 1);
    return LZF_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
// streaming XXH32 over host buffers (stripes on the GPU)
// ------------------------------------------------------------------------------------------------
extern "C" void lzf_xxh32_init(lzf_xxh32_state* st) {
    if (!st) return;
    memset(st, 0, sizeof(*st));
    st->acc[0] = lzf::XP1 + lzf::XP2; st->acc[1] = lzf::XP2; st->acc[2] = 0; st->acc[3] = 0u - lzf::XP1;
}

extern "C" int lzf_xxh32_update(lzf_ctx* c, lzf_xxh32_state* st, const uint8_t* data, size_t n) {
    if (!c || !st || (n && !data)) return LZF_ERR_INVALID_ARG;
    if (n == 0) return LZF_SUCCESS;
    LZF_CU(c, cudaSetDevice(c->device));
    st->total += n;
    // bytes that complete the carried partial stripe, then whole stripes, then the new carry
    const size_t head = st->buflen ? ((16 - st->buflen) < n ? (16 - st->buflen) : n) : 0;
    memcpy(st->buf + st->buflen, data, head);
    const bool head_full = st->buflen + head == 16;
    if (st->buflen && !head_full) { st->buflen += (uint32_t)head; return LZF_SUCCESS; }
    const size_t body = (n - head) & ~(size_t)15;
    const size_t tail = n - head - body;
    const size_t dev_bytes = (head_full ? 16 : 0) + body;
    if (dev_bytes) {
        int rc;
        if ((rc = ensure_dev(c, cur_slot(c)->d_io_in, dev_bytes + 64))) return rc;
        if ((rc = ensure_dev(c, cur_slot(c)->d_desc, 4096))) return rc;
        if ((rc = ensure_host(c, cur_slot(c)->h_desc, 4096))) return rc;
        uint8_t* din = (uint8_t*)cur_slot(c)->d_io_in.p;
        cudaStream_t s = cur_slot(c)->stream;
        memcpy(cur_slot(c)->h_desc.p, st->acc, 16);
        if (head_full) memcpy((uint8_t*)cur_slot(c)->h_desc.p + 16, st->buf, 16);
        LZF_CU(c, cudaMemcpyAsync(cur_slot(c)->d_desc.p, cur_slot(c)->h_desc.p, 32, cudaMemcpyHostToDevice, s));
        if (head_full) LZF_CU(c, cudaMemcpyAsync(din, (uint8_t*)cur_slot(c)->d_desc.p + 16, 16, cudaMemcpyDeviceToDevice, s));
        if (body) LZF_CU(c, cudaMemcpyAsync(din + (head_full ? 16 : 0), data + head, body, cudaMemcpyHostToDevice, s));
        LZF_LAUNCHED(c, lzf_launch_xxh32_stripes(din, dev_bytes / 16, (uint32_t*)cur_slot(c)->d_desc.p, s), 1);
        LZF_CU(c, cudaMemcpyAsync(cur_slot(c)->h_desc.p, cur_slot(c)->d_desc.p, 16, cudaMemcpyDeviceToHost, s));
        LZF_CU(c, cudaStreamSynchronize(s));
        memcpy(st->acc, cur_slot(c)->h_desc.p, 16);
    }
    if (head_full) st->buflen = 0;
    memcpy(st->buf, data + head + body, tail);
    st->buflen = (uint32_t)tail;
    return LZF_SUCCESS;
}

extern "C" uint32_t lzf_xxh32_finish(const lzf_xxh32_state* st) {
    if (!st) return 0;
    auto rotl = [](uint32_t x, int r) { return (x << r) | (x >> (32 - r)); };
    uint32_t h = st->total >= 16 ? rotl(st->acc[0], 1) + rotl(st->acc[1], 7) + rotl(st->acc[2], 12) + rotl(st->acc[3], 18)
                                 : lzf::XP5;
    h += (uint32_t)st->total;
    const uint8_t* p = st->buf;
    uint32_t n = st->buflen;
    while (n >= 4) {
        const uint32_t x = uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24);
        h = rotl(h + x * lzf::XP3, 17) * lzf::XP4; p += 4; n -= 4;
    }
    while (n) { h = rotl(h + uint32_t(*p) * lzf::XP5, 11) * lzf::XP1; p++; n--; }
    h ^= h >> 15; h *= lzf::XP2;
    h ^= h >> 13; h *= lzf::XP3;
    h ^= h >> 16;
    return h;
}

// ------------------------------------------------------------------------------------------------
// single-block host-pointer conveniences (raw::compress2 / raw::decompress_raw shape)
// ------------------------------------------------------------------------------------------------
extern "C" int lzf_raw_compress_into(lzf_ctx* c, const uint8_t* in, size_t n, uint32_t table_kind, uint32_t hashlog,
                                     uint8_t* out, size_t cap, size_t* written, int32_t* status) {
    if (!c || !written || !status || (n && !in) || (cap && !out)) return LZF_ERR_INVALID_ARG;
    *written = 0;
    *status = LZF_OK;
    // assert!(input.len() <= T::payload_size_limit())   src/raw/compress/mod.rs:167
    if (n > 0xffffffffull || (table_kind == LZF_TABLE_U16 && n > 0xffffull)) { *status = LZF_PANIC; return LZF_SUCCESS; }
    LZF_CU(c, cudaSetDevice(c->device));
    const size_t capc = cap > 0xffffffffull ? 0xffffffffull : cap;
    int rc;
    if ((rc = ensure_dev(c, cur_slot(c)->d_io_in, n + 64))) return rc;
    if ((rc = ensure_dev(c, cur_slot(c)->d_io_out, capc + 64))) return rc;
    if ((rc = ensure_dev(c, cur_slot(c)->d_desc, 4096))) return rc;
    if ((rc = ensure_host(c, cur_slot(c)->h_desc, 4096))) return rc;
    uint8_t* h = (uint8_t*)cur_slot(c)->h_desc.p;
    uint8_t* d = (uint8_t*)cur_slot(c)->d_desc.p;
    // layout: in_off u64 @0, out_off u64 @8, in_len u32 @16, out_cap u32 @20 | results: out_len u32 @64, status i32 @68
    memset(h, 0, 128);
    *(uint32_t*)(h + 16) = (uint32_t)n;
    *(uint32_t*)(h + 20) = (uint32_t)capc;
    cudaStream_t s = cur_slot(c)->stream;
    LZF_CU(c, cudaMemcpyAsync(d, h, 128, cudaMemcpyHostToDevice, s));
    if (n) LZF_CU(c, cudaMemcpyAsync(cur_slot(c)->d_io_in.p, in, n, cudaMemcpyHostToDevice, s));
    rc = compress_blocks_impl(c, (const uint8_t*)cur_slot(c)->d_io_in.p, (const uint64_t*)d, (const uint32_t*)(d + 16), 1, hashlog,
                              table_kind, (uint32_t)n, (uint8_t*)cur_slot(c)->d_io_out.p, (const uint64_t*)(d + 8),
                              (const uint32_t*)(d + 20), (uint32_t*)(d + 64), (int32_t*)(d + 68), nullptr, nullptr, s);
    if (rc) return rc;
    LZF_CU(c, cudaMemcpyAsync(h + 64, d + 64, 8, cudaMemcpyDeviceToHost, s));
    LZF_CU(c, cudaStreamSynchronize(s));
    const uint32_t olen = *(uint32_t*)(h + 64);
    *status = *(int32_t*)(h + 68);
    if (*status == LZF_OK && olen) {
        LZF_CU(c, cudaMemcpyAsync(out, cur_slot(c)->d_io_out.p, olen, cudaMemcpyDeviceToHost, s));
        LZF_CU(c, cudaStreamSynchronize(s));
    }
    *written = *status == LZF_OK ? olen : 0;
    return LZF_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
// compress2 with history and a carried table (src/raw/compress/mod.rs:165-170): the EncoderTable lives on the device
// ------------------------------------------------------------------------------------------------
struct lzf_table {
    uint32_t kind = LZF_TABLE_U32, hashlog = 12, nslots = 4096;
    uint64_t offset = 0;                // U32Table::offset / U16Table::offset (:30,81)
    uint32_t* d_slots = nullptr;        // dict, one u32 per slot (:29,80)
};

extern "C" int lzf_table_create(lzf_ctx* c, uint32_t table_kind, uint32_t hashlog, lzf_table** out) {
    if (!c || !out) return LZF_ERR_INVALID_ARG;
    *out = nullptr;
    if (hashlog == 0) hashlog = 12;
    if (hashlog < 8 || hashlog > 16) return fail(c, LZF_ERR_INVALID_ARG, "hashlog must be 0 or 8..16");
    if (table_kind != LZF_TABLE_U32 && table_kind != LZF_TABLE_U16) return fail(c, LZF_ERR_INVALID_ARG, "table_kind");
    LZF_CU(c, cudaSetDevice(c->device));
    lzf_table* t = new (std::nothrow) lzf_table();
    if (!t) return LZF_ERR_OOM;
    t->kind = table_kind; t->hashlog = hashlog;
    t->nslots = table_kind == LZF_TABLE_U16 ? (2u << hashlog) : (1u << hashlog);      // :28,79
    if (cudaMalloc((void**)&t->d_slots, (size_t)t->nslots * 4) != cudaSuccess) { delete t; return fail(c, LZF_ERR_OOM, "cudaMalloc"); }
    if (cudaMemsetAsync(t->d_slots, 0, (size_t)t->nslots * 4, cur_slot(c)->stream) != cudaSuccess ||
        cudaStreamSynchronize(cur_slot(c)->stream) != cudaSuccess) { cudaFree(t->d_slots); delete t; return fail(c, LZF_ERR_CUDA, "cudaMemset"); }
    *out = t;
    return LZF_SUCCESS;
}

extern "C" void lzf_table_destroy(lzf_ctx* c, lzf_table* t) {
    if (!t) return;
    if (c) cudaSetDevice(c->device);
    if (t->d_slots) cudaFree(t->d_slots);
    delete t;
}

extern "C" int lzf_table_reset(lzf_ctx* c, lzf_table* t) {        // *table = T::default()  :32-36,83-87
    if (!c || !t) return LZF_ERR_INVALID_ARG;
    LZF_CU(c, cudaSetDevice(c->device));
    LZF_CU(c, cudaMemsetAsync(t->d_slots, 0, (size_t)t->nslots * 4, cur_slot(c)->stream));
    LZF_CU(c, cudaStreamSynchronize(cur_slot(c)->stream));
    t->offset = 0;
    return LZF_SUCCESS;
}

extern "C" int lzf_table_offset(lzf_ctx* c, lzf_table* t, uint64_t by) {      // EncoderTable::offset  :72-74,97-99
    if (!c || !t) return LZF_ERR_INVALID_ARG;
    t->offset += by;
    return LZF_SUCCESS;
}

namespace {
// What compress2 returns for a call whose table comes within reach of its slot limit, from the sequence stream `seq` the
// same parse writes into an unbounded writer.  The reference runs front to back: sequence j's positions are inserted
// (its probes, then cursor - 2 behind its match, :196,218) BEFORE its bytes are written (:235-236), so it panics in the
// first sequence that inserts a position p with offset + p > slot_limit unless the writer refused an earlier sequence.
// The largest position sequence j inserts is c_j - 2 (c_j = end of its match; its probes lie at least 4 bytes in front);
// the closing literal run probes literal_start, +step, ... while 12 bytes remain (:177-178,225-231).
int32_t settle_panic_zone(const uint8_t* seq, size_t len, size_t n, size_t cursor, uint64_t offset, uint64_t slot_limit, size_t cap) {
    auto violates = [&](uint64_t pos) { return offset + pos > slot_limit; };
    size_t i = 0, pos = cursor;
    while (i < len) {
        const size_t seq_start = i;
        const uint8_t token = seq[i++];
        size_t lit = token >> 4;
        if (lit == 15) { uint8_t b; do { b = seq[i++]; lit += b; } while (b == 255 && i < len); }
        i += lit;
        if (i >= len) {
            // the closing run: literals only (:178-190)
            size_t cur = pos, step_counter = (size_t)1 << 6, step = 1;
            bool panics = false;
            while (cur < n && n - cur >= 12) {
                if (violates(cur)) { panics = true; break; }
                cur += step;
                step = step_counter >> 6;
                if (pos + 1 != cur) step_counter += 1;
            }
            if (seq_start > cap) return LZF_WRITER_FULL;          // an earlier sequence did not fit
            if (panics) return LZF_PANIC;
            return len > cap ? LZF_WRITER_FULL : LZF_OK;
        }
        i += 2;
        size_t ml = (token & 15);
        if (ml == 15) { uint8_t b; do { b = seq[i++]; ml += b; } while (b == 255 && i < len); }
        ml += 4;
        const size_t match_end = pos + lit + ml;
        if (violates(match_end - 2)) return seq_start > cap ? LZF_WRITER_FULL : LZF_PANIC;
        pos = match_end;
    }
    return len > cap ? LZF_WRITER_FULL : LZF_OK;
}
}  // namespace

extern "C" int lzf_raw_compress2(lzf_ctx* c, const uint8_t* in, size_t n, size_t cursor, lzf_table* t,
                                 uint8_t* out, size_t cap, size_t* written, int32_t* status) {
    if (!c || !t || !written || !status || (n && !in) || (cap && !out) || cursor > n) return LZF_ERR_INVALID_ARG;
    *written = 0;
    *status = LZF_OK;
    // assert!(input.len() <= T::payload_size_limit())   :167
    if (n > 0xffffffffull || (t->kind == LZF_TABLE_U16 && n > 0xffffull)) { *status = LZF_PANIC; return LZF_SUCCESS; }
    if (cursor == n) return LZF_SUCCESS;                               // while cursor < input.len() :171 never runs
    // "EncoderTable contract violated" (:67,92) fires when a position that is INSERTED leaves the slot width.  compress2
    // inserts probe positions <= n - 12 (:178,196) and cursor - 2 <= n - 7 behind a match (:218): a call whose positions
    // up to n - 7 all fit can not panic.  Beyond that it depends on where the parse actually inserts, and on whether the
    // writer ran full first: that rare zone (a table within a few bytes of its 64 KiB / 4 GiB limit) is settled below
    // from the sequence stream of an unbounded run.
    const uint64_t slot_limit = t->kind == LZF_TABLE_U16 ? 0xffffull : 0xffffffffull;
    const bool panic_zone = t->offset + n > slot_limit + 7;
    LZF_CU(c, cudaSetDevice(c->device));
    const size_t want_cap = panic_zone ? lzf_compress_bound(n - cursor) : cap;
    const size_t capc = want_cap > 0xffffffffull ? 0xffffffffull : want_cap;
    int rc;
    if ((rc = ensure_dev(c, cur_slot(c)->d_io_in, n + 64))) return rc;
    if ((rc = ensure_dev(c, cur_slot(c)->d_io_out, capc + 64))) return rc;
    if ((rc = ensure_dev(c, cur_slot(c)->d_desc, 4096))) return rc;
    if ((rc = ensure_host(c, cur_slot(c)->h_desc, 4096))) return rc;
    uint8_t* h = (uint8_t*)cur_slot(c)->h_desc.p;
    uint8_t* d = (uint8_t*)cur_slot(c)->d_desc.p;
    // in_off u64 @0, out_off u64 @8, in_len u32 @16, out_cap u32 @20, prefix_len @24, abs_base @28, chain_first @32,
    // chain_count @36 | results: out_len u32 @64, status i32 @68
    memset(h, 0, 128);
    *(uint64_t*)(h + 0) = cursor;
    *(uint32_t*)(h + 16) = (uint32_t)(n - cursor);
    *(uint32_t*)(h + 20) = (uint32_t)capc;
    *(uint32_t*)(h + 24) = (uint32_t)cursor;
    *(uint32_t*)(h + 28) = (uint32_t)t->offset;
    *(uint32_t*)(h + 32) = 0;
    *(uint32_t*)(h + 36) = 1;
    cudaStream_t s = cur_slot(c)->stream;
    LZF_CU(c, cudaMemcpyAsync(d, h, 128, cudaMemcpyHostToDevice, s));
    LZF_CU(c, cudaMemcpyAsync(cur_slot(c)->d_io_in.p, in, n, cudaMemcpyHostToDevice, s));
    lzf::EncodeArgs ch;
    memset(&ch, 0, sizeof(ch));
    ch.prefix_len = (const uint32_t*)(d + 24); ch.abs_base = (const uint32_t*)(d + 28);
    ch.chain_first = (const uint32_t*)(d + 32); ch.chain_count = (const uint32_t*)(d + 36); ch.nchains = 1;
    ch.table_io = t->d_slots;
    ch.allow_slot_wrap = panic_zone ? 1u : 0u;
    // max_block_len 0 ("unknown") keeps plain slots: the carried dict is the reference's, value for value
    rc = compress_blocks_impl(c, (const uint8_t*)cur_slot(c)->d_io_in.p, (const uint64_t*)d, (const uint32_t*)(d + 16), 1, t->hashlog,
                              t->kind, 0, (uint8_t*)cur_slot(c)->d_io_out.p, (const uint64_t*)(d + 8),
                              (const uint32_t*)(d + 20), (uint32_t*)(d + 64), (int32_t*)(d + 68), nullptr, nullptr, s, &ch);
    if (rc) return rc;
    LZF_CU(c, cudaMemcpyAsync(h + 64, d + 64, 8, cudaMemcpyDeviceToHost, s));
    LZF_CU(c, cudaStreamSynchronize(s));
    const uint32_t olen = *(uint32_t*)(h + 64);
    *status = *(int32_t*)(h + 68);
    if (panic_zone) {
        if (*status != LZF_OK) return fail(c, LZF_ERR_CUDA, "unbounded run refused");
        std::vector<uint8_t> seqs;
        try { seqs.resize(olen); } catch (...) { return fail(c, LZF_ERR_OOM, "host memory"); }
        if (olen) {
            LZF_CU(c, cudaMemcpyAsync(seqs.data(), cur_slot(c)->d_io_out.p, olen, cudaMemcpyDeviceToHost, s));
            LZF_CU(c, cudaStreamSynchronize(s));
        }
        *status = settle_panic_zone(seqs.data(), olen, n, cursor, t->offset, slot_limit, cap);
        if (*status == LZF_OK) { memcpy(out, seqs.data(), olen); *written = olen; }
        return LZF_SUCCESS;
    }
    if (*status == LZF_OK && olen) {
        LZF_CU(c, cudaMemcpyAsync(out, cur_slot(c)->d_io_out.p, olen, cudaMemcpyDeviceToHost, s));
        LZF_CU(c, cudaStreamSynchronize(s));
    }
    *written = *status == LZF_OK ? olen : 0;
    return LZF_SUCCESS;
}

extern "C" int lzf_raw_decompress(lzf_ctx* c, const uint8_t* in, size_t n, const uint8_t* prefix, size_t plen,
                                  uint8_t* out, size_t out_cap, size_t out_limit, size_t* out_len, int32_t* status) {
    if (!c || !out_len || !status || (n && !in) || (plen && !prefix) || (out_cap && !out)) return LZF_ERR_INVALID_ARG;
    *out_len = 0;
    *status = LZF_OK;
    if (n > 0x7fffffffull || plen > 0xffffffffull) return fail(c, LZF_ERR_UNSUPPORTED, "block larger than 2 GiB");
    LZF_CU(c, cudaSetDevice(c->device));
    const size_t capc = out_cap > 0xffffffffull ? 0xffffffffull : out_cap;
    const size_t limc = out_limit > 0xffffffffull ? 0xffffffffull : out_limit;
    int rc;
    if ((rc = ensure_dev(c, cur_slot(c)->d_io_in, n + plen + 128))) return rc;
    if ((rc = ensure_dev(c, cur_slot(c)->d_io_out, capc + 64))) return rc;
    if ((rc = ensure_dev(c, cur_slot(c)->d_desc, 4096))) return rc;
    if ((rc = ensure_host(c, cur_slot(c)->h_desc, 4096))) return rc;
    uint8_t* h = (uint8_t*)cur_slot(c)->h_desc.p;
    uint8_t* d = (uint8_t*)cur_slot(c)->d_desc.p;
    const size_t poff = (n + 63) / 64 * 64;
    // in_off @0, out_off @8, prefix_off @16, in_len @24, out_cap @28, out_limit @32, prefix_len @36 | out_len @64, status @68
    memset(h, 0, 128);
    *(uint64_t*)(h + 16) = poff;
    *(uint32_t*)(h + 24) = (uint32_t)n;
    *(uint32_t*)(h + 28) = (uint32_t)capc;
    *(uint32_t*)(h + 32) = (uint32_t)limc;
    *(uint32_t*)(h + 36) = (uint32_t)plen;
    cudaStream_t s = cur_slot(c)->stream;
    LZF_CU(c, cudaMemcpyAsync(d, h, 128, cudaMemcpyHostToDevice, s));
    uint8_t* din = (uint8_t*)cur_slot(c)->d_io_in.p;
    if (n) LZF_CU(c, cudaMemcpyAsync(din, in, n, cudaMemcpyHostToDevice, s));
    if (plen) LZF_CU(c, cudaMemcpyAsync(din + poff, prefix, plen, cudaMemcpyHostToDevice, s));
    rc = decompress_blocks_impl(c, din, (const uint64_t*)d, (const uint32_t*)(d + 24), 1, plen ? din : nullptr,
                                (const uint64_t*)(d + 16), (const uint32_t*)(d + 36), (uint8_t*)cur_slot(c)->d_io_out.p,
                                (const uint64_t*)(d + 8), (const uint32_t*)(d + 28), (const uint32_t*)(d + 32),
                                (uint32_t*)(d + 64), (int32_t*)(d + 68), nullptr, s);
    if (rc) return rc;
    LZF_CU(c, cudaMemcpyAsync(h + 64, d + 64, 8, cudaMemcpyDeviceToHost, s));
    LZF_CU(c, cudaStreamSynchronize(s));
    const uint32_t olen = *(uint32_t*)(h + 64);
    *status = *(int32_t*)(h + 68);
    const size_t ncopy = olen < capc ? olen : capc;
    if (ncopy) {
        LZF_CU(c, cudaMemcpyAsync(out, cur_slot(c)->d_io_out.p, ncopy, cudaMemcpyDeviceToHost, s));
        LZF_CU(c, cudaStreamSynchronize(s));
    }
    *out_len = olen;
    return LZF_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
// frame layer: settings / header
// ------------------------------------------------------------------------------------------------
extern "C" void lzf_settings_default(lzf_settings* s) {     // src/framed/compress.rs:44-55
    if (!s) return;
    memset(s, 0, sizeof(*s));
    s->independent_blocks = 1;
    s->block_checksums = 0;
    s->content_checksum = 1;
    s->block_size = 4u * 1024 * 1024;
    s->hashlog = 12;
}

extern "C" size_t lzf_frame_bound(const lzf_settings* s, size_t n) {
    const size_t bs = (s && s->block_size) ? (size_t)s->block_size : 1;
    const size_t nblocks = (n + bs - 1) / bs;
    return 19 + n + nblocks * 8 + 8;
}

extern "C" int lzf_frame_parse_header(const uint8_t* in, size_t n, lzf_frame_info* info, int32_t* detail) {
    int32_t d = 0;
    lzf_frame_info tmp;
    if (!in && n) return LZF_F_INPUT_ERROR;
    const int rc = lzf::parse_frame_header(in, n, info ? info : &tmp, &d);
    if (detail) *detail = d;
    return rc;
}

namespace {

// header bytes of compress_internal (src/framed/compress.rs:163-200).  Returns LZF_F_*.
int build_header(const lzf_settings* s, uint64_t content_size, uint8_t* hdr, uint32_t* hlen) {
    uint8_t flags = 0;
    if (s->independent_blocks) flags |= lzf::kFlagIndependent;
    if (s->block_checksums) flags |= lzf::kFlagBlockChecksums;
    if (s->content_checksum) flags |= lzf::kFlagContentChecksum;
    if (s->has_dictionary_id) flags |= lzf::kFlagDictionaryId;
    if (s->has_content_size) flags |= lzf::kFlagContentSize;
    uint8_t bd = 0;
    const int b = lzf::bd_new(s->block_size, &bd);            // compress.rs:183
    if (b == 2) return LZF_F_PANIC;
    if (b == 1) return LZF_F_INVALID_BLOCK_SIZE;
    uint32_t h = 0;
    wr32(hdr, LZF_MAGIC); h = 4;
    hdr[h++] = (uint8_t)((1u << 6) | flags);
    hdr[h++] = bd;
    if (s->has_content_size) { wr64(hdr + h, content_size); h += 8; }
    if (s->has_dictionary_id) { wr32(hdr + h, s->dictionary_id); h += 4; }
    hdr[h] = (uint8_t)(lzf::xxh32_scalar(hdr + 4, h - 4) >> 8);   // compress.rs:197-199
    h++;
    *hlen = h;
    return LZF_F_OK;
}

// ------------------------------------------------------------------------------------------------
// frame compress over device-resident plaintext.  All blocks of all frames: ONE encode launch,
// content checksums on a side stream, then layout + assembly.
// ------------------------------------------------------------------------------------------------
int frames_compress_core(lzf_ctx* c, const lzf_settings* s, const uint8_t* d_in, const uint64_t* in_off,
                         const uint64_t* in_len, uint32_t nframes, uint8_t* d_out, const uint64_t* out_off,
                         const uint64_t* out_cap, uint64_t* out_len, int32_t* status, cudaStream_t st,
                         const InputFeed* feed = nullptr, const uint32_t** deferred_hash = nullptr) {
    // deferred_hash (host-buffer pipeline): the frames are laid out and assembled without waiting for the content
    // checksums — only the last 4 bytes of a frame depend on them — and the caller patches those from
    // *deferred_hash (device, one u32 per frame, complete once the slot's side stream has drained)
    if (deferred_hash) *deferred_hash = nullptr;
    if (!s || (nframes && (!in_off || !in_len || !out_off || !out_cap || !out_len || !status)))
        return fail(c, LZF_ERR_INVALID_ARG, "null pointer");
    for (uint32_t f = 0; f < nframes; f++) { out_len[f] = 0; status[f] = LZF_F_OK; }
    if (nframes == 0) return LZF_SUCCESS;
    const bool dependent = !s->independent_blocks;
    const uint64_t dlen = s->dictionary ? s->dictionary_len : 0;
    if (dlen > 0x7fffffffull) return fail(c, LZF_ERR_UNSUPPORTED, "dictionary larger than 2 GiB");
    const bool chained = dependent || dlen > 0;
    uint32_t hashlog = s->hashlog ? s->hashlog : 12;

    // settings-level failures are identical for every frame
    {
        uint8_t hdr[20]; uint32_t hl;
        const int hs = build_header(s, 0, hdr, &hl);
        if (hs != LZF_F_OK) { for (uint32_t f = 0; f < nframes; f++) status[f] = hs; return LZF_SUCCESS; }
    }
    const uint64_t bs = s->block_size;

    // ---- plan
    // A dependent-block stream that leaves the u32 slot width makes the reference panic inside THAT frame ("EncoderTable
    // contract violated", src/raw/compress/mod.rs:67; the last 7 positions of the stream are never inserted): the frame is
    // answered with LZF_F_PANIC and takes no part in the launch (planned as empty), the other frames are unaffected.
    std::vector<uint8_t> frame_panics(nframes, 0);
    for (uint32_t f = 0; f < nframes; f++) frame_panics[f] = dependent && dlen + in_len[f] > c->tune.pos_limit + 7;
    auto planned_len = [&](uint32_t f) -> uint64_t { return frame_panics[f] ? 0 : in_len[f]; };
    uint64_t nblocks64 = 0;
    for (uint32_t f = 0; f < nframes; f++) nblocks64 += (planned_len(f) + bs - 1) / bs;
    if (nblocks64 > 0x7fffffffull) return fail(c, LZF_ERR_UNSUPPORTED, "too many blocks in one call");
    const uint32_t nblocks = (uint32_t)nblocks64;

    Arena da;   // descriptor arena (host-written)
    const size_t o_first = da.take((size_t)nframes * 4), o_nblk = da.take((size_t)nframes * 4);
    const size_t o_hdr = da.take((size_t)nframes * 20);
    const size_t o_foff = da.take((size_t)nframes * 8), o_fcap = da.take((size_t)nframes * 8);
    const size_t o_hoff = da.take((size_t)nframes * 8), o_hlen = da.take((size_t)nframes * 8);
    const size_t o_bin_off = da.take((size_t)nblocks * 8), o_bin_len = da.take((size_t)nblocks * 4);
    const size_t o_bc_off = da.take((size_t)nblocks * 8);
    // dependent blocks / dictionary: history length, stream base and priming per block, chains of blocks that
    // share a table, and the blocks that need a [dictionary | block] staging copy
    const size_t o_pfx = da.take(chained ? (size_t)nblocks * 4 : 0), o_abs = da.take(chained ? (size_t)nblocks * 4 : 0);
    const size_t o_prime = da.take(chained ? (size_t)nblocks * 4 : 0);
    const size_t o_cfirst = da.take(chained ? (size_t)nblocks * 4 : 0), o_ccount = da.take(chained ? (size_t)nblocks * 4 : 0);
    const size_t o_ssrc = da.take(chained ? (size_t)nblocks * 8 : 0), o_sdst = da.take(chained ? (size_t)nblocks * 8 : 0);
    const size_t o_slen = da.take(chained ? (size_t)nblocks * 4 : 0);
    Arena ra;   // result arena (device-written)
    const size_t r_clen = ra.take((size_t)nblocks * 4), r_bst = ra.take((size_t)nblocks * 4);
    const size_t r_xs = ra.take((size_t)nblocks * 4), r_dst = ra.take((size_t)nblocks * 8);
    const size_t r_chash = ra.take((size_t)nframes * 4);
    const size_t r_flen = ra.take((size_t)nframes * 8), r_fst = ra.take((size_t)nframes * 4);

    int rc;
    if ((rc = ensure_host(c, cur_slot(c)->h_desc, da.used))) return rc;
    if ((rc = ensure_dev(c, cur_slot(c)->d_desc, da.used))) return rc;
    if ((rc = ensure_dev(c, cur_slot(c)->d_res, ra.used))) return rc;
    if ((rc = ensure_host(c, cur_slot(c)->h_res, ra.used))) return rc;
    uint8_t* h = (uint8_t*)cur_slot(c)->h_desc.p;
    uint8_t* d = (uint8_t*)cur_slot(c)->d_desc.p;
    uint8_t* r = (uint8_t*)cur_slot(c)->d_res.p;

    uint32_t* first = (uint32_t*)(h + o_first); uint32_t* nblk = (uint32_t*)(h + o_nblk);
    uint8_t* hdrs = h + o_hdr;
    uint64_t* foff = (uint64_t*)(h + o_foff); uint64_t* fcap = (uint64_t*)(h + o_fcap);
    uint64_t* hoff = (uint64_t*)(h + o_hoff); uint64_t* hlen = (uint64_t*)(h + o_hlen);
    uint64_t* bin_off = (uint64_t*)(h + o_bin_off); uint32_t* bin_len = (uint32_t*)(h + o_bin_len);
    uint64_t* bc_off = (uint64_t*)(h + o_bc_off);
    uint32_t* pfx = (uint32_t*)(h + o_pfx); uint32_t* absb = (uint32_t*)(h + o_abs); uint32_t* prime = (uint32_t*)(h + o_prime);
    uint32_t* cfirst = (uint32_t*)(h + o_cfirst); uint32_t* ccount = (uint32_t*)(h + o_ccount);
    uint64_t* ssrc = (uint64_t*)(h + o_ssrc); uint64_t* sdst = (uint64_t*)(h + o_sdst); uint32_t* slen = (uint32_t*)(h + o_slen);
    uint32_t nchains = 0, nstaged = 0;
    uint64_t stage_total = 0, max_pos = 0;
    uint32_t b = 0;
    uint64_t comp_total = 0;
    uint32_t max_block_len = 0;
    for (uint32_t f = 0; f < nframes; f++) {
        first[f] = b;
        const uint64_t n = planned_len(f);
        const uint32_t nb = (uint32_t)((n + bs - 1) / bs);
        nblk[f] = nb;
        uint32_t hl = 0;
        const uint64_t csize = s->has_content_size == 2 ? n : s->content_size;   // compress_with_size vs _unchecked
        build_header(s, csize, hdrs + (size_t)f * 20 + 1, &hl);
        hdrs[(size_t)f * 20] = (uint8_t)hl;
        foff[f] = out_off[f]; fcap[f] = out_cap[f];
        hoff[f] = in_off[f]; hlen[f] = n;
        for (uint32_t i = 0; i < nb; i++, b++) {
            const uint64_t o = (uint64_t)i * bs;
            const uint32_t l = (uint32_t)((n - o) < bs ? (n - o) : bs);           // compress.rs:227 take(block_size)
            bin_off[b] = in_off[f] + o;
            bin_len[b] = l;
            bc_off[b] = comp_total;
            comp_total += ((uint64_t)l + 15) / 16 * 16;
            if (l > max_block_len) max_block_len = l;
            if (chained) {
                // in_buffer in front of this block (compress.rs:218-222,265-275): the dictionary for block 0 and for
                // every independent block; the last 64 KiB of dictionary ++ plaintext for later dependent blocks
                // (entirely plaintext, because every earlier block is a full block of >= 64 KiB)
                const bool dict_history = dlen && (i == 0 || !dependent);
                const uint64_t total_before = dlen + (dependent ? o : 0);
                const uint64_t hist = dict_history ? dlen : (dependent && i > 0 ? (total_before < LZF_WINDOW_SIZE ? total_before : LZF_WINDOW_SIZE) : 0);
                const uint64_t base = dict_history || !dependent ? 0 : total_before - hist;   // table.offset (compress.rs:273)
                pfx[b] = (uint32_t)hist;
                absb[b] = (uint32_t)base;
                prime[b] = dict_history ? (uint32_t)dlen : 0;
                if (base + hist + l > max_pos) max_pos = base + hist + l;
                if (!dependent || i == 0) { cfirst[nchains] = b; ccount[nchains] = 1; nchains++; }
                else ccount[nchains - 1]++;
                if (dict_history) {
                    ssrc[nstaged] = bin_off[b]; slen[nstaged] = l; sdst[nstaged] = stage_total;
                    stage_total += (dlen + l + 31) / 16 * 16;
                    nstaged++;
                }
            }
        }
    }
    if ((rc = ensure_dev(c, cur_slot(c)->d_comp, comp_total + 64))) return rc;
    // with a dictionary every block address becomes absolute: staged blocks live in their own buffer
    const uint8_t* enc_base = d_in;
    if (dlen) {
        if ((rc = ensure_dev(c, cur_slot(c)->d_dict, dlen + 16))) return rc;
        if ((rc = ensure_dev(c, cur_slot(c)->d_aux, stage_total + 64))) return rc;
        LZF_CU(c, cudaMemcpyAsync(cur_slot(c)->d_dict.p, s->dictionary, dlen, cudaMemcpyHostToDevice, st));
        enc_base = nullptr;
        uint32_t k = 0;
        for (uint32_t bb = 0; bb < nblocks; bb++) {
            if (prime[bb]) { bin_off[bb] = (uint64_t)(uintptr_t)((uint8_t*)cur_slot(c)->d_aux.p + sdst[k] + dlen); k++; }
            else bin_off[bb] = (uint64_t)(uintptr_t)(d_in + bin_off[bb]);
        }
    }

    LZF_CU(c, cudaMemcpyAsync(d, h, da.used, cudaMemcpyHostToDevice, st));
    if (s->content_checksum) LZF_CU(c, cudaEventRecord(cur_slot(c)->ev_fork, st));
    // the slice feed only serves independent blocks without history, parsed one warp per block
    const bool fed = feed && !chained && nblocks && plan_segments(c, nblocks, max_block_len, LZF_TABLE_U32) == 1;
    if (feed && !fed) { if ((rc = feed->start())) return rc; LZF_CU(c, cudaStreamWaitEvent(st, feed->ready, 0)); }
    if (nblocks) {
        lzf::EncodeArgs ch;
        memset(&ch, 0, sizeof(ch));
        if (chained) {
            if (nstaged) {
                lzf::StageArgs sa;
                sa.n = nstaged; sa.dict = (const uint8_t*)cur_slot(c)->d_dict.p; sa.dlen = (uint32_t)dlen;
                sa.in = d_in; sa.src_off = (const uint64_t*)(d + o_ssrc); sa.len = (const uint32_t*)(d + o_slen);
                sa.dst = (uint8_t*)cur_slot(c)->d_aux.p; sa.dst_off = (const uint64_t*)(d + o_sdst);
                LZF_LAUNCHED(c, lzf_launch_stage_dict(&sa, max_block_len, st), 1);
            }
            ch.prefix_len = (const uint32_t*)(d + o_pfx); ch.abs_base = (const uint32_t*)(d + o_abs);
            ch.prime_len = (const uint32_t*)(d + o_prime);
            ch.chain_first = (const uint32_t*)(d + o_cfirst); ch.chain_count = (const uint32_t*)(d + o_ccount);
            ch.nchains = nchains; ch.max_pos = max_pos;
        }
        rc = compress_blocks_impl(c, enc_base, (const uint64_t*)(d + o_bin_off), (const uint32_t*)(d + o_bin_len), nblocks,
                                  hashlog, LZF_TABLE_U32, max_block_len, (uint8_t*)cur_slot(c)->d_comp.p,
                                  (const uint64_t*)(d + o_bc_off), nullptr, (uint32_t*)(r + r_clen), (int32_t*)(r + r_bst),
                                  nullptr, s->block_checksums ? (uint32_t*)(r + r_xs) : nullptr, st,
                                  chained ? &ch : nullptr, fed ? feed : nullptr);
        if (rc) return rc;
        if (fed) LZF_CU(c, cudaStreamWaitEvent(st, feed->ready, 0));      // stored blocks are copied from the plaintext
    }
    // content checksum of each frame's plaintext (compress.rs:172,233-235,279-281) on the side stream
    if (s->content_checksum) {
        LZF_CU(c, cudaStreamWaitEvent(cur_slot(c)->side, cur_slot(c)->ev_fork, 0));
        if (fed) LZF_CU(c, cudaStreamWaitEvent(cur_slot(c)->side, feed->ready, 0));       // the whole plaintext
        uint64_t plain_total = 0;
        for (uint32_t f = 0; f < nframes; f++) plain_total += in_len[f];
        LZF_LAUNCHED(c, lzf_launch_xxh32_ranges(d_in, (const uint64_t*)(d + o_hoff), (const uint64_t*)(d + o_hlen), nframes,
                                               (uint32_t*)(r + r_chash), cur_slot(c)->side, plain_total / nframes >= (1u << 20)), 1);
        LZF_CU(c, cudaEventRecord(cur_slot(c)->ev_join, cur_slot(c)->side));
    }
    if (s->content_checksum) {
        if (deferred_hash) *deferred_hash = (const uint32_t*)(r + r_chash);
        else LZF_CU(c, cudaStreamWaitEvent(st, cur_slot(c)->ev_join, 0));
    }
    lzf::LayoutArgs la;
    memset(&la, 0, sizeof(la));
    la.nframes = nframes;
    la.first_block = (const uint32_t*)(d + o_first); la.nblocks = (const uint32_t*)(d + o_nblk);
    la.blk_in_len = (const uint32_t*)(d + o_bin_len); la.blk_comp_len = (const uint32_t*)(r + r_clen);
    la.blk_status = (const int32_t*)(r + r_bst);
    la.block_checksums = s->block_checksums ? 1 : 0; la.content_checksum = s->content_checksum ? 1 : 0;
    la.headers = d + o_hdr;
    la.out = d_out; la.out_off = (const uint64_t*)(d + o_foff); la.out_cap = (const uint64_t*)(d + o_fcap);
    la.content_hash = (const uint32_t*)(r + r_chash);
    la.blk_dst = (uint64_t*)(r + r_dst);
    // frame lengths and statuses are only read by the host: written straight into its pinned, device-visible arena
    uint8_t* hr = (uint8_t*)cur_slot(c)->h_res.p;
    la.frame_len = (uint64_t*)(hr + r_flen); la.frame_status = (int32_t*)(hr + r_fst);
    LZF_LAUNCHED(c, lzf_launch_layout(&la, st), 1);
    if (nblocks) {
        lzf::AssembleArgs aa;
        memset(&aa, 0, sizeof(aa));
        aa.nblocks = nblocks;
        aa.in = enc_base; aa.blk_in_off = (const uint64_t*)(d + o_bin_off); aa.blk_in_len = (const uint32_t*)(d + o_bin_len);
        aa.comp = (const uint8_t*)cur_slot(c)->d_comp.p; aa.blk_comp_off = (const uint64_t*)(d + o_bc_off);
        aa.blk_comp_len = (const uint32_t*)(r + r_clen); aa.blk_status = (const int32_t*)(r + r_bst);
        aa.blk_xxh_stored = s->block_checksums ? (const uint32_t*)(r + r_xs) : nullptr;
        aa.blk_dst = (const uint64_t*)(r + r_dst); aa.out = d_out;
        LZF_LAUNCHED(c, lzf_launch_assemble(&aa, max_block_len, st), 1);
    }
    LZF_CU(c, cudaStreamSynchronize(st));
    const uint64_t* flen = (const uint64_t*)(hr + r_flen);
    const int32_t* fst = (const int32_t*)(hr + r_fst);
    for (uint32_t f = 0; f < nframes; f++) {
        out_len[f] = frame_panics[f] ? 0 : flen[f];
        status[f] = frame_panics[f] ? LZF_F_PANIC : fst[f];
    }
    return LZF_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
// frame decompress over device-resident frames (independent blocks)
// ------------------------------------------------------------------------------------------------
struct DecodeOut {          // optional extra per-frame results
    uint64_t* consumed;     // nullable
    int32_t* detail;        // nullable
    // host-buffer pipeline hook: called once the block kernel is queued on the stream (before the host
    // waits for its results), so the D2H copy of the plaintext can start while the outcome is resolved
    // (argument: upper bound of the plaintext bytes the frames can decode to, from their block counts)
    std::function<int(uint64_t)> after_decode;
    // nullable, one flag per frame: set when the frame's plaintext was (re)written AFTER after_decode ran — frames with
    // short non-final blocks are decoded again at their exact positions — so a copy started by the hook holds the
    // pass-1 placement of that frame and must be repeated
    uint8_t* moved = nullptr;
};

int frames_decompress_core(lzf_ctx* c, const uint8_t* d_in, const uint64_t* in_off, const uint64_t* in_len,
                           uint32_t nframes, uint8_t* d_out, const uint64_t* out_off, const uint64_t* out_cap,
                           uint64_t* out_len, int32_t* status, DecodeOut extra, cudaStream_t st,
                           const uint8_t* d_dict = nullptr, uint64_t dlen = 0) {
    if (nframes && (!in_off || !in_len || !out_off || !out_cap || !out_len || !status))
        return fail(c, LZF_ERR_INVALID_ARG, "null pointer");
    for (uint32_t f = 0; f < nframes; f++) {
        out_len[f] = 0; status[f] = LZF_F_OK;
        if (extra.consumed) extra.consumed[f] = 0;
        if (extra.detail) extra.detail[f] = 0;
    }
    if (nframes == 0) return LZF_SUCCESS;
    int rc;

    // ---- pass 0: parse headers and count blocks (LZ4FrameReader::new + the length-word chase)
    Arena da;
    const size_t o_ioff = da.take((size_t)nframes * 8), o_ilen = da.take((size_t)nframes * 8);
    const size_t o_ooff = da.take((size_t)nframes * 8), o_ocap = da.take((size_t)nframes * 8);
    const size_t o_first = da.take((size_t)nframes * 4), o_nblk = da.take((size_t)nframes * 4);
    const size_t o_hoff = da.take((size_t)nframes * 8), o_hlen = da.take((size_t)nframes * 8);
    const size_t frame_desc_bytes = da.used;
    if ((rc = ensure_host(c, cur_slot(c)->h_desc, da.used))) return rc;
    if ((rc = ensure_dev(c, cur_slot(c)->d_desc, da.used))) return rc;
    uint8_t* h = (uint8_t*)cur_slot(c)->h_desc.p;
    uint8_t* d = (uint8_t*)cur_slot(c)->d_desc.p;
    memcpy(h + o_ioff, in_off, (size_t)nframes * 8);
    memcpy(h + o_ilen, in_len, (size_t)nframes * 8);
    memcpy(h + o_ooff, out_off, (size_t)nframes * 8);
    memcpy(h + o_ocap, out_cap, (size_t)nframes * 8);

    Arena ra;
    const size_t r_walk = ra.take((size_t)nframes * sizeof(lzf::WalkFrame));
    if ((rc = ensure_dev(c, cur_slot(c)->d_res, ra.used))) return rc;
    if ((rc = ensure_host(c, cur_slot(c)->h_res, ra.used))) return rc;
    LZF_CU(c, cudaMemcpyAsync(d, h, o_first, cudaMemcpyHostToDevice, st));
    lzf::WalkArgs wa;
    memset(&wa, 0, sizeof(wa));
    wa.nframes = nframes; wa.mode = 0;
    wa.in = d_in; wa.in_off = (const uint64_t*)(d + o_ioff); wa.in_len = (const uint64_t*)(d + o_ilen);
    // Everything the HOST waits for in this function (walk results, per-block lengths and statuses, checksums) is written
    // by the kernels straight into the slot's pinned, device-visible result arena (unified addressing: the host pointer
    // is the device pointer).  A D2H copy of a few KiB would queue on the one D2H copy engine behind the hundreds of MiB
    // of plaintext other chunks are sending home, and the next launch of this chunk would wait for them.
    wa.frames = (lzf::WalkFrame*)((uint8_t*)cur_slot(c)->h_res.p + r_walk);
    LZF_LAUNCHED(c, lzf_launch_walk(&wa, st), 1);
    std::vector<lzf::WalkFrame> wf(nframes);
    LZF_CU(c, cudaStreamSynchronize(st));
    memcpy(wf.data(), (uint8_t*)cur_slot(c)->h_res.p + r_walk, (size_t)nframes * sizeof(lzf::WalkFrame));

    uint64_t nblocks64 = 0;
    bool any_dependent = false;
    for (uint32_t f = 0; f < nframes; f++) {
        if (wf[f].header_status != LZF_F_OK) { wf[f].nblocks = 0; continue; }
        if (!(wf[f].flags & lzf::kFlagIndependent)) any_dependent = true;
        nblocks64 += wf[f].nblocks;
    }
    if (dlen > 0xffffffffull) return fail(c, LZF_ERR_UNSUPPORTED, "dictionary larger than 4 GiB");
    if (nblocks64 > 0x7fffffffull) return fail(c, LZF_ERR_UNSUPPORTED, "too many blocks in one call");
    const uint32_t nblocks = (uint32_t)nblocks64;

    // ---- pass 1: block descriptors, decode, checksums
    uint32_t* first = (uint32_t*)(h + o_first);
    uint32_t* nblk = (uint32_t*)(h + o_nblk);
    {
        uint32_t b = 0;
        for (uint32_t f = 0; f < nframes; f++) { first[f] = b; nblk[f] = wf[f].nblocks; b += wf[f].nblocks; }
    }
    Arena ba;   // device-only block arrays
    ba.used = frame_desc_bytes;
    const size_t o_bin_off = ba.take((size_t)nblocks * 8), o_bword = ba.take((size_t)nblocks * 4);
    const size_t o_boff = ba.take((size_t)nblocks * 8);
    const size_t o_bcap = ba.take((size_t)nblocks * 4), o_blim = ba.take((size_t)nblocks * 4);
    const size_t o_bplen = ba.take((size_t)nblocks * 8);
    Arena rb;
    const size_t r_olen = rb.take((size_t)nblocks * 4), r_bst = rb.take((size_t)nblocks * 4);
    const size_t r_bxxh = rb.take((size_t)nblocks * 4), r_bcks = rb.take((size_t)nblocks * 4);
    const size_t r_bend = rb.take((size_t)nblocks * 8);
    const size_t r_chash = rb.take((size_t)nframes * 4);
    if (ba.used > cur_slot(c)->d_desc.cap) {
        // growing d_desc would drop the frame arrays: re-upload them afterwards
        if ((rc = ensure_dev(c, cur_slot(c)->d_desc, ba.used))) return rc;
        d = (uint8_t*)cur_slot(c)->d_desc.p;
        LZF_CU(c, cudaMemcpyAsync(d, h, o_first, cudaMemcpyHostToDevice, st));
        wa.in_off = (const uint64_t*)(d + o_ioff); wa.in_len = (const uint64_t*)(d + o_ilen);
    }
    if ((rc = ensure_host(c, cur_slot(c)->h_res, rb.used))) return rc;
    uint8_t* hr = (uint8_t*)cur_slot(c)->h_res.p;
    LZF_CU(c, cudaMemcpyAsync(d + o_first, h + o_first, o_hoff - o_first, cudaMemcpyHostToDevice, st));

    bool any_block_checksums = false;
    for (uint32_t f = 0; f < nframes; f++)
        if (wf[f].header_status == LZF_F_OK && (wf[f].flags & lzf::kFlagBlockChecksums) && wf[f].nblocks) any_block_checksums = true;

    if (nblocks) {
        wa.mode = 1;
        wa.first_block = (const uint32_t*)(d + o_first);
        wa.out_off = (const uint64_t*)(d + o_ooff); wa.out_cap = (const uint64_t*)(d + o_ocap);
        wa.blk_in_off = (uint64_t*)(d + o_bin_off); wa.blk_len_word = (uint32_t*)(d + o_bword);
        wa.blk_checksum = (uint32_t*)(hr + r_bcks);
        wa.blk_out_off = (uint64_t*)(d + o_boff); wa.blk_out_cap = (uint32_t*)(d + o_bcap);
        wa.blk_out_limit = (uint32_t*)(d + o_blim); wa.blk_payload_len = (uint64_t*)(d + o_bplen);
        wa.blk_end = (uint64_t*)(hr + r_bend);
        LZF_LAUNCHED(c, lzf_launch_walk(&wa, st), 1);
        if (any_block_checksums)    // decompress.rs:228-235: hash of the stored payload
            LZF_LAUNCHED(c, lzf_launch_xxh32_ranges(d_in, (const uint64_t*)(d + o_bin_off), (const uint64_t*)(d + o_bplen),
                                                   nblocks, (uint32_t*)(hr + r_bxxh), st), 1);
        // history of every block (src/framed/decompress.rs:238-248): the dictionary for independent blocks
        // and for the first block of a dependent frame; for block i > 0 of a dependent frame the 64 KiB of
        // output right in front of its own slot — exact whenever the blocks before it were full, which
        // pass 1 assumes and the resolution below verifies
        const bool use_hist = any_dependent || dlen > 0;
        const uint64_t* d_pf_off = nullptr; const uint32_t* d_pf_len = nullptr; const int32_t* d_wait = nullptr; uint32_t* d_done = nullptr;
        std::vector<uint8_t> hx;
        if (use_hist) {
            Arena xa;
            const size_t x_pf = xa.take((size_t)nblocks * 8), x_pl = xa.take((size_t)nblocks * 4), x_wait = xa.take((size_t)nblocks * 4);
            const size_t x_done = xa.take((size_t)nblocks * 4);
            hx.assign(xa.used, 0);
            uint64_t* pf = (uint64_t*)(hx.data() + x_pf); uint32_t* pl = (uint32_t*)(hx.data() + x_pl); int32_t* wt = (int32_t*)(hx.data() + x_wait);
            for (uint32_t f = 0; f < nframes; f++) {
                const bool dep = wf[f].header_status == LZF_F_OK && !(wf[f].flags & lzf::kFlagIndependent);
                const uint64_t bms = wf[f].block_maxsize;
                for (uint32_t i = 0; i < wf[f].nblocks; i++) {
                    const uint32_t b = first[f] + i;
                    const uint64_t rel = (uint64_t)i * bms;
                    wt[b] = -1;
                    if (dep && i > 0 && rel <= out_cap[f]) {
                        const uint64_t wlen = rel < LZF_WINDOW_SIZE ? rel : LZF_WINDOW_SIZE;
                        pf[b] = (uint64_t)(uintptr_t)(d_out + out_off[f] + rel - wlen);
                        pl[b] = (uint32_t)wlen;
                        wt[b] = (int32_t)(b - 1);
                    } else if ((!dep || i == 0) && dlen) {
                        pf[b] = (uint64_t)(uintptr_t)d_dict;
                        pl[b] = (uint32_t)dlen;
                    }
                }
            }
            if ((rc = ensure_dev(c, cur_slot(c)->d_aux, xa.used))) return rc;
            uint8_t* x = (uint8_t*)cur_slot(c)->d_aux.p;
            LZF_CU(c, cudaMemcpyAsync(x, hx.data(), xa.used, cudaMemcpyHostToDevice, st));
            d_pf_off = (const uint64_t*)(x + x_pf); d_pf_len = (const uint32_t*)(x + x_pl);
            if (any_dependent) { d_wait = (const int32_t*)(x + x_wait); d_done = (uint32_t*)(x + x_done); }
        }
        rc = decompress_blocks_impl(c, d_in, (const uint64_t*)(d + o_bin_off), (const uint32_t*)(d + o_bword), nblocks,
                                    nullptr, d_pf_off, d_pf_len, d_out, (const uint64_t*)(d + o_boff),
                                    (const uint32_t*)(d + o_bcap), (const uint32_t*)(d + o_blim), (uint32_t*)(hr + r_olen),
                                    (int32_t*)(hr + r_bst), nullptr, st, use_hist, d_wait, d_done);
        if (rc) return rc;
        if (extra.after_decode) {
            uint64_t expect = 0;
            for (uint32_t f = 0; f < nframes; f++) {
                const uint64_t e = (uint64_t)wf[f].nblocks * wf[f].block_maxsize;
                expect += e < out_cap[f] ? e : out_cap[f];
            }
            if ((rc = extra.after_decode(expect))) return rc;
        }
        LZF_CU(c, cudaStreamSynchronize(st));       // lengths, statuses, checksums and block ends are in hr already
    }
    uint32_t* b_olen = (uint32_t*)(hr + r_olen);
    int32_t* b_st = (int32_t*)(hr + r_bst);
    const uint32_t* b_xxh = (const uint32_t*)(hr + r_bxxh);
    const uint32_t* b_cks = (const uint32_t*)(hr + r_bcks);
    const uint64_t* b_end = (const uint64_t*)(hr + r_bend);

    // ---- resolve per-frame outcome in the reference's order (decompress.rs:197-279).  Pass 1 decoded
    // block i of a frame at the fixed slot i * block_maxsize (exact for every frame whose non-final
    // blocks are full, i.e. anything a compressor writes); the kernel reports each block's status
    // and TRUE decoded length even when its slot was too small, so frames with short non-final
    // blocks (hand-crafted, decompress.rs:165-166) are re-decoded below at their exact positions.
    uint64_t* hoff = (uint64_t*)(h + o_hoff);
    uint64_t* hlen = (uint64_t*)(h + o_hlen);
    bool any_hash = false;
    std::vector<uint8_t> want_hash(nframes, 0);
    std::vector<uint32_t> redo_frames;              // independent frames needing exact placement
    std::vector<uint32_t> delivered(nframes, 0);    // blocks whose plaintext the caller receives
    std::vector<uint64_t> p_in_off;                 // host copies of the payload descriptors (filled on demand)
    std::vector<uint32_t> p_word;
    auto fetch_payload_descriptors = [&]() -> int {
        if (!p_in_off.empty() || nblocks == 0) return LZF_SUCCESS;
        p_in_off.resize(nblocks);
        p_word.resize(nblocks);
        LZF_CU(c, cudaMemcpyAsync(p_in_off.data(), d + o_bin_off, (size_t)nblocks * 8, cudaMemcpyDeviceToHost, st));
        LZF_CU(c, cudaMemcpyAsync(p_word.data(), d + o_bword, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, st));
        LZF_CU(c, cudaStreamSynchronize(st));
        return LZF_SUCCESS;
    };
    // One block, decoded synchronously at an exact position with an exact window (slow, exact path of
    // dependent frames with short non-final blocks).
    auto decode_one = [&](uint32_t b, const uint8_t* pfx, uint32_t plen, uint64_t out_pos, uint32_t cap1, uint32_t lim1) -> int {
        int rc1;
        if ((rc1 = ensure_dev(c, cur_slot(c)->d_comp, 4096))) return rc1;
        uint8_t* x = (uint8_t*)cur_slot(c)->d_comp.p;
        uint8_t hb[128];
        memset(hb, 0, sizeof(hb));
        *(uint64_t*)(hb + 0) = p_in_off[b]; *(uint64_t*)(hb + 8) = out_pos; *(uint64_t*)(hb + 16) = (uint64_t)(uintptr_t)pfx;
        *(uint32_t*)(hb + 24) = p_word[b]; *(uint32_t*)(hb + 28) = cap1; *(uint32_t*)(hb + 32) = lim1; *(uint32_t*)(hb + 36) = plen;
        LZF_CU(c, cudaMemcpyAsync(x, hb, 128, cudaMemcpyHostToDevice, st));
        rc1 = decompress_blocks_impl(c, d_in, (const uint64_t*)x, (const uint32_t*)(x + 24), 1, nullptr, (const uint64_t*)(x + 16),
                                     (const uint32_t*)(x + 36), d_out, (const uint64_t*)(x + 8), (const uint32_t*)(x + 28),
                                     (const uint32_t*)(x + 32), (uint32_t*)(x + 64), (int32_t*)(x + 68), nullptr, st, true);
        if (rc1) return rc1;
        LZF_CU(c, cudaMemcpyAsync(hb + 64, x + 64, 8, cudaMemcpyDeviceToHost, st));
        LZF_CU(c, cudaStreamSynchronize(st));
        b_olen[b] = *(uint32_t*)(hb + 64);
        b_st[b] = *(int32_t*)(hb + 68);
        return LZF_SUCCESS;
    };
    for (uint32_t f = 0; f < nframes; f++) {
        hoff[f] = out_off[f]; hlen[f] = 0;
        const lzf::WalkFrame& w = wf[f];
        if (w.header_status != LZF_F_OK) {
            status[f] = w.header_status;
            if (extra.detail) extra.detail[f] = w.header_detail;
            continue;
        }
        const uint64_t bms = w.block_maxsize;
        const bool bc = (w.flags & lzf::kFlagBlockChecksums) != 0;
        const bool dep = !(w.flags & lzf::kFlagIndependent);
        const uint64_t capf = out_cap[f];
        uint64_t o = 0;
        int fs = LZF_F_OK, det = 0;
        uint64_t consumed = w.consumed;
        bool early_stop = false, irregular = false;
        bool exact = false;          // dependent frame: from here on blocks are (re)decoded one by one, exactly placed
        uint32_t i = 0;
        for (; i < w.nblocks; i++) {
            const uint32_t b = first[f] + i;
            if (bc && b_xxh[b] != b_cks[b]) { fs = LZF_F_BLOCK_CHECKSUM_FAIL; break; }                     // :228-235
            if (exact) {
                // window = last 64 KiB of dictionary ++ output so far (decompress.rs:253-269)
                if ((rc = fetch_payload_descriptors())) return rc;
                const uint8_t* pfx;
                uint32_t plen;
                if (o >= LZF_WINDOW_SIZE) { pfx = d_out + out_off[f] + o - LZF_WINDOW_SIZE; plen = LZF_WINDOW_SIZE; }
                else if (o == 0) { pfx = d_dict; plen = (uint32_t)dlen; }
                else {
                    const uint64_t from_dict = dlen < LZF_WINDOW_SIZE - o ? dlen : LZF_WINDOW_SIZE - o;
                    if ((rc = ensure_dev(c, cur_slot(c)->d_aux, 2 * LZF_WINDOW_SIZE))) return rc;
                    uint8_t* wb = (uint8_t*)cur_slot(c)->d_aux.p;
                    if (from_dict) LZF_CU(c, cudaMemcpyAsync(wb, d_dict + dlen - from_dict, from_dict, cudaMemcpyDeviceToDevice, st));
                    LZF_CU(c, cudaMemcpyAsync(wb + from_dict, d_out + out_off[f], o, cudaMemcpyDeviceToDevice, st));
                    pfx = wb; plen = (uint32_t)(from_dict + o);
                }
                const uint64_t room = capf - o;
                if ((rc = decode_one(b, pfx, plen, out_off[f] + o, (uint32_t)(room < bms ? room : bms), (uint32_t)bms))) return rc;
            }
            const int bst = b_st[b];
            if (bst >= LZF_UNEXPECTED_END && bst <= LZF_INVALID_DEDUP_OFFSET) { fs = LZF_F_CODEC_ERROR; det = bst; break; }   // :247-248
            const uint64_t ol = b_olen[b];
            if (ol > bms) { fs = LZF_F_BLOCK_SIZE_OVERFLOW; break; }                                        // :272-274
            if (o + ol > capf) { fs = LZF_F_WRITE_ERROR; break; }
            if (!exact) {
                const uint64_t rel = (uint64_t)i * bms;
                const uint64_t slot_cap = rel < capf ? (bms < capf - rel ? bms : capf - rel) : 0;
                if (ol && (o != rel || ol > slot_cap)) irregular = true;
                // a dependent frame stops being regular at the first short block: everything after it saw the
                // wrong window in pass 1 and is decoded again, one block at a time
                if (dep && ol != bms) exact = true;
            }
            o += ol;
            if (ol == 0) { early_stop = true; consumed = b_end[b]; i++; break; }   // read_to_end sees Ok(0): decompress.rs:54-61,286
        }
        delivered[f] = i;
        if (fs == LZF_F_OK && !early_stop) {
            fs = w.term_status;                                  // EndMark / truncation / length-word overflow
            if (fs == LZF_F_OK && (w.flags & lzf::kFlagContentChecksum)) { want_hash[f] = 1; any_hash = true; }
        }
        status[f] = fs;
        out_len[f] = o;
        hlen[f] = o;
        if (irregular && !dep) redo_frames.push_back(f);
        if (extra.moved) extra.moved[f] = (irregular || exact) ? 1 : 0;
        if (extra.detail) extra.detail[f] = det;
        if (extra.consumed) extra.consumed[f] = consumed;
    }
    if (!redo_frames.empty()) {
        uint32_t nredo = 0;
        for (uint32_t f : redo_frames) nredo += delivered[f];
        if ((rc = fetch_payload_descriptors())) return rc;
        Arena xa;
        const size_t x_in_off = xa.take((size_t)nredo * 8), x_word = xa.take((size_t)nredo * 4);
        const size_t x_out_off = xa.take((size_t)nredo * 8), x_cap = xa.take((size_t)nredo * 4);
        const size_t x_lim = xa.take((size_t)nredo * 4), x_olen = xa.take((size_t)nredo * 4), x_st = xa.take((size_t)nredo * 4);
        const size_t x_pf = xa.take((size_t)nredo * 8), x_pl = xa.take((size_t)nredo * 4);
        std::vector<uint8_t> xh(xa.used);
        uint32_t k = 0;
        for (uint32_t f : redo_frames) {
            uint64_t o = 0;
            for (uint32_t i = 0; i < delivered[f]; i++, k++) {
                const uint32_t b = first[f] + i;
                ((uint64_t*)(xh.data() + x_in_off))[k] = p_in_off[b];
                ((uint32_t*)(xh.data() + x_word))[k] = p_word[b];
                ((uint64_t*)(xh.data() + x_out_off))[k] = out_off[f] + o;
                ((uint32_t*)(xh.data() + x_cap))[k] = b_olen[b];
                ((uint32_t*)(xh.data() + x_lim))[k] = (uint32_t)wf[f].block_maxsize;
                ((uint64_t*)(xh.data() + x_pf))[k] = (uint64_t)(uintptr_t)d_dict;
                ((uint32_t*)(xh.data() + x_pl))[k] = (uint32_t)dlen;
                o += b_olen[b];
            }
        }
        if ((rc = ensure_dev(c, cur_slot(c)->d_comp, xa.used))) return rc;     // scratch free on this path
        uint8_t* x = (uint8_t*)cur_slot(c)->d_comp.p;
        LZF_CU(c, cudaMemcpyAsync(x, xh.data(), xa.used, cudaMemcpyHostToDevice, st));
        rc = decompress_blocks_impl(c, d_in, (const uint64_t*)(x + x_in_off), (const uint32_t*)(x + x_word), nredo,
                                    nullptr, (const uint64_t*)(x + x_pf), (const uint32_t*)(x + x_pl), d_out,
                                    (const uint64_t*)(x + x_out_off), (const uint32_t*)(x + x_cap), (const uint32_t*)(x + x_lim),
                                    (uint32_t*)(x + x_olen), (int32_t*)(x + x_st), nullptr, st, true);
        if (rc) return rc;
        LZF_CU(c, cudaStreamSynchronize(st));    // xh must outlive the copy
    }
    if (any_hash) {       // content checksum over the frame's plaintext (decompress.rs:207-211,276-278)
        LZF_CU(c, cudaMemcpyAsync(d + o_hoff, h + o_hoff, frame_desc_bytes - o_hoff, cudaMemcpyHostToDevice, st));
        uint64_t plain_total = 0;
        for (uint32_t f = 0; f < nframes; f++) plain_total += hlen[f];
        LZF_LAUNCHED(c, lzf_launch_xxh32_ranges(d_out, (const uint64_t*)(d + o_hoff), (const uint64_t*)(d + o_hlen), nframes,
                                               (uint32_t*)(hr + r_chash), st, plain_total / nframes >= (1u << 20)), 1);
        LZF_CU(c, cudaStreamSynchronize(st));
        const uint32_t* ch = (const uint32_t*)(hr + r_chash);
        for (uint32_t f = 0; f < nframes; f++)
            if (want_hash[f] && ch[f] != wf[f].content_checksum) status[f] = LZF_F_FRAME_CHECKSUM_FAIL;
    }
    return LZF_SUCCESS;
}

// Host-buffer batches usually come as one contiguous run of frames; then a single cudaMemcpy moves
// the whole run and the device layout mirrors the host layout.  Otherwise frames are packed.
struct HostLayout {
    bool dense = false;
    uint64_t base = 0, span = 0;
    std::vector<uint64_t> dev_off;
};
HostLayout plan_layout(const uint64_t* off, const uint64_t* len, uint32_t n) {
    HostLayout L;
    L.dev_off.resize(n);
    bool dense = n > 0;
    for (uint32_t f = 0; f + 1 < n && dense; f++) dense = off[f + 1] == off[f] + len[f];
    if (dense) {
        L.dense = true;
        L.base = off[0];
        L.span = off[n - 1] + len[n - 1] - off[0];
        for (uint32_t f = 0; f < n; f++) L.dev_off[f] = off[f] - L.base;
    } else {
        uint64_t t = 0;
        for (uint32_t f = 0; f < n; f++) { L.dev_off[f] = t; t += (len[f] + 255) / 256 * 256; }
        L.span = t;
    }
    return L;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// frame entry points
// ------------------------------------------------------------------------------------------------
extern "C" int lzf_frames_compress_device(lzf_ctx* c, const lzf_settings* s, const uint8_t* d_in, const uint64_t* in_off,
                                          const uint64_t* in_len, uint32_t nframes, uint8_t* d_out,
                                          const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len,
                                          int32_t* status) {
    if (!c) return LZF_ERR_INVALID_ARG;
    LZF_CU(c, cudaSetDevice(c->device));
    return frames_compress_core(c, s, d_in, in_off, in_len, nframes, d_out, out_off, out_cap, out_len, status, cur_slot(c)->stream);
}

namespace {

// Consecutive frames are grouped into chunks of roughly `target` payload bytes; chunk i runs in
// pipeline slot i % kSlots.
std::vector<uint32_t> plan_chunks(const uint64_t* len, uint32_t n, uint64_t target, bool ramp = false) {
    std::vector<uint32_t> start;
    uint64_t acc = 0;
    // ramp (the decompress pipeline): the first chunks are short — 1/8, 1/4, 1/2 of the target — so that the first
    // plaintext starts travelling back after a fraction of a chunk's H2D copy instead of a whole one, and so are the
    // last ones, so that little is left to drain once the H2D engine has nothing more to send
    uint64_t total = 0, done = 0;
    for (uint32_t f = 0; f < n; f++) total += len[f];
    uint64_t cur = ramp ? target / 8 : target;
    for (uint32_t f = 0; f < n; f++) {
        if (f == 0 || acc >= cur) {
            start.push_back(f);
            if (f != 0 && ramp) {
                cur = cur * 2 < target ? cur * 2 : target;
                const uint64_t left = total - done;
                while (cur > target / 8 && left < 2 * cur) cur /= 2;       // the tail shrinks again
            }
            acc = 0;
        }
        acc += len[f];
        done += len[f];
    }
    start.push_back(n);
    return start;
}

// Runs `work(i)` for every chunk on up to kSlots host threads, each bound to its own pipeline slot
// (stream + scratch).  Every worker is synchronous on its own stream; the overlap of the H2D copy of
// one chunk with the kernels of another and the D2H copy of a third comes from the streams running
// side by side, and kernels of different chunks share the SMs.
template <typename F>
int run_chunks(lzf_ctx* c, uint32_t nchunks, F work, uint32_t max_workers = kSlots) {
    if (nchunks == 0) return LZF_SUCCESS;
    std::atomic<uint32_t> next{0};
    std::atomic<int> rc_all{LZF_SUCCESS};
    auto worker = [&](int k) {
        cudaSetDevice(c->device);
        tls_slot = &c->slots[k];
        for (;;) {
            const uint32_t i = next.fetch_add(1);
            if (i >= nchunks || rc_all.load() != LZF_SUCCESS) break;
            const int rc = work(i, c->slots[k]);
            if (rc != LZF_SUCCESS) rc_all.store(rc);
        }
        cudaStreamSynchronize(c->slots[k].stream);
        tls_slot = nullptr;
    };
#ifdef LZF_SIMT_EMU
    const uint32_t nworkers = 1;                 // the CPU SIMT test harness is single-threaded
#else
    const uint32_t nworkers = nchunks < max_workers ? nchunks : max_workers;
#endif
    std::vector<std::thread> threads;
    for (uint32_t k = 1; k < nworkers; k++) threads.emplace_back(worker, (int)k);
    worker(0);
    for (auto& t : threads) t.join();
    return rc_all.load();
}

int compress_chunk(lzf_ctx* c, lzf_slot& sl, const lzf_settings* s, uint32_t f0, uint32_t f1, const uint8_t* in,
                   const uint64_t* in_off, const uint64_t* in_len, uint8_t* out, const uint64_t* out_off,
                   const uint64_t* out_cap, uint64_t* out_len, int32_t* status) {
    const uint32_t n = f1 - f0;
    std::vector<uint64_t> dcap(n);
    for (uint32_t f = 0; f < n; f++) {
        const uint64_t need = lzf_frame_bound(s, in_len[f0 + f]);
        dcap[f] = out_cap[f0 + f] < need ? out_cap[f0 + f] : need;
    }
    const HostLayout li = plan_layout(in_off + f0, in_len + f0, n);
    const HostLayout lo = plan_layout(out_off + f0, dcap.data(), n);
    int rc;
    const bool trace = c->tune.trace;                                 // phase times of the chunk on stderr
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    if ((rc = ensure_dev(c, sl.d_io_in, li.span + 256))) return rc;
    if ((rc = ensure_dev(c, sl.d_io_out, lo.span + 256))) return rc;
    uint8_t* din = (uint8_t*)sl.d_io_in.p;
    const uint8_t* dout = (const uint8_t*)sl.d_io_out.p;
    // One warp parses one block at its own (latency-bound) pace, far below PCIe speed, so a chunk of large
    // independent blocks is fed while the kernel already runs: slice k of EVERY block travels before slice k + 1
    // of any (one strided copy per slice), followed by a 4-byte copy that bumps the progress word the warps poll.
    const uint64_t bs = s->block_size;
    const uint64_t slice = c->tune.feed_slice, min_blocks = c->tune.feed_min_blocks;   // 64 KiB: the first slice of 4096 blocks is there after 5 ms
    bool sliced = li.dense && s->independent_blocks && !(s->dictionary && s->dictionary_len) && slice >= 4096 &&
                  bs >= 4 * slice && bs % slice == 0 && bs / slice <= kMaxSlices && li.span >= min_blocks * bs;
    for (uint32_t f = 0; f < n && sliced; f++) sliced = in_len[f0 + f] % bs == 0;
    InputFeed feed{nullptr, 0, nullptr, nullptr};
    if (sliced) {
        uint32_t* d_progress = sl.d_counter + 32;
        feed.d_progress = d_progress; feed.slice_bytes = (uint32_t)slice; feed.ready = sl.ev_feed1;
        feed.start = [&, d_progress]() -> int {
            const uint64_t rows = li.span / bs;
            LZF_CU(c, cudaMemsetAsync(d_progress, 0, 4, sl.stream));
            LZF_CU(c, cudaEventRecord(sl.ev_feed0, sl.stream));
            LZF_CU(c, cudaStreamWaitEvent(sl.copy, sl.ev_feed0, 0));
            for (uint64_t k = 0; k < bs / slice; k++) {
                LZF_CU(c, cudaMemcpy2DAsync(din + k * slice, bs, in + li.base + k * slice, bs, slice, rows, cudaMemcpyHostToDevice, sl.copy));
                LZF_CU(c, cudaMemcpyAsync(d_progress, sl.h_seq + k, 4, cudaMemcpyHostToDevice, sl.copy));
            }
            LZF_CU(c, cudaEventRecord(sl.ev_feed1, sl.copy));
            return LZF_SUCCESS;
        };
    } else if (li.dense) {
        if (li.span) LZF_CU(c, cudaMemcpyAsync(din, in + li.base, li.span, cudaMemcpyHostToDevice, sl.stream));
    } else {
        for (uint32_t f = 0; f < n; f++)
            if (in_len[f0 + f])
                LZF_CU(c, cudaMemcpyAsync(din + li.dev_off[f], in + in_off[f0 + f], in_len[f0 + f], cudaMemcpyHostToDevice, sl.stream));
    }
    const uint32_t* d_hash = nullptr;      // content checksums still being computed on the side stream
    rc = frames_compress_core(c, s, din, li.dev_off.data(), in_len + f0, n, (uint8_t*)sl.d_io_out.p, lo.dev_off.data(),
                              dcap.data(), out_len + f0, status + f0, sl.stream, sliced ? &feed : nullptr, &d_hash);
    if (sliced) cudaStreamSynchronize(sl.copy);
    if (rc) return rc;
    const double t_core = since();
    // compressed frames are much shorter than their capacity: copy each frame's bytes
    for (uint32_t f = f0; f < f1; f++)
        if (status[f] == LZF_F_OK && out_len[f])
            LZF_CU(c, cudaMemcpyAsync(out + out_off[f], dout + lo.dev_off[f - f0], out_len[f], cudaMemcpyDeviceToHost, sl.stream));
    const double t_issued = since();
    std::vector<uint32_t> hashes;
    if (d_hash) {
        // the checksums ran beside the assembly and the D2H copies; they are the last 4 bytes of each frame
        // (src/framed/compress.rs:279-281)
        hashes.resize(n);
        LZF_CU(c, cudaStreamSynchronize(sl.side));
        LZF_CU(c, cudaMemcpyAsync(hashes.data(), d_hash, (size_t)n * 4, cudaMemcpyDeviceToHost, sl.stream));
    }
    LZF_CU(c, cudaStreamSynchronize(sl.stream));
    if (d_hash)
        for (uint32_t f = f0; f < f1; f++)
            if (status[f] == LZF_F_OK && out_len[f] >= 4) wr32(out + out_off[f] + out_len[f] - 4, hashes[f - f0]);
    if (trace)
        fprintf(stderr, "lzf trace: compress chunk of %u frames (%s feed): kernels + results %.1f ms, D2H issued %.1f ms, done %.1f ms\n",
                n, sliced ? "sliced" : "plain", t_core, t_issued, since());
    return LZF_SUCCESS;
}

}  // namespace

extern "C" int lzf_frames_compress(lzf_ctx* c, const lzf_settings* s, const uint8_t* in, const uint64_t* in_off,
                                   const uint64_t* in_len, uint32_t nframes, uint8_t* out, const uint64_t* out_off,
                                   const uint64_t* out_cap, uint64_t* out_len, int32_t* status) {
    if (!c || !s) return LZF_ERR_INVALID_ARG;
    if (nframes && (!in_off || !in_len || !out_off || !out_cap || !out_len || !status))
        return fail(c, LZF_ERR_INVALID_ARG, "null pointer");
    LZF_CU(c, cudaSetDevice(c->device));
    uint64_t target = c->compress_chunk_bytes;
    if (target == 0) {
        target = (uint64_t)c->num_sms * 28 * (s->block_size ? s->block_size : (4ull << 20));
        if (target < (256ull << 20)) target = 256ull << 20;
        if (target > (24ull << 30)) target = 24ull << 30;
    }
    const std::vector<uint32_t> chunks = plan_chunks(in_len, nframes, target);
    // wave-sized chunks hold three buffers of their own size each, and two block kernels never share an SM: two
    // slots are enough to overlap the copies of one chunk with the kernel of the next
    return run_chunks(c, (uint32_t)chunks.size() - 1, [&](uint32_t i, lzf_slot& sl) {
        return compress_chunk(c, sl, s, chunks[i], chunks[i + 1], in, in_off, in_len, out, out_off, out_cap, out_len, status);
    }, target >= (2ull << 30) ? 2u : (uint32_t)kSlots);
}

extern "C" int lzf_frame_compress(lzf_ctx* c, const lzf_settings* s, const uint8_t* in, size_t n,
                                  uint8_t* out, size_t cap, size_t* written, int32_t* status) {
    if (!c || !s || !written || !status) return LZF_ERR_INVALID_ARG;
    const uint64_t in_off = 0, in_len = n, out_off = 0, out_cap = cap;
    uint64_t out_len = 0;
    const int rc = lzf_frames_compress(c, s, in, &in_off, &in_len, 1, out, &out_off, &out_cap, &out_len, status);
    *written = (size_t)out_len;
    return rc;
}

extern "C" int lzf_frames_decompress_device(lzf_ctx* c, const uint8_t* d_in, const uint64_t* in_off, const uint64_t* in_len,
                                            uint32_t nframes, uint8_t* d_out, const uint64_t* out_off,
                                            const uint64_t* out_cap, uint64_t* out_len, int32_t* status, int32_t* detail) {
    if (!c) return LZF_ERR_INVALID_ARG;
    LZF_CU(c, cudaSetDevice(c->device));
    DecodeOut ex{nullptr, detail, nullptr, nullptr};
    return frames_decompress_core(c, d_in, in_off, in_len, nframes, d_out, out_off, out_cap, out_len, status, ex, cur_slot(c)->stream);
}

namespace {

int decompress_chunk(lzf_ctx* c, lzf_slot& sl, uint32_t f0, uint32_t f1, const uint8_t* in, const uint64_t* in_off,
                     const uint64_t* in_len, uint8_t* out, const uint64_t* out_off, const uint64_t* out_cap,
                     uint64_t* out_len, int32_t* status, int32_t* detail, uint64_t* consumed,
                     const uint8_t* dict, uint64_t dlen) {
    const uint32_t n = f1 - f0;
    std::vector<uint64_t> dcap(n);
    for (uint32_t f = 0; f < n; f++) {
        // a frame of C bytes decodes to at most ~C/5 blocks of <= 4 MiB; the device copy of the output
        // is bounded by what the caller can take anyway
        const uint64_t worst = (in_len[f0 + f] / 5 + 1) * (4ull << 20);
        dcap[f] = out_cap[f0 + f] < worst ? out_cap[f0 + f] : worst;
    }
    const HostLayout li = plan_layout(in_off + f0, in_len + f0, n);
    const HostLayout lo = plan_layout(out_off + f0, dcap.data(), n);
    int rc;
    if ((rc = ensure_dev(c, sl.d_io_in, li.span + 256))) return rc;
    if ((rc = ensure_dev(c, sl.d_io_out, lo.span + 256))) return rc;
    uint8_t* din = (uint8_t*)sl.d_io_in.p;
    const uint8_t* dout = (const uint8_t*)sl.d_io_out.p;
    if (li.dense) {
        if (li.span) LZF_CU(c, cudaMemcpyAsync(din, in + li.base, li.span, cudaMemcpyHostToDevice, sl.stream));
    } else {
        for (uint32_t f = 0; f < n; f++)
            if (in_len[f0 + f])
                LZF_CU(c, cudaMemcpyAsync(din + li.dev_off[f], in + in_off[f0 + f], in_len[f0 + f], cudaMemcpyHostToDevice, sl.stream));
    }
    if (dlen) {
        if ((rc = ensure_dev(c, sl.d_dict, dlen + 16))) return rc;
        LZF_CU(c, cudaMemcpyAsync(sl.d_dict.p, dict, dlen, cudaMemcpyHostToDevice, sl.stream));
    }
    // The plaintext of a dense run of frames travels back as ONE copy that starts as soon as the block kernel
    // has run, on the slot's side stream, while this thread resolves the per-frame outcome (and the content
    // checksums run) on the main stream.  Whatever the resolution then finds irregular (short frames, errors,
    // re-decoded frames) is copied again, frame by frame, after both streams have drained.
    bool early_copy = false;
    std::vector<uint8_t> moved(n, 0);
    DecodeOut ex{consumed ? consumed + f0 : nullptr, detail ? detail + f0 : nullptr, nullptr, moved.data()};
    if (lo.dense && lo.span) {
        ex.after_decode = [&](uint64_t expect) -> int {
            if (lo.span > expect + expect / 4) return LZF_SUCCESS;       // generous capacities: copy what was decoded, later
            LZF_CU(c, cudaEventRecord(sl.ev_fork, sl.stream));
            LZF_CU(c, cudaStreamWaitEvent(sl.side, sl.ev_fork, 0));
            LZF_CU(c, cudaMemcpyAsync(out + lo.base, dout, lo.span, cudaMemcpyDeviceToHost, sl.side));
            early_copy = true;
            return LZF_SUCCESS;
        };
    }
    rc = frames_decompress_core(c, din, li.dev_off.data(), in_len + f0, n, (uint8_t*)sl.d_io_out.p, lo.dev_off.data(),
                                dcap.data(), out_len + f0, status + f0, ex, sl.stream, dlen ? (const uint8_t*)sl.d_dict.p : nullptr, dlen);
    if (early_copy) LZF_CU(c, cudaStreamSynchronize(sl.side));
    if (rc) return rc;
    bool full = lo.dense;      // every frame filled its capacity exactly: one copy moves the whole run
    for (uint32_t f = f0; f < f1 && full; f++) full = out_len[f] == dcap[f - f0];
    if (full) {
        if (lo.span && !early_copy) LZF_CU(c, cudaMemcpyAsync(out + lo.base, dout, lo.span, cudaMemcpyDeviceToHost, sl.stream));
        // the early copy left while frames with short non-final blocks still sat at their pass-1 positions: those
        // frames travel again, now that the side stream has drained and the exact re-decode is queued on the main one
        for (uint32_t f = f0; f < f1 && early_copy; f++)
            if (moved[f - f0] && out_len[f])
                LZF_CU(c, cudaMemcpyAsync(out + out_off[f], dout + lo.dev_off[f - f0], out_len[f], cudaMemcpyDeviceToHost, sl.stream));
    } else {
        for (uint32_t f = f0; f < f1; f++)
            if (out_len[f])
                LZF_CU(c, cudaMemcpyAsync(out + out_off[f], dout + lo.dev_off[f - f0], out_len[f], cudaMemcpyDeviceToHost, sl.stream));
    }
    LZF_CU(c, cudaStreamSynchronize(sl.stream));
    return LZF_SUCCESS;
}

int frames_decompress_host(lzf_ctx* c, const uint8_t* in, const uint64_t* in_off, const uint64_t* in_len,
                           uint32_t nframes, uint8_t* out, const uint64_t* out_off, const uint64_t* out_cap,
                           uint64_t* out_len, int32_t* status, int32_t* detail, uint64_t* consumed,
                           const uint8_t* dict = nullptr, uint64_t dlen = 0) {
    if (nframes && (!in_off || !in_len || !out_off || !out_cap || !out_len || !status))
        return fail(c, LZF_ERR_INVALID_ARG, "null pointer");
    LZF_CU(c, cudaSetDevice(c->device));
    // chunk by compressed + plaintext bytes
    std::vector<uint64_t> weight(nframes);
    for (uint32_t f = 0; f < nframes; f++) {
        const uint64_t worst = (in_len[f] / 5 + 1) * (4ull << 20);
        weight[f] = in_len[f] + (out_cap[f] < worst ? out_cap[f] : worst);
    }
    const std::vector<uint32_t> chunks = plan_chunks(weight.data(), nframes, c->chunk_bytes, true);
    return run_chunks(c, (uint32_t)chunks.size() - 1, [&](uint32_t i, lzf_slot& sl) {
        return decompress_chunk(c, sl, chunks[i], chunks[i + 1], in, in_off, in_len, out, out_off, out_cap, out_len, status,
                                detail, consumed, dict, dlen);
    });
}
}  // namespace

extern "C" int lzf_frames_decompress(lzf_ctx* c, const uint8_t* in, const uint64_t* in_off, const uint64_t* in_len,
                                     uint32_t nframes, uint8_t* out, const uint64_t* out_off, const uint64_t* out_cap,
                                     uint64_t* out_len, int32_t* status, int32_t* detail) {
    if (!c) return LZF_ERR_INVALID_ARG;
    return frames_decompress_host(c, in, in_off, in_len, nframes, out, out_off, out_cap, out_len, status, detail, nullptr);
}

extern "C" int lzf_frame_decompress(lzf_ctx* c, const uint8_t* in, size_t n, const uint8_t* dict, size_t dlen,
                                    uint8_t* out, size_t cap, size_t* written, size_t* consumed,
                                    int32_t* status, int32_t* detail) {
    if (!c || !written || !status) return LZF_ERR_INVALID_ARG;
    const uint64_t in_off = 0, in_len = n, out_off = 0, out_cap = cap;
    uint64_t out_len = 0, cons = 0;
    int32_t det = 0;
    const int rc = frames_decompress_host(c, in, &in_off, &in_len, 1, out, &out_off, &out_cap, &out_len, status, &det, &cons,
                                          dlen ? dict : nullptr, dlen);
    *written = (size_t)out_len;
    if (consumed) *consumed = (size_t)cons;
    if (detail) *detail = det;
    return rc;
}
