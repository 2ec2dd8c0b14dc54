/*
 * lzf_oracle.c — TEST INFRASTRUCTURE ONLY (see lzf_oracle.h).
 *
 * Plain-C restatement of the lz-fear raw block codec, XXH32 and frame glue.
 * Written from the behaviour of the reference (file:line cited per function,
 * relative to /root/reference); it is a checker for the CUDA path and the CPU
 * baseline ("port") in bench.py — never part of the shipped library.
 */
#include "lzf_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* little-endian loads (byteorder crate: LE / NativeEndian on a LE host)     */
/* ------------------------------------------------------------------------ */
static inline uint16_t ld16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
static inline uint32_t ld32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t ld64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline void st32(uint8_t* p, uint32_t v) { memcpy(p, &v, 4); }
static inline void st64(uint8_t* p, uint64_t v) { memcpy(p, &v, 8); }

/* ------------------------------------------------------------------------ */
/* XXH32 — public algorithm; call sites src/framed/compress.rs:172,197-199,  */
/* 233-235,260-262,279-281 and src/framed/decompress.rs:112-133,207-211,     */
/* 229-234,276-278 (twox-hash XxHash32::with_seed(0), write*, finish)        */
/* ------------------------------------------------------------------------ */
#define P1 2654435761u
#define P2 2246822519u
#define P3 3266489917u
#define P4 668265263u
#define P5 374761393u
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static inline uint32_t xxh_round(uint32_t acc, uint32_t x) { return rotl32(acc + x * P2, 13) * P1; }

void lzfo_xxh32_init(lzfo_xxh32_state* s, uint32_t seed) {
    s->acc[0] = seed + P1 + P2;
    s->acc[1] = seed + P2;
    s->acc[2] = seed;
    s->acc[3] = seed - P1;
    s->buflen = 0;
    s->total = 0;
    s->seed = seed;
}

void lzfo_xxh32_update(lzfo_xxh32_state* s, const void* data, size_t n) {
    const uint8_t* p = (const uint8_t*)data;
    s->total += n;
    if (s->buflen) {
        size_t take = 16 - s->buflen;
        if (take > n) take = n;
        memcpy(s->buf + s->buflen, p, take);
        s->buflen += (uint32_t)take;
        p += take;
        n -= take;
        if (s->buflen < 16) return;
        for (int i = 0; i < 4; i++) s->acc[i] = xxh_round(s->acc[i], ld32(s->buf + 4 * i));
        s->buflen = 0;
    }
    uint32_t a0 = s->acc[0], a1 = s->acc[1], a2 = s->acc[2], a3 = s->acc[3];
    while (n >= 16) {
        a0 = xxh_round(a0, ld32(p));
        a1 = xxh_round(a1, ld32(p + 4));
        a2 = xxh_round(a2, ld32(p + 8));
        a3 = xxh_round(a3, ld32(p + 12));
        p += 16;
        n -= 16;
    }
    s->acc[0] = a0; s->acc[1] = a1; s->acc[2] = a2; s->acc[3] = a3;
    if (n) {
        memcpy(s->buf, p, n);
        s->buflen = (uint32_t)n;
    }
}

uint32_t lzfo_xxh32_finish(const lzfo_xxh32_state* s) {
    uint32_t h;
    if (s->total >= 16)
        h = rotl32(s->acc[0], 1) + rotl32(s->acc[1], 7) + rotl32(s->acc[2], 12) + rotl32(s->acc[3], 18);
    else
        h = s->seed + P5;
    h += (uint32_t)s->total;
    const uint8_t* p = s->buf;
    uint32_t n = s->buflen;
    while (n >= 4) { h = rotl32(h + ld32(p) * P3, 17) * P4; p += 4; n -= 4; }
    while (n) { h = rotl32(h + (*p) * P5, 11) * P1; p++; n--; }
    h ^= h >> 15; h *= P2;
    h ^= h >> 13; h *= P3;
    h ^= h >> 16;
    return h;
}

uint32_t lzfo_xxh32(const void* data, size_t n, uint32_t seed) {
    lzfo_xxh32_state s;
    lzfo_xxh32_init(&s, seed);
    lzfo_xxh32_update(&s, data, n);
    return lzfo_xxh32_finish(&s);
}

/* ------------------------------------------------------------------------ */
/* EncoderTable, U32Table, U16Table — src/raw/compress/mod.rs:14-101         */
/* ------------------------------------------------------------------------ */
struct lzfo_table {
    int kind;
    unsigned hashlog;
    size_t nslots;
    size_t offset;   /* :30, :81 */
    uint32_t* d32;
    uint16_t* d16;
};

lzfo_table* lzfo_table_new(int kind, unsigned hashlog) {
    lzfo_table* t = (lzfo_table*)calloc(1, sizeof(*t));
    if (!t) return NULL;
    t->kind = kind;
    t->hashlog = hashlog ? hashlog : 12;                      /* :15 HASHLOG = 12 */
    if (kind == LZFO_TABLE_U32) {
        t->nslots = (size_t)1 << t->hashlog;                  /* :14,:29 */
        t->d32 = (uint32_t*)calloc(t->nslots, sizeof(uint32_t));   /* :34 zeroed */
    } else {
        t->nslots = (size_t)2 << t->hashlog;                  /* :80 DICTIONARY_SIZE*2 */
        t->d16 = (uint16_t*)calloc(t->nslots, sizeof(uint16_t));   /* :85 zeroed */
    }
    return t;
}

lzfo_table* lzfo_table_clone(const lzfo_table* t) {
    lzfo_table* c = lzfo_table_new(t->kind, t->hashlog);
    c->offset = t->offset;
    if (t->d32) memcpy(c->d32, t->d32, t->nslots * sizeof(uint32_t));
    if (t->d16) memcpy(c->d16, t->d16, t->nslots * sizeof(uint16_t));
    return c;
}

static void table_assign(lzfo_table* dst, const lzfo_table* src) {
    dst->offset = src->offset;
    if (src->d32) memcpy(dst->d32, src->d32, src->nslots * sizeof(uint32_t));
    if (src->d16) memcpy(dst->d16, src->d16, src->nslots * sizeof(uint16_t));
}

void lzfo_table_free(lzfo_table* t) {
    if (!t) return;
    free(t->d32);
    free(t->d16);
    free(t);
}

size_t lzfo_table_payload_size_limit(const lzfo_table* t) {
    return t->kind == LZFO_TABLE_U32 ? (size_t)UINT32_MAX : (size_t)UINT16_MAX;   /* :75, :100 */
}

/* hash_for_u32, 64-bit little-endian branch — :40-51 */
static inline size_t hash_for_u32(const uint8_t* input, size_t avail, unsigned hashlog) {
    uint64_t v = avail >= 8 ? ld64(input) : 0;                /* :43 get(..8) ... unwrap_or(0) */
    return (size_t)(((v << 24) * 889523592379ull) >> (64 - hashlog));   /* :48,:50 */
}
/* hash_for_u16 — :58-61.  NativeEndian::read_u32 panics with < 4 bytes. */
static inline int hash_for_u16(const uint8_t* input, size_t avail, unsigned hashlog, size_t* h) {
    if (avail < 4) return LZFO_PANIC;
    uint32_t v = ld32(input);
    *h = (size_t)((v * 2654435761u) >> (32 - hashlog - 1));
    return LZFO_OK;
}

/* EncoderTable::replace — :64-71 (U32), :89-96 (U16) */
int lzfo_table_replace(lzfo_table* t, const uint8_t* input, size_t len, size_t pos, size_t* old) {
    size_t o = pos + t->offset;                               /* :65 */
    if (pos > len) return LZFO_PANIC;                         /* &input[offset..] out of range */
    size_t value;
    if (t->kind == LZFO_TABLE_U32) {
        if (o > UINT32_MAX) return LZFO_PANIC;                /* :67 expect */
        size_t h = hash_for_u32(input + pos, len - pos, t->hashlog);
        value = t->d32[h];                                    /* :68 mem::swap */
        t->d32[h] = (uint32_t)o;
    } else {
        if (o > UINT16_MAX) return LZFO_PANIC;                /* :92 expect */
        size_t h;
        if (hash_for_u16(input + pos, len - pos, t->hashlog, &h)) return LZFO_PANIC;
        value = t->d16[h];                                    /* :93 */
        t->d16[h] = (uint16_t)o;
    }
    *old = value > t->offset ? value - t->offset : 0;         /* :70 saturating_sub */
    return LZFO_OK;
}

void lzfo_table_offset(lzfo_table* t, size_t by) { t->offset += by; }   /* :72-74 */

/* ------------------------------------------------------------------------ */
/* NoPartialWrites — src/framed/compress.rs:294-308                          */
/* ------------------------------------------------------------------------ */
typedef struct { uint8_t* p; size_t remaining; } bounded_writer;
static inline int bw_write(bounded_writer* w, const void* data, size_t n) {
    if (w->remaining < n) return LZFO_WRITER_FULL;            /* :298-301 */
    memcpy(w->p, data, n);
    w->p += n;
    w->remaining -= n;
    return LZFO_OK;
}
static inline int bw_u8(bounded_writer* w, uint8_t b) { return bw_write(w, &b, 1); }

/* write_lsic_head — src/raw/compress/mod.rs:239-242 */
static inline void write_lsic_head(uint8_t* token, unsigned shift, size_t value) {
    uint8_t i = (uint8_t)(value < 0xF ? value : 0xF);
    *token |= (uint8_t)(i << shift);
}
/* write_lsic_tail — :243-260 */
static int write_lsic_tail(bounded_writer* w, size_t value) {
    if (value < 0xF) return LZFO_OK;
    value -= 0xF;
    while (value >= 4 * 0xFF) {                               /* :251-254 */
        uint32_t m = UINT32_MAX;
        int rc = bw_write(w, &m, 4);
        if (rc) return rc;
        value -= 4 * 0xFF;
    }
    while (value >= 0xFF) {                                   /* :255-258 */
        int rc = bw_u8(w, 0xFF);
        if (rc) return rc;
        value -= 0xFF;
    }
    return bw_u8(w, (uint8_t)value);                          /* :259 */
}

/* count_matching_bytes — :117-145 */
static size_t count_matching_bytes(const uint8_t* a, size_t alen, const uint8_t* b, size_t blen) {
    size_t n = alen < blen ? alen : blen;
    size_t m = 0;
    while (m + 8 <= n) {                                      /* chunks_exact(REGSIZE).zip */
        uint64_t x = ld64(a + m) ^ ld64(b + m);
        if (x == 0) {
            m += 8;
        } else {
            return m + (size_t)(__builtin_ctzll(x) / 8);      /* :136 */
        }
    }
    while (m < n && a[m] == b[m]) m++;                        /* :143 */
    return m;
}

size_t lzfo_compress_bound(size_t n) { return n + n / 255 + 16; }

/* compress2 — :165-238 ; write_group — :150-163 */
int lzfo_compress2(const uint8_t* input, size_t len, size_t cursor_in, lzfo_table* table,
                   uint8_t* out, size_t cap, size_t* written) {
    bounded_writer w = { out, cap };
    int rc;
    *written = 0;
    if (len > lzfo_table_payload_size_limit(table)) return LZFO_PANIC;      /* :167 assert */

    const size_t init_cursor = cursor_in;
    size_t cursor = cursor_in;
    while (cursor < len) {                                                  /* :171 */
        const size_t literal_start = cursor;
        size_t step_counter = (size_t)1 << 6;                               /* :174 ACCELERATION << SKIP_TRIGGER */
        size_t step = 1;
        size_t dup_offset = 0, dup_extra = 0;
        for (;;) {                                                          /* :177 */
            size_t remaining = len > cursor ? len - cursor : 0;             /* saturating_sub */
            if (remaining < 12) {                                           /* :178-190 */
                size_t literal_len = len - literal_start;
                uint8_t token = 0;
                write_lsic_head(&token, 4, literal_len);
                if ((rc = bw_u8(&w, token))) return rc;
                if ((rc = write_lsic_tail(&w, literal_len))) return rc;
                if ((rc = bw_write(&w, input + literal_start, literal_len))) return rc;
                *written = cap - w.remaining;
                return LZFO_OK;
            }
            const size_t batch_len = len - 5 - cursor;                      /* :195 */
            size_t candidate;
            if ((rc = lzfo_table_replace(table, input, len, cursor, &candidate))) return rc;   /* :196 */

            if (cursor != init_cursor && cursor - candidate <= 0xFFFF) {    /* :200-201 */
                size_t matching = count_matching_bytes(input + cursor, batch_len,
                                                       input + candidate, len - candidate);   /* :203-204 */
                if (matching >= 4) {                                        /* :206 checked_sub(MINMATCH) */
                    size_t extra = matching - 4;
                    dup_offset = cursor - candidate;                        /* :208 */
                    size_t max_backtrack = cursor - literal_start;          /* :211 */
                    size_t backtrack = 0;                                   /* :212 */
                    while (backtrack < max_backtrack && backtrack < candidate &&
                           input[cursor - 1 - backtrack] == input[candidate - 1 - backtrack])
                        backtrack++;
                    extra += backtrack;                                     /* :214 */
                    cursor += matching;                                     /* :215 */
                    size_t dummy;
                    if ((rc = lzfo_table_replace(table, input, len, cursor - 2, &dummy))) return rc;   /* :218 */
                    dup_extra = extra;
                    break;                                                  /* :220 */
                }
            }
            cursor += step;                                                 /* :225 */
            step = step_counter >> 6;                                       /* :226 */
            if (literal_start + 1 != cursor) step_counter += 1;             /* :229-231 */
        }
        /* :235-236 + write_group :150-163 */
        size_t literal_end = cursor - dup_extra - 4;
        size_t literal_len = literal_end - literal_start;
        uint8_t token = 0;
        write_lsic_head(&token, 4, literal_len);
        write_lsic_head(&token, 0, dup_extra);
        if ((rc = bw_u8(&w, token))) return rc;
        if ((rc = write_lsic_tail(&w, literal_len))) return rc;
        if ((rc = bw_write(&w, input + literal_start, literal_len))) return rc;
        uint16_t off16 = (uint16_t)dup_offset;
        if ((rc = bw_write(&w, &off16, 2))) return rc;
        if ((rc = write_lsic_tail(&w, dup_extra))) return rc;
    }
    *written = cap - w.remaining;
    return LZFO_OK;
}

int lzfo_compress_block(const uint8_t* input, size_t len, int table_kind, unsigned hashlog,
                        uint8_t* out, size_t cap, size_t* written) {
    lzfo_table* t = lzfo_table_new(table_kind, hashlog);
    if (!t) return LZFO_PANIC;
    int rc = lzfo_compress2(input, len, 0, t, out, cap, written);
    lzfo_table_free(t);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* decompress_raw — src/raw/decompress.rs:30-43, 58-78, 80-138                */
/* ------------------------------------------------------------------------ */
typedef struct { const uint8_t* in; size_t n; size_t pos; } rd;

/* read_lsic — :30-43 */
static int read_lsic(rd* r, uint8_t initial, size_t* value) {
    size_t v = initial;
    if (v == 0xF) {
        for (;;) {
            if (r->pos >= r->n) return LZFO_UNEXPECTED_END;   /* :35 read_u8()? */
            uint8_t more = r->in[r->pos++];
            v += more;
            if (more != 0xFF) break;
        }
    }
    *value = v;
    return LZFO_OK;
}

int lzfo_decompress_raw(const uint8_t* in, size_t n, const uint8_t* prefix, size_t plen,
                        uint8_t* out, size_t out_cap, size_t out_limit, size_t* out_len) {
    rd r = { in, n, 0 };
    size_t olen = *out_len;          /* output.len(): pre-existing bytes are addressable history */
    int rc;
    int status = LZFO_OK;
    /* Once the physical cap is exceeded we stop writing but keep parsing, so that a codec error
     * later in the block is still reported exactly as the reference (whose Vec just grows) would. */
    int dry = olen > out_cap;
    while (r.pos < r.n) {                                     /* :61 while let Ok(token) */
        uint8_t token = r.in[r.pos++];
        size_t lit;
        if ((rc = read_lsic(&r, token >> 4, &lit))) { status = rc; break; }      /* :63 */
        if (r.n - r.pos < lit) { status = LZFO_UNEXPECTED_END; break; }          /* :67 read_exact */
        if (!dry && olen + lit > out_cap) dry = 1;
        if (!dry) memcpy(out + olen, r.in + r.pos, lit);
        olen += lit;
        r.pos += lit;

        if (r.n - r.pos < 2) {                                /* :70 if let Ok(offset) — else falls through; */
            r.pos = r.n;                                      /* recent std: failed read_exact leaves the cursor at EOF */
            continue;
        }
        size_t offset = ld16(r.in + r.pos);
        r.pos += 2;
        size_t mlen;
        if ((rc = read_lsic(&r, token & 0xF, &mlen))) { status = rc; break; }    /* :71 */
        mlen += 4;
        if (olen + mlen > out_limit) { status = LZFO_MEMORY_LIMIT_EXCEEDED; break; }   /* :72-74 */
        /* copy_overlapping :80-138 — every arm is observationally the sequential byte loop */
        if (offset == 0) { status = LZFO_ZERO_DEDUP_OFFSET; break; }             /* :83 */
        if (offset > olen && offset - olen > plen) { status = LZFO_INVALID_DEDUP_OFFSET; break; }   /* :84-89 */
        if (!dry && olen + mlen > out_cap) dry = 1;
        if (!dry) {
            size_t k = 0;
            if (offset > olen) {                              /* :84-99 prefix arm */
                size_t need = offset - olen;
                size_t take = need < mlen ? need : mlen;
                memcpy(out + olen, prefix + plen - need, take);
                k = take;
            }
            /* the reference's own fast paths (:100-136), so that the CPU baseline is not slower than the crate:
             * memset for offset 1, memcpy when the ranges do not overlap, a 16-byte pattern buffer for offsets
             * 2 / 4 / 8, single bytes otherwise.  Every arm writes the bytes of the sequential loop. */
            uint8_t* dst = out + olen + k;
            const size_t rest = mlen - k;
            if (rest == 0) {
            } else if (offset == 1) {                          /* :102 */
                memset(dst, dst[-1], rest);
            } else if (rest <= offset) {                       /* :104-111 */
                memcpy(dst, dst - offset, rest);
            } else if (offset == 2 || offset == 4 || offset == 8) {   /* :112-127 */
                uint8_t buf[16];
                for (size_t i = 0; i < 16; i += offset) memcpy(buf + i, dst - offset, offset);
                size_t i = 0;
                for (; i + 16 <= rest; i += 16) memcpy(dst + i, buf, 16);
                if (i < rest) memcpy(dst + i, buf, rest - i);
            } else {                                           /* :128-135 */
                for (size_t i = 0; i < rest; i++) dst[i] = dst[(ptrdiff_t)i - (ptrdiff_t)offset];
            }
        }
        olen += mlen;
    }
    *out_len = olen;
    if (status == LZFO_OK && olen > out_cap) status = LZFO_OUTPUT_CAP;
    return status;
}

/* ------------------------------------------------------------------------ */
/* frame header — src/framed/header.rs, src/framed/mod.rs:16-20              */
/* ------------------------------------------------------------------------ */
#define LZF_MAGIC 0x184D2204u
#define LZF_INCOMPRESSIBLE 0x80000000u
#define LZF_WINDOW_SIZE 65536u
#define FLAG_INDEPENDENT 0x20
#define FLAG_BLOCK_CHECKSUMS 0x10
#define FLAG_CONTENT_SIZE 0x08
#define FLAG_CONTENT_CHECKSUM 0x04
#define FLAG_DICTIONARY_ID 0x01

void lzfo_settings_default(lzfo_settings* s) {               /* src/framed/compress.rs:44-55 */
    memset(s, 0, sizeof(*s));
    s->independent_blocks = 1;
    s->block_checksums = 0;
    s->content_checksum = 1;
    s->block_size = 4u * 1024 * 1024;
    s->hashlog = 12;
}

/* BlockDescriptor::block_maxsize — header.rs:72-80 */
static int bd_block_maxsize(uint8_t bd, uint64_t* size) {
    unsigned s = (bd >> 4) & 7;
    if (s >= 4 && s < 8) { *size = (uint64_t)1 << (s * 2 + 8); return 0; }
    return LZFO_P_UNIMPLEMENTED_BLOCKSIZE;
}
/* BlockDescriptor::new — header.rs:53-62.  Returns 0 ok, 1 None, 2 panic (unwrap at :55). */
static int bd_new(uint64_t block_maxsize, uint8_t* bd) {
    unsigned tz = block_maxsize ? (unsigned)__builtin_ctzll(block_maxsize) : 64;
    unsigned maybe = ((tz > 8 ? tz - 8 : 0) / 2) & 0xFF;
    uint8_t b = (uint8_t)(maybe << 4);
    if (b & 0x8F) return 2;                                   /* parse(..).unwrap() */
    uint64_t sz;
    if (bd_block_maxsize(b, &sz) || sz != block_maxsize) return 1;
    *bd = b;
    return 0;
}

size_t lzfo_frame_bound(const lzfo_settings* s, size_t n) {
    size_t bs = (size_t)s->block_size ? (size_t)s->block_size : 1;
    size_t nblocks = (n + bs - 1) / bs;
    return 19 + n + nblocks * 8 + 8;
}

/* compress_internal — src/framed/compress.rs:159-282 */
int lzfo_frame_compress(const lzfo_settings* s, const uint8_t* in, size_t n,
                        uint8_t* out, size_t cap, size_t* written) {
    size_t o = 0;
    *written = 0;
    uint8_t flags = 0;
    if (s->independent_blocks) flags |= FLAG_INDEPENDENT;     /* :164-179 */
    if (s->block_checksums) flags |= FLAG_BLOCK_CHECKSUMS;
    if (s->content_checksum) flags |= FLAG_CONTENT_CHECKSUM;
    if (s->has_dictionary_id) flags |= FLAG_DICTIONARY_ID;
    if (s->has_content_size) flags |= FLAG_CONTENT_SIZE;
    uint8_t bd;
    int b = bd_new(s->block_size, &bd);                       /* :183 */
    if (b == 2) return LZFO_F_PANIC;
    if (b == 1) return LZFO_F_INVALID_BLOCK_SIZE;

    uint8_t header[19];
    size_t h = 0;
    st32(header, LZF_MAGIC); h = 4;                           /* :186-195 */
    header[h++] = (uint8_t)((1 << 6) | flags);
    header[h++] = bd;
    if (s->has_content_size) { st64(header + h, s->content_size); h += 8; }
    if (s->has_dictionary_id) { st32(header + h, s->dictionary_id); h += 4; }
    header[h] = (uint8_t)(lzfo_xxh32(header + 4, h - 4, 0) >> 8);   /* :197-199 */
    h++;
    if (cap - o < h) return LZFO_F_WRITE_ERROR;
    memcpy(out + o, header, h); o += h;                       /* :200 */

    const unsigned hashlog = s->hashlog ? s->hashlog : 12;
    const size_t block_size = (size_t)s->block_size;
    const uint8_t* dict = s->dictionary;
    const size_t dlen = dict ? (size_t)s->dictionary_len : 0;

    lzfo_table* template_table = lzfo_table_new(LZFO_TABLE_U32, hashlog);   /* :202 */
    int rc = LZFO_F_OK;
    if (dict) {                                               /* :204-214 windows(8).step_by(3) */
        for (size_t off = 0; off + 8 <= dlen; off += 3) {
            size_t dummy;
            if (lzfo_table_replace(template_table, dict, dlen, off, &dummy)) { rc = LZFO_F_PANIC; break; }
        }
    }
    /* in_buffer: reserve the most it can ever hold */
    size_t inbuf_cap = (dlen > LZF_WINDOW_SIZE ? dlen : LZF_WINDOW_SIZE) + block_size + 16;
    uint8_t* in_buffer = (uint8_t*)malloc(inbuf_cap);
    uint8_t* out_buffer = (uint8_t*)malloc(block_size ? block_size : 1);    /* :219 */
    size_t in_len = 0;
    if (dlen) { memcpy(in_buffer, dict, dlen); in_len = dlen; }             /* :218 */
    lzfo_table* table = lzfo_table_clone(template_table);                    /* :220 */
    lzfo_xxh32_state content_hasher;
    lzfo_xxh32_init(&content_hasher, 0);                                     /* :172 */

    size_t ipos = 0;
    while (rc == LZFO_F_OK) {                                                /* :221 */
        size_t window_offset = in_len;                                       /* :222 */
        size_t read_bytes = n - ipos < block_size ? n - ipos : block_size;   /* :227 */
        memcpy(in_buffer + in_len, in + ipos, read_bytes);
        in_len += read_bytes; ipos += read_bytes;
        if (read_bytes == 0) break;                                          /* :229-231 */
        if (s->content_checksum) lzfo_xxh32_update(&content_hasher, in_buffer + window_offset, read_bytes);  /* :233-235 */

        size_t wlen = 0;
        int crc = lzfo_compress2(in_buffer, in_len, window_offset, table, out_buffer, read_bytes, &wlen);   /* :242-243 */
        const uint8_t* payload;
        size_t payload_len;
        uint32_t word;
        if (crc == LZFO_OK) {                                                /* :244-249 */
            word = (uint32_t)wlen; payload = out_buffer; payload_len = wlen;
        } else if (crc == LZFO_WRITER_FULL) {                                /* :250-255 */
            word = (uint32_t)read_bytes | LZF_INCOMPRESSIBLE;
            payload = in_buffer + window_offset; payload_len = read_bytes;
        } else { rc = LZFO_F_PANIC; break; }
        size_t need = 4 + payload_len + (s->block_checksums ? 4 : 0);
        if (cap - o < need) { rc = LZFO_F_WRITE_ERROR; break; }
        st32(out + o, word); o += 4;
        memcpy(out + o, payload, payload_len); o += payload_len;             /* :258 */
        if (s->block_checksums) { st32(out + o, lzfo_xxh32(payload, payload_len, 0)); o += 4; }   /* :259-263 */

        if (s->independent_blocks) {                                         /* :265-270 */
            in_len = 0;
            if (dlen) { memcpy(in_buffer, dict, dlen); in_len = dlen; }
            table_assign(table, template_table);
        } else if (in_len > LZF_WINDOW_SIZE) {                               /* :271-275 */
            size_t forget = in_len - LZF_WINDOW_SIZE;
            lzfo_table_offset(table, forget);
            memmove(in_buffer, in_buffer + forget, LZF_WINDOW_SIZE);
            in_len = LZF_WINDOW_SIZE;
        }
    }
    if (rc == LZFO_F_OK) {
        size_t need = 4 + (s->content_checksum ? 4 : 0);
        if (cap - o < need) rc = LZFO_F_WRITE_ERROR;
        else {
            st32(out + o, 0); o += 4;                                        /* :277 EndMark */
            if (s->content_checksum) { st32(out + o, lzfo_xxh32_finish(&content_hasher)); o += 4; }   /* :279-281 */
        }
    }
    free(in_buffer);
    free(out_buffer);
    lzfo_table_free(table);
    lzfo_table_free(template_table);
    *written = o;
    return rc;
}

/* LZ4FrameReader::new — src/framed/decompress.rs:101-161 */
int lzfo_frame_parse_header(const uint8_t* in, size_t n, lzfo_frame_info* info, int* detail) {
    size_t p = 0;
    if (detail) *detail = 0;
    memset(info, 0, sizeof(*info));
    if (n - p < 4) return LZFO_F_INPUT_ERROR;                 /* :103 */
    uint32_t magic = ld32(in); p += 4;
    if (magic != LZF_MAGIC) return LZFO_F_WRONG_MAGIC;        /* :104-106 */
    if (n - p < 1) return LZFO_F_INPUT_ERROR;
    uint8_t flags_byte = in[p++];                             /* :108 */
    /* Flags::parse — header.rs:31-42 */
    if ((flags_byte >> 6) != 1) { if (detail) *detail = LZFO_P_UNSUPPORTED_VERSION; return LZFO_F_HEADER_PARSE_ERROR; }
    if (flags_byte & 2) { if (detail) *detail = LZFO_P_RESERVED_FLAG_BITS; return LZFO_F_HEADER_PARSE_ERROR; }
    uint8_t flags = flags_byte & (FLAG_INDEPENDENT | FLAG_BLOCK_CHECKSUMS | FLAG_CONTENT_SIZE | FLAG_CONTENT_CHECKSUM | FLAG_DICTIONARY_ID);
    if (n - p < 1) return LZFO_F_INPUT_ERROR;
    uint8_t bd = in[p++];                                     /* :110 */
    if (bd & 0x8F) { if (detail) *detail = LZFO_P_RESERVED_BD_BITS; return LZFO_F_HEADER_PARSE_ERROR; }   /* header.rs:65-68 */
    lzfo_xxh32_state hs;
    lzfo_xxh32_init(&hs, 0);                                  /* :112-114 */
    lzfo_xxh32_update(&hs, &flags_byte, 1);
    lzfo_xxh32_update(&hs, &bd, 1);
    if (flags & FLAG_CONTENT_SIZE) {                          /* :116-122 */
        if (n - p < 8) return LZFO_F_INPUT_ERROR;
        info->content_size = ld64(in + p);
        info->has_content_size = 1;
        lzfo_xxh32_update(&hs, in + p, 8);
        p += 8;
    }
    if (flags & FLAG_DICTIONARY_ID) {                         /* :124-130 */
        if (n - p < 4) return LZFO_F_INPUT_ERROR;
        info->dictionary_id = ld32(in + p);
        info->has_dictionary_id = 1;
        lzfo_xxh32_update(&hs, in + p, 4);
        p += 4;
    }
    if (n - p < 1) return LZFO_F_INPUT_ERROR;
    uint8_t want = in[p++];                                   /* :132 */
    uint8_t have = (uint8_t)(lzfo_xxh32_finish(&hs) >> 8);
    if (want != have) return LZFO_F_HEADER_CHECKSUM_FAIL;     /* :134-136 */
    uint64_t bms;
    int pe = bd_block_maxsize(bd, &bms);                      /* :153 */
    if (pe) { if (detail) *detail = pe; return LZFO_F_HEADER_PARSE_ERROR; }
    info->flags = flags;
    info->block_maxsize = bms;
    info->header_len = p;
    return LZFO_F_OK;
}

/* decode_block loop — src/framed/decompress.rs:197-279, driven like decompress_frame :283-288 */
int lzfo_frame_decompress(const uint8_t* in, size_t n, const uint8_t* dict, size_t dlen,
                          uint8_t* out, size_t cap, size_t* written, size_t* consumed, int* detail) {
    lzfo_frame_info info;
    int dummy_detail;
    if (!detail) detail = &dummy_detail;
    *written = 0;
    if (consumed) *consumed = 0;
    int rc = lzfo_frame_parse_header(in, n, &info, detail);
    if (rc) return rc;
    size_t p = info.header_len;
    size_t o = 0;
    const size_t bms = (size_t)info.block_maxsize;
    const int dependent = !(info.flags & FLAG_INDEPENDENT);
    lzfo_xxh32_state ch;
    lzfo_xxh32_init(&ch, 0);                                  /* :138-142 */
    uint8_t* window = NULL;                                   /* carryover_window :144-148 */
    size_t wlen = 0;
    size_t wcap = (dlen > LZF_WINDOW_SIZE ? dlen : LZF_WINDOW_SIZE) + LZF_WINDOW_SIZE;
    if (dependent) window = (uint8_t*)malloc(wcap);
    /* a block may transiently decode to block_maxsize + C bytes before the :272 check */
    uint8_t* blk = (uint8_t*)malloc(2 * bms + 16);
    rc = LZFO_F_OK;
    for (;;) {
        if (n - p < 4) { rc = LZFO_F_INPUT_ERROR; break; }    /* :205 */
        uint32_t block_length = ld32(in + p); p += 4;
        if (block_length == 0) {                              /* :206-215 */
            if (info.flags & FLAG_CONTENT_CHECKSUM) {
                if (n - p < 4) { rc = LZFO_F_INPUT_ERROR; break; }
                uint32_t checksum = ld32(in + p); p += 4;
                if (lzfo_xxh32_finish(&ch) != checksum) { rc = LZFO_F_FRAME_CHECKSUM_FAIL; break; }
            }
            break;
        }
        int is_compressed = (block_length & LZF_INCOMPRESSIBLE) == 0;   /* :217-218 */
        block_length &= ~LZF_INCOMPRESSIBLE;
        if (block_length > (uint32_t)bms) { rc = LZFO_F_BLOCK_SIZE_OVERFLOW; break; }   /* :220-222 */
        if (n - p < block_length) { rc = LZFO_F_INPUT_ERROR; break; }   /* :226 read_exact */
        const uint8_t* buf = in + p; p += block_length;
        if (info.flags & FLAG_BLOCK_CHECKSUMS) {              /* :228-235 */
            if (n - p < 4) { rc = LZFO_F_INPUT_ERROR; break; }
            uint32_t checksum = ld32(in + p); p += 4;
            if (lzfo_xxh32(buf, block_length, 0) != checksum) { rc = LZFO_F_BLOCK_CHECKSUM_FAIL; break; }
        }
        const uint8_t* dec_prefix = dict;                     /* :238-245 */
        size_t dec_plen = dlen;
        if (dependent) {
            if (wlen == 0 && dlen) { memcpy(window, dict, dlen); wlen = dlen; }
            dec_prefix = window; dec_plen = wlen;
        }
        size_t outlen = 0;
        if (is_compressed) {                                  /* :247-248 */
            int st = lzfo_decompress_raw(buf, block_length, dec_prefix, dec_plen, blk, 2 * bms + 16, bms, &outlen);
            if (st != LZFO_OK) { *detail = st; rc = LZFO_F_CODEC_ERROR; break; }
        } else {                                              /* :249-251 */
            memcpy(blk, buf, block_length);
            outlen = block_length;
        }
        if (dependent) {                                      /* :253-269 */
            if (outlen < LZF_WINDOW_SIZE) {
                size_t available = wlen + outlen;
                if (available > LZF_WINDOW_SIZE) {
                    size_t surplus = available - LZF_WINDOW_SIZE;
                    memmove(window, window + surplus, wlen - surplus);
                    wlen -= surplus;
                }
                memcpy(window + wlen, blk, outlen); wlen += outlen;
            } else {
                memcpy(window, blk + outlen - LZF_WINDOW_SIZE, LZF_WINDOW_SIZE);
                wlen = LZF_WINDOW_SIZE;
            }
        }
        if (outlen > bms) { rc = LZFO_F_BLOCK_SIZE_OVERFLOW; break; }   /* :272-274 */
        if (info.flags & FLAG_CONTENT_CHECKSUM) lzfo_xxh32_update(&ch, blk, outlen);   /* :276-278 */
        if (cap - o < outlen) { rc = LZFO_F_WRITE_ERROR; break; }
        memcpy(out + o, blk, outlen); o += outlen;
        /* LZ4FrameIoReader::read (:54-61) hands an empty fill_buf() straight to the caller as
         * Ok(0), which read_to_end (:286) takes for end-of-stream: a block that decodes to zero
         * bytes ends decompress_frame early and successfully, with the rest of the frame unread. */
        if (outlen == 0) break;
    }
    free(window);
    free(blk);
    *written = o;
    if (consumed) *consumed = p;
    return rc;
}

/* ------------------------------------------------------------------------ */
/* multi-threaded batch drivers (bench.py cpu_baseline / --impl reference)   */
/* ------------------------------------------------------------------------ */
typedef struct {
    int compress;
    const uint8_t* in; const uint64_t* in_off; const uint32_t* in_len; uint32_t nblocks; unsigned hashlog;
    uint8_t* out; const uint64_t* out_off; const uint32_t* out_cap; const uint32_t* out_limit;
    uint32_t* out_len; int32_t* status;
    volatile uint32_t next;
} mt_job;

static void* mt_worker(void* arg) {
    mt_job* j = (mt_job*)arg;
    lzfo_table* t = j->compress ? lzfo_table_new(LZFO_TABLE_U32, j->hashlog) : NULL;
    for (;;) {
        uint32_t b = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (b >= j->nblocks) break;
        if (j->compress) {
            memset(t->d32, 0, t->nslots * sizeof(uint32_t));
            t->offset = 0;
            size_t w = 0;
            /* output capacity = the block's own plaintext length (src/framed/compress.rs:242) */
            int st = lzfo_compress2(j->in + j->in_off[b], j->in_len[b], 0, t, j->out + j->out_off[b], j->in_len[b], &w);
            j->out_len[b] = (uint32_t)w;
            j->status[b] = st;
        } else {
            size_t olen = 0;
            int st = lzfo_decompress_raw(j->in + j->in_off[b], j->in_len[b], NULL, 0, j->out + j->out_off[b],
                                         j->out_cap[b], j->out_limit[b], &olen);
            j->out_len[b] = (uint32_t)olen;
            j->status[b] = st;
        }
    }
    lzfo_table_free(t);
    return NULL;
}

static int mt_run(mt_job* j, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 512) nthreads = 512;
    pthread_t th[512];
    j->next = 0;
    int started = 0;
    for (int i = 1; i < nthreads; i++)
        if (pthread_create(&th[started], NULL, mt_worker, j) == 0) started++;
    mt_worker(j);
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
    return 0;
}

int lzfo_compress_blocks_mt(const uint8_t* in, const uint64_t* in_off, const uint32_t* in_len,
                            uint32_t nblocks, unsigned hashlog, uint8_t* out, const uint64_t* out_off,
                            uint32_t* out_len, int32_t* status, int nthreads) {
    mt_job j;
    memset(&j, 0, sizeof(j));
    j.compress = 1; j.in = in; j.in_off = in_off; j.in_len = in_len; j.nblocks = nblocks;
    j.hashlog = hashlog ? hashlog : 12; j.out = out; j.out_off = out_off; j.out_len = out_len; j.status = status;
    return mt_run(&j, nthreads);
}

int lzfo_decompress_blocks_mt(const uint8_t* in, const uint64_t* in_off, const uint32_t* in_len,
                              uint32_t nblocks, uint8_t* out, const uint64_t* out_off,
                              const uint32_t* out_cap, const uint32_t* out_limit, uint32_t* out_len,
                              int32_t* status, int nthreads) {
    mt_job j;
    memset(&j, 0, sizeof(j));
    j.compress = 0; j.in = in; j.in_off = in_off; j.in_len = in_len; j.nblocks = nblocks;
    j.out = out; j.out_off = out_off; j.out_cap = out_cap; j.out_limit = out_limit;
    j.out_len = out_len; j.status = status;
    return mt_run(&j, nthreads);
}

/* ------------------------------------------------------------------------ */
/* second CPU bar: C lz4 (liblz4.so.1, dlopen'ed — the image ships no header) */
/* README.md:11,18 places lz-fear at ~1x C for decode and 2-3x slower for     */
/* encode, so this is the stricter bar.  Same thread pool, one block per task */
/* ------------------------------------------------------------------------ */
#include <dlfcn.h>
typedef int (*lz4_compress_fn)(const char*, char*, int, int);
typedef int (*lz4_decompress_fn)(const char*, char*, int, int);
typedef struct {
    lz4_compress_fn comp; lz4_decompress_fn decomp;
    const uint8_t* in; const uint64_t* in_off; const uint32_t* in_len; uint32_t nblocks;
    uint8_t* out; const uint64_t* out_off; const uint32_t* out_cap; uint32_t* out_len; int32_t* status;
    volatile uint32_t next;
} lz4_job;

static void* lz4_worker(void* arg) {
    lz4_job* j = (lz4_job*)arg;
    for (;;) {
        uint32_t b = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (b >= j->nblocks) break;
        int r;
        if (j->comp) r = j->comp((const char*)j->in + j->in_off[b], (char*)j->out + j->out_off[b], (int)j->in_len[b], (int)j->out_cap[b]);
        else r = j->decomp((const char*)j->in + j->in_off[b], (char*)j->out + j->out_off[b], (int)j->in_len[b], (int)j->out_cap[b]);
        j->out_len[b] = r > 0 ? (uint32_t)r : 0;
        j->status[b] = r > 0 ? 0 : 1;       /* compress: 0 = does not fit (stored block); decompress: < 0 = malformed */
    }
    return NULL;
}

/* returns 0, or -1 when liblz4.so.1 is not installed */
int lzfo_liblz4_blocks_mt(int compress, const uint8_t* in, const uint64_t* in_off, const uint32_t* in_len, uint32_t nblocks,
                          uint8_t* out, const uint64_t* out_off, const uint32_t* out_cap, uint32_t* out_len,
                          int32_t* status, int nthreads) {
    static void* h = NULL;
    if (!h) h = dlopen("liblz4.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return -1;
    lz4_job j;
    memset(&j, 0, sizeof(j));
    if (compress) j.comp = (lz4_compress_fn)dlsym(h, "LZ4_compress_default");
    else j.decomp = (lz4_decompress_fn)dlsym(h, "LZ4_decompress_safe");
    if (!j.comp && !j.decomp) return -1;
    j.in = in; j.in_off = in_off; j.in_len = in_len; j.nblocks = nblocks;
    j.out = out; j.out_off = out_off; j.out_cap = out_cap; j.out_len = out_len; j.status = status;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 512) nthreads = 512;
    pthread_t th[512];
    int started = 0;
    for (int i = 1; i < nthreads; i++)
        if (pthread_create(&th[started], NULL, lz4_worker, &j) == 0) started++;
    lz4_worker(&j);
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
    return 0;
}
