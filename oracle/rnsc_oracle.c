/* oracle/rnsc_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Plain-C restatement of the reference's per-chunk raw-signal codecs. Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library;
 * the product (librnsc.so) never does. Every function cites the reference file:line it follows.
 * Pinned by tests/test_oracle_*.py against (a) the 22 stored VBZ frames of the reference's
 * pod5/test_data/multi_fast5_zip_v3.pod5 (golden fixtures under tests/golden/), (b) the reference's own
 * unit-test input recipes, and (c) outputs of the reference's code compiled here (oracle/_ref).
 *
 * The zstd stage is third-party (libzstd, NOT under /root/reference; upstream pins zstd/1.5.4 in
 * pod5/conanfile.py:97, this image ships libzstd.so.1.5.5): encode calls the image's libzstd through
 * dlopen exactly like the reference links it; decode can use either libzstd or this repo's RFC 8878
 * restatement (zstd_decode_oracle.c).
 *
 * Codec ids (shared with include/rnsc.h): 0 VBZ, 1 VBZ0, 2 C1, 3 C2, 4 C3, 5 C4, 6 C5, 7 RC.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <dlfcn.h>

uint64_t orc_zstd_frame_content_size(const uint8_t* src, size_t size);
int orc_zstd_decode_frame(const uint8_t* src, size_t size, uint8_t* dst, size_t cap, size_t* out_size);

enum { VBZ = 0, VBZ0 = 1, C1 = 2, C2 = 3, C3 = 4, C4 = 5, C5 = 6, RC = 7 };

/* ------------------------------------------------------------------------------------------------
 * libzstd binding (the same shared object the reference would link)
 * ---------------------------------------------------------------------------------------------- */
static size_t (*z_compress)(void*, size_t, const void*, size_t, int);
static size_t (*z_decompress)(void*, size_t, const void*, size_t);
static unsigned long long (*z_fcs)(const void*, size_t);
static size_t (*z_bound)(size_t);
static unsigned (*z_iserr)(size_t);
static unsigned (*z_version)(void);
static int z_load(void)
{
    if (z_compress) return 0;
    void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return -1;
    z_decompress = (size_t(*)(void*, size_t, const void*, size_t))dlsym(h, "ZSTD_decompress");
    z_fcs = (unsigned long long (*)(const void*, size_t))dlsym(h, "ZSTD_getFrameContentSize");
    z_bound = (size_t(*)(size_t))dlsym(h, "ZSTD_compressBound");
    z_iserr = (unsigned (*)(size_t))dlsym(h, "ZSTD_isError");
    z_version = (unsigned (*)(void))dlsym(h, "ZSTD_versionNumber");
    z_compress = (size_t(*)(void*, size_t, const void*, size_t, int))dlsym(h, "ZSTD_compress");
    return (z_compress && z_decompress && z_fcs && z_bound && z_iserr) ? 0 : -1;
}
unsigned orc_zstd_version(void) { return z_load() ? 0 : z_version(); }

/* ------------------------------------------------------------------------------------------------
 * shared arithmetic
 * ---------------------------------------------------------------------------------------------- */
/* svb16/svb16.h:11-21, pgnano/svb16/svb16.h:11-22 */
uint32_t orc_key_len1(uint32_t n) { return (n >> 3) + (((n & 7) + 7) >> 3); }
uint32_t orc_key_len2(uint32_t n) { return (n >> 2) + (((n & 3) + 3) >> 2); }
/* svb16/encode_scalar.hpp:11-14 */
static uint16_t zz_enc(uint16_t v) { return (uint16_t)((v + v) ^ (uint16_t)((int16_t)v >> 15)); }
/* svb16/decode_scalar.hpp:14-17 */
static uint16_t zz_dec(uint16_t v) { return (uint16_t)((v >> 1) ^ (uint16_t)(0 - (v & 1))); }

typedef struct { uint8_t* p[5]; size_t n[5]; } streams_t;

/* nibble appender shared by VBZ0/C4/C5 (C5.hpp:111-126; VBZ_0.hpp:143-160): low nibble first */
typedef struct { uint8_t* p; size_t nib; } nibw;
static void nib_put(nibw* w, unsigned v) {
    if ((w->nib & 1) == 0) w->p[w->nib >> 1] = (uint8_t)(v & 15); else w->p[w->nib >> 1] |= (uint8_t)((v & 15) << 4);
    w->nib++;
}

/* Pre-zstd streams of every splitter. out = 5 regions of stride (2n+16); sizes in wire order.
 *   VBZ : encode_scalar.hpp:17-66            -> keys1, data
 *   VBZ0: VBZ_0.hpp:61-172                   -> keys2, nibble data
 *   C1  : C1.hpp:58-109                      -> keys1, data
 *   C2  : C2.hpp:53-109                      -> keys1, low(all), high(large)
 *   C3  : C3.hpp:54-110                      -> keys1, ll, lh, h
 *   C4  : C4.hpp:54-145                      -> keys2, S, M, Llow, Lhigh (no offsets)
 *   C5  : C5.hpp:58-152                      -> keys2, S, M, Llow, Lhigh (offsets 1/17/273)       */
int orc_streams(int codec, const int16_t* in, size_t n, uint8_t* out, size_t* sizes)
{
    size_t stride = 2 * n + 16;
    uint8_t* s[5]; for (int i = 0; i < 5; ++i) { s[i] = out + stride * i; sizes[i] = 0; }
    if (codec < VBZ || codec > C5) return 1;
    if (n == 0) return 0;                                     /* every encoder returns empty streams */
    int two_bit = (codec == VBZ0 || codec == C4 || codec == C5);
    size_t klen = two_bit ? orc_key_len2((uint32_t)n) : orc_key_len1((uint32_t)n);
    memset(s[0], 0, klen);
    size_t c1 = 0, c2 = 0, c3 = 0, c4 = 0; nibw nw = {s[1], 0};
    uint16_t prev = 0;
    for (size_t i = 0; i < n; ++i) {
        uint16_t z = zz_enc((uint16_t)((uint16_t)in[i] - prev)); prev = (uint16_t)in[i];
        switch (codec) {
        case VBZ: case C1:
            if (z < 256) s[1][c1++] = (uint8_t)z;
            else { s[1][c1++] = (uint8_t)z; s[1][c1++] = (uint8_t)(z >> 8); s[0][i >> 3] |= (uint8_t)(1 << (i & 7)); }
            break;
        case C2:
            s[1][c1++] = (uint8_t)z;
            if (z >= 256) { s[2][c2++] = (uint8_t)(z >> 8); s[0][i >> 3] |= (uint8_t)(1 << (i & 7)); }
            break;
        case C3:
            if (z >= 256) { s[2][c2++] = (uint8_t)z; s[3][c3++] = (uint8_t)(z >> 8); s[0][i >> 3] |= (uint8_t)(1 << (i & 7)); }
            else s[1][c1++] = (uint8_t)z;
            break;
        case C4: case C5: case VBZ0: {
            unsigned code; uint16_t v = z;
            if (codec == C4) code = z == 0 ? 0 : (z < 16 ? 1 : (z < 256 ? 2 : 3));
            else if (z == 0) code = 0;
            else if (z <= 16) { code = 1; v = (uint16_t)(z - 1); }
            else if (z <= 272) { code = 2; v = (uint16_t)(z - 17); }
            else { code = 3; v = (uint16_t)(z - 273); }
            s[0][i >> 2] |= (uint8_t)(code << (2 * (i & 3)));
            if (codec == VBZ0) { for (unsigned k = 0; k < (code == 3 ? 4u : code); ++k) { nib_put(&nw, v & 15); v >>= 4; } }
            else if (code == 1) nib_put(&nw, v);
            else if (code == 2) s[2][c2++] = (uint8_t)v;
            else if (code == 3) { s[3][c3++] = (uint8_t)v; s[4][c4++] = (uint8_t)(v >> 8); }
            break; }
        }
    }
    sizes[0] = klen;
    if (two_bit) { sizes[1] = (nw.nib + 1) >> 1; sizes[2] = c2; sizes[3] = c3; sizes[4] = c4; }
    else { sizes[1] = c1; sizes[2] = c2; sizes[3] = c3; }
    return 0;
}

/* Inverse splitters over the contiguous decompressed buffer [keys|s1|s2|s3|s4] (the reference's
 * "intermediate"): decode_scalar.hpp:36-75 (VBZ), VBZ_0.hpp:186-290, C1.hpp:141-179, C2.hpp:122-178,
 * C3.hpp:131-189, C4.hpp:167-249, C5.hpp:174-258. d[] = decompressed size of each frame in wire order.
 * Returns 0 and *consumed as the reference's decode_* return value (pointer - buffer start); the
 * caller applies the "Remaining data at end of signal buffer" check. Returns 2 if a stream would be
 * read out of bounds (the reference only asserts). */
int orc_unpack(int codec, const uint8_t* buf, const size_t* d, size_t n, int16_t* out, size_t* consumed)
{
    size_t total = 0; int nstreams = (codec == VBZ || codec == VBZ0 || codec == C1) ? 2 : (codec == C2 ? 3 : (codec == C3 ? 4 : 5));
    if (codec == VBZ || codec == VBZ0) nstreams = 1;          /* one frame holding keys||data */
    for (int i = 0; i < nstreams; ++i) total += d[i];
    int two_bit = (codec == VBZ0 || codec == C4 || codec == C5);
    size_t klen = n ? (two_bit ? orc_key_len2((uint32_t)n) : orc_key_len1((uint32_t)n)) : 0;
    if (n == 0) { /* each decoder returns the begin() of one of its spans */
        *consumed = (codec == C4 || codec == C5) ? (klen + d[1] + d[2] + d[3]) : klen; return 0; }
    if (klen > total) return 2;
    const uint8_t* keys = buf;
    /* stream start offsets inside buf */
    size_t o1 = klen, o2, o3, o4;
    if (codec == VBZ || codec == VBZ0 || codec == C1) { o2 = o3 = o4 = total; }
    else { o2 = o1 + d[1]; o3 = o2 + d[2]; o4 = (codec == C3) ? total : o3 + d[3]; if (codec == C2) o3 = o4 = total; }
    size_t p1 = o1, p2 = o2, p3 = o3, p4 = o4, nib = 0; uint16_t prev = 0;
    for (size_t i = 0; i < n; ++i) {
        uint16_t z = 0;
        if (!two_bit) {
            int k = (keys[i >> 3] >> (i & 7)) & 1;
            if (codec == VBZ || codec == C1) {
                if (p1 + 1 + k > total) return 2;
                z = buf[p1++]; if (k) z |= (uint16_t)(buf[p1++] << 8);
            } else if (codec == C2) {
                if (p1 >= total) return 2; z = buf[p1++];
                if (k) { if (p2 >= total) return 2; z |= (uint16_t)(buf[p2++] << 8); }
            } else { /* C3 */
                if (k) { if (p2 >= total || p3 >= total) return 2; z = buf[p2++]; z |= (uint16_t)(buf[p3++] << 8); }
                else { if (p1 >= total) return 2; z = buf[p1++]; }
            }
        } else {
            int code = (keys[i >> 2] >> (2 * (i & 3))) & 3;
            if (codec == VBZ0) {
                int nn = code == 3 ? 4 : code; unsigned v = 0;
                for (int k = 0; k < nn; ++k) { size_t b = o1 + (nib >> 1); if (b >= total) return 2;
                    v |= (unsigned)((buf[b] >> (4 * (nib & 1))) & 15) << (4 * k); nib++; }
                z = (uint16_t)(code == 0 ? 0 : v + (code == 1 ? 1 : (code == 2 ? 17 : 273)));
            } else {
                unsigned v = 0;
                if (code == 1) { size_t b = o1 + (nib >> 1); if (b >= total) return 2; v = (buf[b] >> (4 * (nib & 1))) & 15; nib++; }
                else if (code == 2) { if (p2 >= total) return 2; v = buf[p2++]; }
                else if (code == 3) { if (p3 >= total || p4 >= total) return 2; v = (unsigned)(buf[p4++] << 8) + buf[p3++]; }
                if (codec == C5 && code) v += (code == 1 ? 1 : (code == 2 ? 17 : 273));
                z = (uint16_t)v;
            }
        }
        prev = (uint16_t)(prev + zz_dec(z));
        out[i] = (int16_t)prev;
    }
    switch (codec) {
    case VBZ: case C1: *consumed = p1; break;
    case VBZ0: *consumed = o1 + ((nib + 1) >> 1); break;
    case C2: *consumed = p2; break;
    case C3: *consumed = p3; break;
    default: *consumed = p4; break;
    }
    return 0;
}

/* signal_compression.cpp:30-35 (VBZ); pgnano/compressor.h:39-45 + header.h:15 (every Cx and RC) */
size_t orc_max_size(int codec, size_t n)
{
    if (codec == VBZ) { if (z_load()) return 0; return z_bound(orc_key_len1((uint32_t)n) + 2 * n); }
    size_t sz = 2 * n + 10 + 16;
    return sz > 1024 ? sz : 1024;
}

static int codec_nframes(int codec) { return codec == VBZ || codec == VBZ0 ? 1 : (codec == C1 ? 2 : (codec == C2 ? 3 : (codec == C3 ? 4 : 5))); }

/* Whole-chunk encode. VBZ: signal_compression.cpp:37-78. VBZ0: VBZ_0.hpp:316-362. C1: C1.hpp:202-285.
 * C2: C2.hpp:195-316. C3: C3.hpp:212-354. C4: C4.hpp:274-470. C5: C5.hpp:282-474.
 * Container: u64 size prefix before every frame but the last. Status codes are rnsc.h's: 0 ok;
 * 6 zstd failure ("Failed to compress ..."); 3 "Not enough space in destination buffer" (C4/C5 check it, C5.hpp:420-427;
 * C1-C3 overrun in the reference -- reported as the same error here). */
int orc_encode(int codec, const int16_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_size)
{
    if (z_load()) return 9;
    size_t stride = 2 * n + 16, sizes[5];
    uint8_t* st = (uint8_t*)malloc(stride * 5 + 16);
    if (orc_streams(codec, in, n, st, sizes)) { free(st); return 9; }
    int rc = 0; size_t op = 0;
    if (codec == VBZ || codec == VBZ0) {                    /* one frame over keys||data, straight into dst */
        uint8_t* cat = (uint8_t*)malloc(sizes[0] + sizes[1] + 16);
        memcpy(cat, st, sizes[0]); memcpy(cat + sizes[0], st + stride, sizes[1]);
        size_t c = z_compress(out, cap, cat, sizes[0] + sizes[1], 1);
        if (z_iserr(c)) rc = 6; else op = c;
        free(cat);
    } else {
        int nf = codec_nframes(codec);
        uint8_t* tmp[5] = {0}; size_t cs[5]; size_t total = 0;
        for (int f = 0; f < nf && !rc; ++f) {
            size_t b = z_bound(sizes[f]); tmp[f] = (uint8_t*)malloc(b);
            cs[f] = z_compress(tmp[f], b, st + stride * f, sizes[f], 1);
            if (z_iserr(cs[f])) rc = 6;
            total += cs[f] + (f + 1 < nf ? 8 : 0);
        }
        if (!rc && total > cap) rc = 3;
        for (int f = 0; f < nf && !rc; ++f) {
            if (f + 1 < nf) { uint64_t v = cs[f]; memcpy(out + op, &v, 8); op += 8; }
            memcpy(out + op, tmp[f], cs[f]); op += cs[f];
        }
        for (int f = 0; f < nf; ++f) free(tmp[f]);
    }
    free(st);
    *out_size = op;
    return rc;
}

/* Whole-chunk decode. VBZ: signal_compression.cpp:96-141; Cx: A.8 / C5.hpp:477-683 and siblings.
 * use_own_zstd != 0 routes frames through zstd_decode_oracle.c instead of libzstd.
 * Returns 0 ok; 1 "Input data not compressed by zstd"; 2 "failed to decompress"; 4 "Remaining data at
 * end of signal buffer"; 5 malformed container. */
int orc_decode(int codec, const uint8_t* in, size_t clen, size_t n, int16_t* out, int use_own_zstd)
{
    if (z_load()) return 9;
    int nf = codec_nframes(codec);
    const uint8_t* fp[5]; size_t fl[5], d[5] = {0, 0, 0, 0, 0}; size_t pos = 0, total = 0;
    for (int f = 0; f < nf; ++f) {
        if (f + 1 < nf) { if (pos + 8 > clen) return 5; uint64_t v; memcpy(&v, in + pos, 8); pos += 8; if (v > clen - pos) return 5; fl[f] = (size_t)v; }
        else fl[f] = clen - pos;
        fp[f] = in + pos; pos += fl[f];
        unsigned long long c = use_own_zstd ? orc_zstd_frame_content_size(fp[f], fl[f]) : z_fcs(fp[f], fl[f]);
        if (c >= (unsigned long long)-2) return 1;
        /* a valid chunk of n samples never decodes to more than its key bytes + 2 bytes per sample (+ slack): reject
         * hostile frame headers before sizing the buffer from them, as the kernels do (plan.cuh: fcs > stream_bound) */
        if (c > (unsigned long long)n / 4 + 2ull * n + 16) return 4;
        d[f] = (size_t)c; total += d[f];
    }
    uint8_t* buf = (uint8_t*)malloc(total + 16); size_t wp = 0; int rc = 0;
    if (!buf) return 2;
    for (int f = 0; f < nf && !rc; ++f) {
        size_t got = 0;
        if (use_own_zstd) { if (orc_zstd_decode_frame(fp[f], fl[f], buf + wp, d[f], &got)) rc = 2; }
        else { got = z_decompress(buf + wp, d[f], fp[f], fl[f]); if (z_iserr(got)) rc = 2; }
        if (!rc && got != d[f]) rc = 2;
        wp += d[f];
    }
    if (!rc) {
        size_t consumed = 0;
        /* the reference slices streams by the *decompressed* sizes (decode_N01 etc.) */
        size_t dd[5] = {d[0], d[1], d[2], d[3], d[4]};
        rc = orc_unpack(codec, buf, dd, n, out, &consumed);
        if (!rc && consumed != total) rc = 4;
        else if (rc) rc = 4;
    }
    free(buf);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * Legacy pgnano range coder (RC): pgnano/compressor.cpp:39-88, decompressor.cpp:45-73,
 * commonLib/codecs.cpp:6-20, fqzcomp/clr.h:44-130, fqzcomp/simple_model.h:80-92,169-176,206-221,325-343
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint16_t f[257]; uint32_t tot; } rc_model;
static void model_init(rc_model* m) { for (int i = 0; i < 256; ++i) m->f[i] = 1; m->f[256] = 0; m->tot = 256; }
static void model_bump(rc_model* m, int s)   /* simple_model.h:216-220 + normalize :169-176 */
{
    m->f[s] = (uint16_t)(m->f[s] + 8); m->tot += 8;
    if (m->tot > 65504) { m->tot = 0; for (int i = 0; m->f[i]; ++i) { m->f[i] = (uint16_t)(m->f[i] - (m->f[i] >> 1)); m->tot += m->f[i]; } }
}
typedef struct { uint64_t low; uint32_t range, code; uint8_t* out; const uint8_t* in; } rc_state;
#define RC_TOP (1u << 24)
static void rc_encode(rc_state* r, uint32_t cum, uint32_t freq, uint32_t tot)   /* clr.h:80-104 */
{
    r->range /= tot; r->low += (uint32_t)(cum * r->range); r->range *= freq;
    while (r->range < RC_TOP) {
        if ((uint8_t)((r->low ^ (r->low + r->range)) >> 56)) r->range = (((uint32_t)r->low | (RC_TOP - 1)) - (uint32_t)r->low);
        *r->out++ = (uint8_t)(r->low >> 56); r->range <<= 8; r->low <<= 8;
    }
}
static void sym_encode(rc_state* r, rc_model* m, int s)      /* simple_model.h:206-221 */
{
    uint32_t acc = 0; for (int i = 0; i < s; ++i) acc += m->f[i];
    rc_encode(r, acc, m->f[s], m->tot); model_bump(m, s);
}
static int sym_decode(rc_state* r, rc_model* m)              /* simple_model.h:325-343, clr.h:115-130 */
{
    r->range /= m->tot; uint32_t fq = r->code / r->range;
    uint32_t acc = 0; int s = 0;
    while ((acc += m->f[s]) <= fq) { s++; if (s > 255) { s = 255; break; } }
    acc -= m->f[s];
    uint32_t temp = acc * r->range; r->low += temp; r->code -= temp; r->range *= m->f[s];
    while (r->range < RC_TOP) {
        if ((uint8_t)((r->low ^ (r->low + r->range)) >> 56)) r->range = (((uint32_t)r->low | (RC_TOP - 1)) - (uint32_t)r->low);
        r->code = (r->code << 8) | *r->in++; r->range <<= 8; r->low <<= 8;
    }
    model_bump(m, s);
    return s;
}
/* codecs.cpp:6-12 on the int16-truncated difference == zigzag on the mod-2^16 difference */
/* compressor.cpp:72-88 (header) + :39-70 (body). out must hold orc_max_size(RC,n). Returns bytes written,
 * or 0 if the reference's raw-fallback branch (bytes_written > 2n, compressor.cpp:60-65) would be taken:
 * that branch is broken in the reference (copies in the wrong direction) and is an error here. */
size_t orc_rc_encode(const int16_t* in, size_t n, uint8_t* out, uint8_t pore_type)
{
    out[0] = 0; for (int i = 0; i < 8; ++i) out[1 + i] = (uint8_t)((uint64_t)n >> (8 * i)); out[9] = pore_type;
    rc_state r; r.low = 0; r.range = 0xFFFFFFFFu; r.code = 0; r.out = out + 10; r.in = NULL;
    rc_model H, L0, L1; model_init(&H); model_init(&L0); model_init(&L1);
    uint16_t prev = 0;
    for (size_t i = 0; i < n; ++i) {
        uint16_t m = zz_enc((uint16_t)((uint16_t)in[i] - prev)); prev = (uint16_t)in[i];
        int hi = m >> 8, lo = m & 0xFF;
        sym_encode(&r, &H, hi);
        sym_encode(&r, hi ? &L1 : &L0, lo);
        if ((size_t)(r.out - (out + 10)) > 2 * n) return 0;
    }
    for (int i = 0; i < 8; ++i) { *r.out++ = (uint8_t)(r.low >> 56); r.low <<= 8; }   /* clr.h:65-69 */
    return (size_t)(r.out - out);
}
/* decompressor.cpp:16-73. in must have >= 8 readable bytes after the coded body (the 8 flush bytes
 * always provide this for streams produced by orc_rc_encode / the reference). */
int orc_rc_decode(const uint8_t* in, size_t clen, size_t n, int16_t* out)
{
    if (clen < 18) return 1;
    if (in[0] & 1) return 3;                                   /* raw flag: never produced, rejected */
    uint64_t hn = 0; for (int i = 0; i < 8; ++i) hn |= (uint64_t)in[1 + i] << (8 * i);
    if (hn != n) return 2;
    rc_state r; r.low = 0; r.range = 0xFFFFFFFFu; r.code = 0; r.in = in + 10; r.out = NULL;
    for (int i = 0; i < 8; ++i) r.code = (r.code << 8) | *r.in++;   /* clr.h:54-59 (u32: first 4 shift out) */
    rc_model H, L0, L1; model_init(&H); model_init(&L0); model_init(&L1);
    uint16_t prev = 0;
    for (size_t i = 0; i < n; ++i) {
        if ((size_t)(r.in - in) > clen + 8) return 4;
        int hi = sym_decode(&r, &H);
        int lo = sym_decode(&r, hi ? &L1 : &L0);
        prev = (uint16_t)(prev + zz_dec((uint16_t)((hi << 8) | lo)));
        out[i] = (int16_t)prev;
    }
    return 0;
}
