/* oracle/zstd_decode_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Plain-C restatement of the Zstandard frame *decoder* (RFC 8878), i.e. of the third-party call
 * ZSTD_getFrameContentSize + ZSTD_decompress that the reference makes at
 *   pod5/c++/pod5_format/signal_compression.cpp:103,117 and pgnano/svb16/C5.hpp:494-565,588-655
 * (libzstd is NOT under /root/reference; upstream pins zstd/1.5.4 in pod5/conanfile.py:97, the
 * container has libzstd.so.1.5.5). It is the executable specification for the CUDA frame decoder in
 * rawnanoporesignalcompression_b200/csrc/zstd_decode.cuh and is itself pinned against libzstd 1.5.5
 * by tests/test_oracle_zstd.py (stored frames of pod5/test_data + frames produced by libzstd here).
 *
 * Supported: everything ZSTD_compress can emit for a one-shot call: single/multi-block frames,
 * raw/RLE/compressed blocks, raw/RLE/Huffman(1|4 streams)/treeless literals, predefined/RLE/FSE/
 * repeat sequence tables, repeat offsets, optional checksum (skipped, not verified), window
 * descriptor. Not supported: dictionaries, skippable frames, legacy formats.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

#define ZERR_SRC_SIZE 1   /* truncated input */
#define ZERR_MAGIC 2
#define ZERR_UNSUPPORTED 3
#define ZERR_CORRUPT 4
#define ZERR_DST_SIZE 5

typedef struct { uint16_t newState; uint8_t symbol; uint8_t nbBits; } fse_entry;

/* ---- forward (LSB-first) bit reader used by FSE table descriptions ---- */
typedef struct { const uint8_t* p; size_t size; size_t bitpos; } fwd_bits;
static uint32_t fwd_peek(const fwd_bits* b, int n)
{
    uint64_t v = 0; size_t byte = b->bitpos >> 3;
    for (int i = 0; i < 5; ++i) if (byte + i < b->size) v |= (uint64_t)b->p[byte + i] << (8 * i);
    return (uint32_t)((v >> (b->bitpos & 7)) & ((1ull << n) - 1));
}

/* FSE_readNCount: returns bytes consumed or 0 on error. norm[] receives counts (-1 = "less than 1"). */
static size_t fse_read_ncount(const uint8_t* src, size_t size, int16_t* norm, int max_sym, int max_log,
                              int* table_log, int* nsym)
{
    fwd_bits b = {src, size, 0};
    if (size < 1) return 0;
    int tl = (int)fwd_peek(&b, 4) + 5; b.bitpos += 4;
    if (tl > max_log) return 0;
    int remaining = (1 << tl) + 1, threshold = 1 << tl, nbBits = tl + 1, sym = 0, prev0 = 0;
    while (remaining > 1 && sym <= max_sym) {
        if (prev0) {
            int n0 = sym;
            for (;;) {                        /* 2-bit repeat flags: 3 = "3 more zeros and continue" */
                uint32_t r = fwd_peek(&b, 2); b.bitpos += 2;
                n0 += (int)r;
                if (r != 3) break;
            }
            if (n0 > max_sym + 1) return 0;
            while (sym < n0) norm[sym++] = 0;
            if (sym > max_sym) break;
        }
        int max = (2 * threshold - 1) - remaining, count;
        uint32_t lo = fwd_peek(&b, nbBits - 1);
        if ((int)lo < max) { count = (int)lo; b.bitpos += nbBits - 1; }
        else {
            count = (int)fwd_peek(&b, nbBits);
            if (count >= threshold) count -= max;
            b.bitpos += nbBits;
        }
        count--;
        remaining -= count < 0 ? -count : count;
        norm[sym++] = (int16_t)count;
        prev0 = (count == 0);
        while (remaining < threshold) { nbBits--; threshold >>= 1; }
    }
    if (remaining != 1) return 0;
    size_t used = (b.bitpos + 7) >> 3;
    if (used > size) return 0;
    *table_log = tl; *nsym = sym;
    return used;
}

static int highbit32(uint32_t v) { int r = 0; while (v >>= 1) r++; return r; }

/* FSE_buildDTable / ZSTD_buildFSETable (symbol spread with step (size>>1)+(size>>3)+3). */
static void fse_build(fse_entry* t, const int16_t* norm, int nsym, int tl)
{
    int size = 1 << tl, high = size - 1;
    uint16_t next[256];
    for (int s = 0; s < nsym; ++s) {
        if (norm[s] == -1) { t[high--].symbol = (uint8_t)s; next[s] = 1; }
        else next[s] = (uint16_t)norm[s];
    }
    int step = (size >> 1) + (size >> 3) + 3, mask = size - 1, pos = 0;
    for (int s = 0; s < nsym; ++s)
        for (int i = 0; i < norm[s]; ++i) {
            t[pos].symbol = (uint8_t)s;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    for (int u = 0; u < size; ++u) {
        uint8_t s = t[u].symbol;
        uint32_t ns = next[s]++;
        int nb = tl - highbit32(ns);
        t[u].nbBits = (uint8_t)nb;
        t[u].newState = (uint16_t)((ns << nb) - size);
    }
}

/* ---- backward (MSB-first from the end) bit reader: Huffman streams, FSE streams ---- */
typedef struct { const uint8_t* start; int64_t bitpos; /* number of unread bits */ int overflow; } bwd_bits;
static int bwd_init(bwd_bits* b, const uint8_t* p, size_t size)
{
    if (size == 0) return -1;
    uint8_t last = p[size - 1];
    if (last == 0) return -1;
    b->start = p; b->overflow = 0;
    b->bitpos = (int64_t)size * 8 - (8 - highbit32(last));   /* drop padding + the end-mark bit */
    return 0;
}
/* read n bits (n<=32); bits beyond the start read as 0 (zstd semantics) and set overflow */
static uint32_t bwd_read(bwd_bits* b, int n)
{
    if (n == 0) return 0;
    int64_t lo = b->bitpos - n;               /* index of lowest bit of the field */
    uint64_t v = 0;
    for (int i = 0; i < n; ++i) {
        int64_t bi = lo + i;
        if (bi >= 0) v |= (uint64_t)((b->start[bi >> 3] >> (bi & 7)) & 1) << i;
    }
    b->bitpos = lo;
    if (lo < 0) b->overflow = 1;
    return (uint32_t)v;
}
static uint32_t bwd_peek(const bwd_bits* b, int n)
{
    bwd_bits c = *b; return bwd_read(&c, n);
}

/* ---- Huffman ---- */
typedef struct { uint8_t symbol; uint8_t nbBits; } huf_entry;
typedef struct { huf_entry t[1 << 12]; int maxBits; int valid; } huf_table;

/* Decode the Huffman tree description; returns bytes consumed or 0 on error. */
static size_t huf_read_table(huf_table* ht, const uint8_t* src, size_t size)
{
    uint8_t w[256]; int nw = 0;
    if (size < 1) return 0;
    int hb = src[0]; size_t used;
    if (hb >= 128) {                              /* direct: 4-bit weights */
        nw = hb - 127;
        used = 1 + (size_t)((nw + 1) / 2);
        if (used > size) return 0;
        for (int i = 0; i < nw; ++i) w[i] = (i & 1) ? (src[1 + i / 2] & 15) : (src[1 + i / 2] >> 4);
    } else {                                      /* FSE-compressed weights, two interleaved states */
        used = 1 + (size_t)hb;
        if (used > size || hb < 1) return 0;
        int16_t norm[16]; int tl, ns;
        size_t h = fse_read_ncount(src + 1, (size_t)hb, norm, 12, 6, &tl, &ns);
        if (!h) return 0;
        fse_entry ft[64];
        fse_build(ft, norm, ns, tl);
        bwd_bits b;
        if (bwd_init(&b, src + 1 + h, (size_t)hb - h)) return 0;
        uint32_t s1 = bwd_read(&b, tl), s2 = bwd_read(&b, tl);
        for (;;) {   /* FSE_decompress_usingDTable tail semantics */
            if (nw >= 254) return 0;
            w[nw++] = ft[s1].symbol;
            { int nb = ft[s1].nbBits; int64_t before = b.bitpos; s1 = ft[s1].newState + bwd_read(&b, nb);
              if (before - nb < 0) { if (nw >= 255) return 0; w[nw++] = ft[s2].symbol; break; } }
            if (nw >= 254) return 0;
            w[nw++] = ft[s2].symbol;
            { int nb = ft[s2].nbBits; int64_t before = b.bitpos; s2 = ft[s2].newState + bwd_read(&b, nb);
              if (before - nb < 0) { if (nw >= 255) return 0; w[nw++] = ft[s1].symbol; break; } }
        }
    }
    /* last weight is implied */
    uint32_t total = 0;
    for (int i = 0; i < nw; ++i) { if (w[i] > 12) return 0; total += w[i] ? (1u << (w[i] - 1)) : 0; }
    if (total == 0) return 0;
    int maxBits = highbit32(total) + 1;
    if (maxBits > 12) return 0;
    uint32_t rest = (1u << maxBits) - total;
    if (rest & (rest - 1)) return 0;              /* must be a power of two */
    w[nw++] = (uint8_t)(highbit32(rest) + 1);
    /* fill: weight ascending, symbols in natural order inside a weight */
    uint32_t rankStart[14] = {0}, cnt[14] = {0};
    for (int i = 0; i < nw; ++i) cnt[w[i]]++;
    if (cnt[1] < 2 || (cnt[1] & 1)) return 0;
    uint32_t pos = 0;
    for (int k = 1; k <= maxBits; ++k) { rankStart[k] = pos; pos += cnt[k] << (k - 1); }
    for (int i = 0; i < nw; ++i) {
        if (!w[i]) continue;
        uint32_t len = 1u << (w[i] - 1), st = rankStart[w[i]];
        for (uint32_t u = 0; u < len; ++u) { ht->t[st + u].symbol = (uint8_t)i; ht->t[st + u].nbBits = (uint8_t)(maxBits + 1 - w[i]); }
        rankStart[w[i]] += len;
    }
    ht->maxBits = maxBits; ht->valid = 1;
    return used;
}

static int huf_decode_stream(const huf_table* ht, const uint8_t* src, size_t size, uint8_t* dst, size_t n)
{
    bwd_bits b;
    if (bwd_init(&b, src, size)) return -1;
    for (size_t i = 0; i < n; ++i) {
        uint32_t v = bwd_peek(&b, ht->maxBits);
        huf_entry e = ht->t[v];
        dst[i] = e.symbol;
        b.bitpos -= e.nbBits;
    }
    return b.bitpos == 0 ? 0 : -1;                /* stream must be consumed exactly */
}

/* ---- sequences ---- */
static const uint32_t LL_base[36] = {0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,18,20,22,24,28,32,40,48,64,
                                     0x80,0x100,0x200,0x400,0x800,0x1000,0x2000,0x4000,0x8000,0x10000};
static const uint8_t LL_bits[36] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,6,7,8,9,10,11,12,13,14,15,16};
static const uint32_t ML_base[53] = {3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,30,31,32,
                                     33,34,35,37,39,41,43,47,51,59,67,83,99,0x83,0x103,0x203,0x403,0x803,0x1003,0x2003,
                                     0x4003,0x8003,0x10003};
static const uint8_t ML_bits[53] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,
                                    1,1,1,1,2,2,3,3,4,4,5,7,8,9,10,11,12,13,14,15,16};
static const int16_t LL_default[36] = {4,3,2,2,2,2,2,2,2,2,2,2,2,1,1,1,2,2,2,2,2,2,2,2,2,3,2,1,1,1,1,1,-1,-1,-1,-1};
static const int16_t ML_default[53] = {1,4,3,2,2,2,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,
                                       1,1,-1,-1,-1,-1,-1,-1,-1};
static const int16_t OF_default[29] = {1,1,1,1,1,1,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1};

typedef struct { fse_entry t[512]; int log; int valid; } seq_table;

typedef struct {
    huf_table huf;
    seq_table ll, of, ml;
    uint32_t rep[3];
} frame_ctx;

/* returns bytes consumed from src, or (size_t)-1 on error */
static size_t seq_table_setup(seq_table* st, int mode, const uint8_t* src, size_t size,
                              const int16_t* defnorm, int defn, int deflog, int max_sym, int max_log)
{
    switch (mode) {
    case 0: fse_build(st->t, defnorm, defn, deflog); st->log = deflog; st->valid = 1; return 0;
    case 1: if (size < 1 || src[0] > max_sym) return (size_t)-1;
            st->t[0].symbol = src[0]; st->t[0].nbBits = 0; st->t[0].newState = 0; st->log = 0; st->valid = 1; return 1;
    case 2: { int16_t norm[64]; int tl, ns;
              size_t h = fse_read_ncount(src, size, norm, max_sym, max_log, &tl, &ns);
              if (!h) return (size_t)-1;
              fse_build(st->t, norm, ns, tl); st->log = tl; st->valid = 1; return h; }
    default: return st->valid ? 0 : (size_t)-1;   /* repeat */
    }
}

static int decode_block(frame_ctx* fc, const uint8_t* src, size_t size, uint8_t* dst_base, size_t dst_pos,
                        size_t dst_cap, size_t* produced, uint8_t* litbuf)
{
    if (size < 1) return ZERR_CORRUPT;
    /* ---- literals section ---- */
    int ltype = src[0] & 3, sf = (src[0] >> 2) & 3;
    size_t lhs, regen, comp = 0; int nstreams = 1;
    const uint8_t* lit = NULL;       /* literal bytes (pointer into src for raw, litbuf otherwise) */
    if (ltype < 2) {
        if (sf == 0 || sf == 2) { lhs = 1; regen = src[0] >> 3; }
        else if (sf == 1) { if (size < 2) return ZERR_CORRUPT; lhs = 2; regen = (src[0] >> 4) | ((size_t)src[1] << 4); }
        else { if (size < 3) return ZERR_CORRUPT; lhs = 3; regen = (src[0] >> 4) | ((size_t)src[1] << 4) | ((size_t)src[2] << 12); }
    } else {
        if (size < 3) return ZERR_CORRUPT;
        uint64_t h = 0; for (int i = 0; i < 5 && (size_t)i < size; ++i) h |= (uint64_t)src[i] << (8 * i);
        if (sf == 0 || sf == 1) { lhs = 3; regen = (h >> 4) & 0x3FF; comp = (h >> 14) & 0x3FF; nstreams = sf ? 4 : 1; }
        else if (sf == 2) { lhs = 4; regen = (h >> 4) & 0x3FFF; comp = (h >> 18) & 0x3FFF; nstreams = 4; }
        else { if (size < 5) return ZERR_CORRUPT; lhs = 5; regen = (h >> 4) & 0x3FFFF; comp = (h >> 22) & 0x3FFFF; nstreams = 4; }
    }
    if (regen > 128 * 1024) return ZERR_CORRUPT;
    size_t lit_section;
    if (ltype == 0) { lit_section = lhs + regen; if (lit_section > size) return ZERR_CORRUPT; lit = src + lhs; }
    else if (ltype == 1) { lit_section = lhs + 1; if (lit_section > size) return ZERR_CORRUPT; memset(litbuf, src[lhs], regen); lit = litbuf; }
    else {
        lit_section = lhs + comp; if (lit_section > size) return ZERR_CORRUPT;
        const uint8_t* p = src + lhs; size_t left = comp;
        if (ltype == 2) { size_t u = huf_read_table(&fc->huf, p, left); if (!u) return ZERR_CORRUPT; p += u; left -= u; }
        else if (!fc->huf.valid) return ZERR_CORRUPT;
        if (nstreams == 1) { if (huf_decode_stream(&fc->huf, p, left, litbuf, regen)) return ZERR_CORRUPT; }
        else {
            if (left < 6) return ZERR_CORRUPT;
            size_t s1 = p[0] | (p[1] << 8), s2 = p[2] | (p[3] << 8), s3 = p[4] | (p[5] << 8);
            if (6 + s1 + s2 + s3 > left) return ZERR_CORRUPT;
            size_t s4 = left - 6 - s1 - s2 - s3, q = (regen + 3) / 4;
            if (3 * q > regen) return ZERR_CORRUPT;
            const uint8_t* a = p + 6;
            if (huf_decode_stream(&fc->huf, a, s1, litbuf, q)) return ZERR_CORRUPT;
            if (huf_decode_stream(&fc->huf, a + s1, s2, litbuf + q, q)) return ZERR_CORRUPT;
            if (huf_decode_stream(&fc->huf, a + s1 + s2, s3, litbuf + 2 * q, q)) return ZERR_CORRUPT;
            if (huf_decode_stream(&fc->huf, a + s1 + s2 + s3, s4, litbuf + 3 * q, regen - 3 * q)) return ZERR_CORRUPT;
        }
        lit = litbuf;
    }
    /* ---- sequences section ---- */
    const uint8_t* p = src + lit_section; size_t left = size - lit_section;
    if (left < 1) return ZERR_CORRUPT;
    size_t nseq = p[0];
    if (nseq == 0) { p += 1; left -= 1; }
    else if (nseq < 128) { p += 1; left -= 1; }
    else if (nseq < 255) { if (left < 2) return ZERR_CORRUPT; nseq = ((nseq - 128) << 8) + p[1]; p += 2; left -= 2; }
    else { if (left < 3) return ZERR_CORRUPT; nseq = p[1] + ((size_t)p[2] << 8) + 0x7F00; p += 3; left -= 3; }
    uint8_t* out = dst_base + dst_pos; size_t cap = dst_cap - dst_pos, op = 0, lp = 0;
    if (nseq == 0) {
        if (left != 0) return ZERR_CORRUPT;
        if (regen > cap) return ZERR_DST_SIZE;
        memcpy(out, lit, regen); *produced = regen; return 0;
    }
    if (left < 1) return ZERR_CORRUPT;
    int modes = p[0]; p += 1; left -= 1;
    if (modes & 3) return ZERR_CORRUPT;
    size_t u;
    u = seq_table_setup(&fc->ll, (modes >> 6) & 3, p, left, LL_default, 36, 6, 35, 9); if (u == (size_t)-1) return ZERR_CORRUPT; p += u; left -= u;
    u = seq_table_setup(&fc->of, (modes >> 4) & 3, p, left, OF_default, 29, 5, 31, 8); if (u == (size_t)-1) return ZERR_CORRUPT; p += u; left -= u;
    u = seq_table_setup(&fc->ml, (modes >> 2) & 3, p, left, ML_default, 53, 6, 52, 9); if (u == (size_t)-1) return ZERR_CORRUPT; p += u; left -= u;
    bwd_bits b;
    if (bwd_init(&b, p, left)) return ZERR_CORRUPT;
    uint32_t sl = bwd_read(&b, fc->ll.log), so = bwd_read(&b, fc->of.log), sm = bwd_read(&b, fc->ml.log);
    for (size_t i = 0; i < nseq; ++i) {
        int llc = fc->ll.t[sl].symbol, ofc = fc->of.t[so].symbol, mlc = fc->ml.t[sm].symbol;
        if (llc > 35 || mlc > 52 || ofc > 31) return ZERR_CORRUPT;
        uint32_t ofv = (1u << ofc) + bwd_read(&b, ofc);       /* order: offset, match length, literal length */
        uint32_t ml = ML_base[mlc] + bwd_read(&b, ML_bits[mlc]);
        uint32_t ll = LL_base[llc] + bwd_read(&b, LL_bits[llc]);
        uint32_t offset;
        if (ofv > 3) { offset = ofv - 3; fc->rep[2] = fc->rep[1]; fc->rep[1] = fc->rep[0]; fc->rep[0] = offset; }
        else {
            uint32_t idx = ofv - 1 + (ll == 0);
            if (idx == 0) offset = fc->rep[0];
            else {
                offset = idx < 3 ? fc->rep[idx] : fc->rep[0] - 1;
                if (offset == 0) return ZERR_CORRUPT;
                if (idx > 1) fc->rep[2] = fc->rep[1];
                fc->rep[1] = fc->rep[0]; fc->rep[0] = offset;
            }
        }
        if (i + 1 < nseq) {                                   /* state update order: LL, ML, OF */
            sl = fc->ll.t[sl].newState + bwd_read(&b, fc->ll.t[sl].nbBits);
            sm = fc->ml.t[sm].newState + bwd_read(&b, fc->ml.t[sm].nbBits);
            so = fc->of.t[so].newState + bwd_read(&b, fc->of.t[so].nbBits);
        }
        if (b.overflow) return ZERR_CORRUPT;
        if (lp + ll > regen) return ZERR_CORRUPT;
        if (op + ll + ml > cap) return ZERR_DST_SIZE;
        memcpy(out + op, lit + lp, ll); op += ll; lp += ll;
        if (offset > dst_pos + op) return ZERR_CORRUPT;
        for (uint32_t k = 0; k < ml; ++k) out[op + k] = out[op + k - offset];   /* overlapping copy */
        op += ml;
    }
    if (b.bitpos != 0) return ZERR_CORRUPT;
    size_t rest = regen - lp;
    if (op + rest > cap) return ZERR_DST_SIZE;
    memcpy(out + op, lit + lp, rest); op += rest;
    *produced = op;
    return 0;
}

/* Frame header. Returns 0 and fills *header_size, *fcs (or (uint64_t)-1 when unknown), *has_checksum. */
static int parse_frame_header(const uint8_t* src, size_t size, size_t* header_size, uint64_t* fcs, int* has_checksum)
{
    if (size < 5) return ZERR_SRC_SIZE;
    if (!(src[0] == 0x28 && src[1] == 0xB5 && src[2] == 0x2F && src[3] == 0xFD)) return ZERR_MAGIC;
    int fhd = src[4], fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did = fhd & 3;
    if (fhd & 0x08) return ZERR_UNSUPPORTED;                  /* reserved bit */
    size_t pos = 5;
    if (!single) pos += 1;                                    /* window descriptor (unused: dst is contiguous) */
    static const int did_sz[4] = {0, 1, 2, 4};
    if (did) { for (int i = 0; i < did_sz[did]; ++i) if (pos + i < size && src[pos + i]) return ZERR_UNSUPPORTED; pos += did_sz[did]; }
    int fcs_sz = fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
    if (pos + fcs_sz > size) return ZERR_SRC_SIZE;
    uint64_t v = 0; for (int i = 0; i < fcs_sz; ++i) v |= (uint64_t)src[pos + i] << (8 * i);
    if (fcs_sz == 2) v += 256;
    *fcs = fcs_sz ? v : (uint64_t)-1;
    *header_size = pos + fcs_sz; *has_checksum = (fhd >> 2) & 1;
    return 0;
}

/* ZSTD_getFrameContentSize restatement: returns content size, (uint64_t)-1 unknown, (uint64_t)-2 error. */
uint64_t orc_zstd_frame_content_size(const uint8_t* src, size_t size)
{
    size_t hs; uint64_t fcs; int ck;
    if (parse_frame_header(src, size, &hs, &fcs, &ck)) return (uint64_t)-2;
    return fcs;
}

/* ZSTD_decompress restatement for exactly one frame occupying [src, src+size). 0 on success. */
int orc_zstd_decode_frame(const uint8_t* src, size_t size, uint8_t* dst, size_t cap, size_t* out_size)
{
    size_t pos; uint64_t fcs; int ck;
    int r = parse_frame_header(src, size, &pos, &fcs, &ck);
    if (r) return r;
    frame_ctx* fc = (frame_ctx*)calloc(1, sizeof(frame_ctx));
    uint8_t* litbuf = (uint8_t*)malloc(128 * 1024 + 32);
    fc->rep[0] = 1; fc->rep[1] = 4; fc->rep[2] = 8;
    size_t op = 0; int last = 0;
    while (!last) {
        if (pos + 3 > size) { r = ZERR_SRC_SIZE; break; }
        uint32_t bh = src[pos] | (src[pos + 1] << 8) | ((uint32_t)src[pos + 2] << 16);
        pos += 3; last = bh & 1; int type = (bh >> 1) & 3; size_t bsz = bh >> 3;
        if (type == 0) { if (pos + bsz > size) { r = ZERR_SRC_SIZE; break; } if (op + bsz > cap) { r = ZERR_DST_SIZE; break; }
                         memcpy(dst + op, src + pos, bsz); op += bsz; pos += bsz; }
        else if (type == 1) { if (pos + 1 > size) { r = ZERR_SRC_SIZE; break; } if (op + bsz > cap) { r = ZERR_DST_SIZE; break; }
                              memset(dst + op, src[pos], bsz); op += bsz; pos += 1; }
        else if (type == 2) { if (pos + bsz > size) { r = ZERR_SRC_SIZE; break; } size_t prod = 0;
                              r = decode_block(fc, src + pos, bsz, dst, op, cap, &prod, litbuf); if (r) break; op += prod; pos += bsz; }
        else { r = ZERR_CORRUPT; break; }
    }
    if (!r) { if (ck) pos += 4; if (pos != size) r = ZERR_CORRUPT; }
    if (!r && fcs != (uint64_t)-1 && fcs != op) r = ZERR_CORRUPT;
    free(fc); free(litbuf);
    *out_size = op;
    return r;
}
