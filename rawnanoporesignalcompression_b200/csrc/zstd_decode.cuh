// zstd_decode.cuh -- CTA-cooperative Zstandard frame decoder (RFC 8878) for sm_100a.
//
// Replaces the reference's third-party calls ZSTD_getFrameContentSize + ZSTD_decompress
// (pod5_format/signal_compression.cpp:103,117; pgnano/svb16/C5.hpp:494-565,588-655 and siblings).
// The executable specification is oracle/zstd_decode_oracle.c (pinned against libzstd 1.5.5).
//
// One CTA decodes one frame at a time. What is parallel:
//   * Huffman literals: every one of the 1 or 4 backward bitstreams is cut into NT/nstreams
//     sub-sequences; each thread decodes one sub-sequence starting WARM bits early so that it has
//     self-synchronised by the time it reaches its own range; the chain start[k+1] == end[k] is then
//     verified and repaired (rarely needed), a CTA scan of the symbol counts assigns output offsets and
//     a second pass writes the symbols. Decode table (<= 2^12 x u16) lives in shared memory.
//   * table construction, raw/RLE copies: strided over the CTA.
// What is serial (one thread / one warp): header parsing, FSE weight decoding, FSE sequence decoding and
// the in-order execution of sequences (warp 0, lane-parallel byte copies).
#pragma once
#include "block_prims.cuh"

namespace rnsc {

constexpr int HUF_LOG_MAX = 12;
constexpr uint32_t HUF_WARM_BITS = 192;   // warm-up window before a sub-sequence (self-synchronisation)
constexpr uint32_t HUF_SUB_MIN_BITS = 256; // minimum sub-sequence length
constexpr int SEQ_BATCH = 128;
constexpr uint32_t ZSTD_BLOCK_MAX = 128 * 1024;

struct ZstdWs {
    uint16_t huf[1 << HUF_LOG_MAX];      // (nbBits << 8) | symbol
    uint32_t fse_ll[512], fse_ml[512], fse_of[256];  // (newState << 16) | (nbBits << 8) | symbol
    uint32_t sub_start[NT], sub_end[NT], sub_cnt[NT];
    uint32_t seq_ll[SEQ_BATCH], seq_ml[SEQ_BATCH], seq_of[SEQ_BATCH];
    uint32_t scan_a[NWARP + 1], scan_b[NWARP + 1];
    uint32_t rank_cnt[16], rank_start[16];
    uint32_t fse_w[64];
    int16_t norm[64];
    uint16_t next16[64];
    uint8_t weights[256];
    uint8_t sym_rank[256];
    // frame / block state (written by thread 0, read by all after a barrier)
    uint32_t rep[3];
    int huf_log, huf_valid, ll_log, ll_valid, of_log, of_valid, ml_log, ml_valid;
    int err, berr;
    uint32_t pos, fcs_lo, has_ck, fcs_known;
    uint32_t b_last, b_type, b_size;
    uint32_t ltype, regen, lcomp, nstreams, lhs, nseq, seq_modes, seq_off;
    uint32_t s_beg[4], s_size[4], s_count[4], s_dst[4], s_tbits[4], s_sub[4], s_nsub[4];
    uint32_t lit_is_dst;
    uint32_t out_pos;
    uint32_t seq_ll_n, seq_done;
};

// ---------------------------------------------------------------------------------------------------
// single-thread helpers
// ---------------------------------------------------------------------------------------------------
__device__ inline int highbit32(uint32_t v) { return 31 - __clz((int)v); }

// LSB-first forward reader (FSE table descriptions)
struct FwdBits { const uint8_t* p; uint32_t size; uint32_t bitpos; };
// 32 bits starting at bit `bit` of the byte string p: the two aligned 32-bit words that cover them, funnel
// shifted. Reads at most 3 bytes below and 7 bytes above the bytes needed (inside the chunk or its padding).
__device__ __forceinline__ uint32_t gbits32(const uint8_t* p, uint32_t bit)
{
    const uintptr_t a = (uintptr_t)p + (bit >> 3), al = a & ~(uintptr_t)3;
    const uint32_t sh = ((uint32_t)(a & 3) << 3) + (bit & 7);
    return __funnelshift_r(ldg32((const uint32_t*)al), ldg32((const uint32_t*)(al + 4)), sh);
}
__device__ inline uint32_t fwd_peek(const FwdBits& b, int n)
{
    const uint32_t total = b.size * 8;
    if (b.bitpos >= total) return 0;
    uint32_t v = gbits32(b.p, b.bitpos);
    const uint32_t avail = total - b.bitpos;                  // bits past the end of the section read as zero
    if (avail < 32) v &= (1u << avail) - 1u;
    return v & ((1u << n) - 1u);                              // n <= 10 here
}

// FSE_readNCount. Returns bytes consumed, 0 on error.
__device__ inline uint32_t fse_read_ncount(const uint8_t* src, uint32_t size, int16_t* norm, int max_sym, int max_log, int* table_log, int* nsym)
{
    FwdBits b{src, size, 0};
    if (size < 1) return 0;
    int tl = (int)fwd_peek(b, 4) + 5; b.bitpos += 4;
    if (tl > max_log) return 0;
    int remaining = (1 << tl) + 1, threshold = 1 << tl, nbBits = tl + 1, sym = 0, prev0 = 0;
    while (remaining > 1 && sym <= max_sym) {
        if (prev0) {
            int n0 = sym;
            for (;;) { uint32_t r = fwd_peek(b, 2); b.bitpos += 2; n0 += (int)r; if (r != 3) break; if (n0 > max_sym + 1) return 0; }
            if (n0 > max_sym + 1) return 0;
            while (sym < n0) norm[sym++] = 0;
            if (sym > max_sym) break;
        }
        int mx = (2 * threshold - 1) - remaining, count;
        uint32_t lo = fwd_peek(b, nbBits - 1);
        if ((int)lo < mx) { count = (int)lo; b.bitpos += nbBits - 1; }
        else { count = (int)fwd_peek(b, nbBits); if (count >= threshold) count -= mx; b.bitpos += nbBits; }
        count--;
        remaining -= count < 0 ? -count : count;
        norm[sym++] = (int16_t)count;
        prev0 = (count == 0);
        while (remaining < threshold) { nbBits--; threshold >>= 1; }
    }
    if (remaining != 1) return 0;
    uint32_t used = (b.bitpos + 7) >> 3;
    if (used > size) return 0;
    *table_log = tl; *nsym = sym;
    return used;
}

// FSE_buildDTable. t[u] = (newState << 16) | (nbBits << 8) | symbol. `next` = 64 x u16 of scratch.
__device__ inline void fse_build(uint32_t* t, const int16_t* norm, int nsym, int tl, uint16_t* next)
{
    int size = 1 << tl, high = size - 1;
    for (int s = 0; s < nsym; ++s) {
        if (norm[s] == -1) { t[high--] = (uint32_t)s; next[s] = 1; }
        else next[s] = (uint16_t)norm[s];
    }
    int step = (size >> 1) + (size >> 3) + 3, mask = size - 1, pos = 0;
    for (int s = 0; s < nsym; ++s)
        for (int i = 0; i < norm[s]; ++i) {
            t[pos] = (uint32_t)s;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    for (int u = 0; u < size; ++u) {
        uint32_t s = t[u] & 0xFF;
        uint32_t ns = next[s]++;
        int nb = tl - highbit32(ns);
        t[u] = (((ns << nb) - (uint32_t)size) << 16) | ((uint32_t)nb << 8) | s;
    }
}

// MSB-first backward reader over global bytes (FSE streams; serial use only). The two aligned words that cover the
// current position stay in registers: a read is a compare, a funnel shift and a mask, and one word is fetched when the
// position crosses a word boundary (every 32 bits) instead of two per read.
struct BwdBits { const uint8_t* start; int bitpos; int overflow; uintptr_t wa; uint32_t w0, w1; };
__device__ inline int bwd_init(BwdBits& b, const uint8_t* p, uint32_t size)
{
    if (size == 0) return -1;
    uint32_t last = ldg8(p + size - 1);
    if (last == 0) return -1;
    b.start = p; b.overflow = 0; b.wa = 0; b.w0 = b.w1 = 0;
    b.bitpos = (int)size * 8 - (8 - highbit32(last));
    return 0;
}
__device__ __forceinline__ uint32_t bwd_bits32(BwdBits& b, uint32_t bit)      // gbits32(b.start, bit) through the cached words
{
    const uintptr_t a = (uintptr_t)b.start + (bit >> 3), al = a & ~(uintptr_t)3;
    if (al != b.wa) {
        if (al + 4 == b.wa) {
            b.w1 = b.w0; b.w0 = ldg32((const uint32_t*)al);
            // the stream is read towards its start: entering a sector, ask for the one below it (never below the stream's own first sector)
            if ((al & 28u) == 28u) { const uintptr_t lowest = (uintptr_t)b.start & ~(uintptr_t)31; if (al - 28 > lowest) prefetch_sector((const void*)(al - 60)); }
        }
        else { b.w0 = ldg32((const uint32_t*)al); b.w1 = ldg32((const uint32_t*)(al + 4)); }
        b.wa = al;
    }
    return __funnelshift_r(b.w0, b.w1, ((uint32_t)(a & 3) << 3) + (bit & 7));
}
__device__ inline uint32_t bwd_read(BwdBits& b, int n)
{
    if (n == 0) return 0;
    const int lo = b.bitpos - n;
    b.bitpos = lo;
    if (lo >= 0) return bwd_bits32(b, (uint32_t)lo) & (0xFFFFFFFFu >> (32 - n));     // n <= 31
    b.overflow = 1;                                           // bits below the start of the stream read as zero
    const int avail = n + lo;
    if (avail <= 0) return 0;
    return ((gbits32(b.start, 0) & (0xFFFFFFFFu >> (32 - avail))) << (-lo)) & (0xFFFFFFFFu >> (32 - n));
}

// ---------------------------------------------------------------------------------------------------
// Huffman: tree description (thread 0), table build (CTA), parallel stream decode (CTA)
// ---------------------------------------------------------------------------------------------------
// Warp reconvergence point for code that runs one independent problem per LANE (the planner): `SYNC` instantiations
// must be executed by all lanes of the warp the same number of times (single-exit, flag-driven control flow), so
// that the long loops that follow run converged instead of once per diverged lane.
template <bool SYNC> __device__ __forceinline__ void warp_converge() { if (SYNC) __syncwarp(); }

// Reads the Huffman tree description into ws->weights; returns bytes consumed (0 = error or size == 0) and *nw.
// Single exit; with SYNC every lane of the warp must call it (pass size = 0 to skip the work).
template <typename WS, bool SYNC = false>
__device__ inline uint32_t huf_read_weights(WS* ws, const uint8_t* src, uint32_t size, int* nw_out)
{
    uint8_t* w = ws->weights; int nw = 0;
    bool ok = size >= 1;
    const uint32_t hb = ok ? ldg8(src) : 0;
    const bool direct = hb >= 128;
    uint32_t used = 0;
    // ---- direct representation: one 4-bit weight per symbol ----
    if (ok && direct) { nw = (int)hb - 127; used = 1 + (uint32_t)((nw + 1) / 2); if (used > size) ok = false; }
    {
        const int lim = (ok && direct) ? nw : 0;
        for (int i = 0; i < lim; ++i) { const uint32_t by = ldg8(src + 1 + i / 2); w[i] = (uint8_t)((i & 1) ? (by & 15) : (by >> 4)); }
    }
    // ---- FSE-compressed weights ----
    int tl = 0, ns = 0; uint32_t h = 0;
    if (ok && !direct) { used = 1 + hb; if (used > size || hb < 1) ok = false; }
    if (ok && !direct) { h = fse_read_ncount(src + 1, hb, ws->norm, 12, 6, &tl, &ns); if (!h) ok = false; }
    warp_converge<SYNC>();
    if (ok && !direct) fse_build(ws->fse_w, ws->norm, ns, tl, ws->next16);
    warp_converge<SYNC>();
    if (ok && !direct) {
        BwdBits b;
        if (bwd_init(b, src + 1 + h, hb - h)) ok = false;
        else {
            uint32_t s1 = bwd_read(b, tl), s2 = bwd_read(b, tl);
            for (;;) {
                if (nw >= 254) { ok = false; break; }
                uint32_t e = ws->fse_w[s1]; w[nw++] = (uint8_t)(e & 0xFF);
                s1 = (e >> 16) + bwd_read(b, (int)((e >> 8) & 0xFF));
                if (b.overflow) { if (nw >= 255) ok = false; else w[nw++] = (uint8_t)(ws->fse_w[s2] & 0xFF); break; }
                if (nw >= 254) { ok = false; break; }
                e = ws->fse_w[s2]; w[nw++] = (uint8_t)(e & 0xFF);
                s2 = (e >> 16) + bwd_read(b, (int)((e >> 8) & 0xFF));
                if (b.overflow) { if (nw >= 255) ok = false; else w[nw++] = (uint8_t)(ws->fse_w[s1] & 0xFF); break; }
            }
        }
    }
    warp_converge<SYNC>();
    // ---- one pass: validity, total weight and the per-weight symbol counts. The counters are 8-bit fields of two
    //      64-bit registers (at most 255 weights are listed): no memory read-modify-write chain, and no local array
    //      (a data-dependent index would be expanded into a compare/select chain per symbol). Left in ws->rank_cnt. ----
    uint32_t* cnt = ws->rank_cnt;
    uint32_t total = 0;
    if (ok) {
        unsigned long long c_lo = 0, c_hi = 0;              // weights 0-7, 8-15
        bool big = false;
        for (int i = 0; i < nw; ++i) {
            uint32_t wi = w[i]; if (wi > HUF_LOG_MAX) { big = true; wi = 0; }
            total += (1u << wi) >> 1;
            const unsigned long long one = 1ull << ((wi & 7u) * 8u);
            if (wi & 8u) c_hi += one; else c_lo += one;
        }
        for (int k = 0; k < 8; ++k) { cnt[k] = (uint32_t)(c_lo >> (8 * k)) & 0xFFu; cnt[8 + k] = (uint32_t)(c_hi >> (8 * k)) & 0xFFu; }
        if (big || total == 0) ok = false;
    }
    int maxBits = 0;
    if (ok) {
        maxBits = highbit32(total) + 1;
        const uint32_t rest = (1u << maxBits) - total;
        if (maxBits > HUF_LOG_MAX || (rest & (rest - 1))) ok = false;
        else {
            const uint32_t lastw = (uint32_t)highbit32(rest) + 1;
            w[nw++] = (uint8_t)lastw; cnt[lastw]++;
            if (cnt[1] < 2 || (cnt[1] & 1)) ok = false;
        }
    }
    warp_converge<SYNC>();
    if (ok) {
        for (int i = nw; i < 256; ++i) w[i] = 0;
        // rank starts: weight ascending == code length descending (HUF_readDTableX1 layout)
        uint32_t p = 0;
        for (int k = 1; k <= maxBits; ++k) { ws->rank_start[k] = p; p += cnt[k] << (k - 1); }
        ws->huf_log = maxBits;
        *nw_out = nw;
    }
    return ok ? used : 0;
}

// CTA: fill ws->huf from ws->weights / ws->rank_start. Symbols of equal weight take consecutive runs of
// 2^(w-1) entries in natural symbol order.
__device__ inline void huf_build_table(ZstdWs* ws)
{
    static_assert(NT == 512, "huf_build_table assigns two threads per symbol");
    const int tid = threadIdx.x;
    const int maxBits = ws->huf_log;
    if (tid < 256) {
        uint32_t w = ws->weights[tid];
        uint32_t r = 0;
        if (w) for (int j = 0; j < tid; ++j) r += (ws->weights[j] == w);
        ws->sym_rank[tid] = (uint8_t)r;   // rank < 256 always fits (at most 255 earlier symbols)
    }
    __syncthreads();
    // two threads per symbol split the run; long runs (w large) are at most 2048 entries
    {
        int sym = tid >> 1, half = tid & 1;
        uint32_t w = ws->weights[sym];
        if (w) {
            uint32_t len = 1u << (w - 1), st = ws->rank_start[w] + (uint32_t)ws->sym_rank[sym] * len;
            uint16_t e = (uint16_t)(((uint32_t)(maxBits + 1 - w) << 8) | (uint32_t)sym);
            uint32_t b = half ? (len + 1) / 2 : 0, en = half ? len : (len + 1) / 2;
            for (uint32_t u = b; u < en; ++u) ws->huf[st + u] = e;
        }
    }
    __syncthreads();
}

// Per-thread bit window over a backward stream. Absolute bit index 0 = bit 0 of the 32-bit word at
// `words` (a 4-byte aligned address <= stream begin, moved one word down so that a >= -32 never indexes
// below it). `a` = absolute index of the lowest bit of the current 32-bit view.
struct HufCursor {
    const uint32_t* words; int q; int r; uint32_t lo, hi;
    __device__ __forceinline__ void init(const uint32_t* w, int a) { words = w; q = a >> 5; r = a & 31; lo = ldg32(words + q); hi = ldg32(words + q + 1); }
    __device__ __forceinline__ uint32_t view() const { return __funnelshift_r(lo, hi, (uint32_t)r); }
    __device__ __forceinline__ void skip(int nb) { r -= nb; if (r < 0) { r += 32; q -= 1; hi = lo; lo = ldg32(words + (q < 0 ? 0 : q)); } }
};

// CTA: decode the streams described in ws->s_* into `dst` (generic pointer). Returns 0 on success
// (same value in every thread).
__device__ inline int huf_decode_parallel(ZstdWs* ws, const uint8_t* src_base, uint8_t* dst)
{
    const int tid = threadIdx.x;
    const int ns = (int)ws->nstreams;
    const int tps = NT / ns;                 // threads per stream
    const int j = tid / tps, k = tid % tps;  // stream, sub-sequence
    const int L = ws->huf_log;
    const uint32_t T = ws->s_tbits[j], sub = ws->s_sub[j], nsub = ws->s_nsub[j];
    const bool active = (uint32_t)k < nsub;
    // stream geometry
    const uint8_t* sb = src_base + ws->s_beg[j];
    const uintptr_t al = ((uintptr_t)sb & ~(uintptr_t)3) - 4;        // one word below the aligned begin
    const uint32_t* words = (const uint32_t*)al;
    const int bit0 = 8 * (int)((uintptr_t)sb - al);                  // absolute index of the stream's bit 0 (32..56)
    const int mark = bit0 + (int)T;                                  // absolute index of the end-mark bit
    uint32_t my_start = 0, my_end = 0, my_cnt = 0;
    if (active) {
        const uint32_t s0 = (uint32_t)k * sub;
        uint32_t p = s0 > HUF_WARM_BITS ? s0 - HUF_WARM_BITS : 0;
        HufCursor c; c.init(words, mark - (int)p - 32);
        while (p < s0) { uint32_t e = ws->huf[c.view() >> (32 - L)]; int nb = (int)(e >> 8); c.skip(nb); p += (uint32_t)nb; }
        my_start = p;
        const uint32_t lim = min(s0 + sub, T);
        while (p < lim) { uint32_t e = ws->huf[c.view() >> (32 - L)]; int nb = (int)(e >> 8); c.skip(nb); p += (uint32_t)nb; my_cnt++; }
        my_end = p;
    }
    ws->sub_start[tid] = my_start; ws->sub_end[tid] = my_end; ws->sub_cnt[tid] = my_cnt;
    __syncthreads();
    // verify / repair the chain: start[k] must equal end[k-1]
    for (int round = 0; round < NT; ++round) {
        bool bad = active && k > 0 && ws->sub_end[tid - 1] != my_start;
        uint32_t want = bad ? ws->sub_end[tid - 1] : 0;
        if (!__syncthreads_or(bad)) break;
        if (bad) {
            const uint32_t s0 = (uint32_t)k * sub, lim = min(s0 + sub, T);
            uint32_t p = want; my_start = p; my_cnt = 0;
            if (p < lim) {
                HufCursor c; c.init(words, mark - (int)p - 32);
                while (p < lim) { uint32_t e = ws->huf[c.view() >> (32 - L)]; int nb = (int)(e >> 8); c.skip(nb); p += (uint32_t)nb; my_cnt++; }
            }
            my_end = p;
            ws->sub_start[tid] = my_start; ws->sub_end[tid] = my_end; ws->sub_cnt[tid] = my_cnt;
        }
        __syncthreads();
    }
    // output offsets: exclusive scan of symbol counts within each stream (one CTA scan, stream base removed)
    uint32_t total;
    uint32_t ex = block_exclusive_scan<uint32_t>(active ? my_cnt : 0u, ws->scan_a, &total);
    if (k == 0) ws->s_dst[j] = ex;           // scan value at the first thread of stream j (tps threads per stream)
    __syncthreads();
    const uint32_t off = ex - ws->s_dst[j];
    // checks: every stream consumed exactly and produced exactly its share
    bool bad = false;
    if (active && (uint32_t)k == nsub - 1) bad = (my_end != T) || (off + my_cnt != ws->s_count[j]);
    if (k == 0 && nsub == 0) bad = ws->s_count[j] != 0;
    if (__syncthreads_or(bad)) return 1;
    if (active && my_cnt) {
        uint8_t* o = dst + (size_t)j * ws->s_count[0] + off;   // streams 0..2 hold s_count[0] symbols each
        HufCursor c; c.init(words, mark - (int)my_start - 32);
        for (uint32_t i = 0; i < my_cnt; ++i) { uint32_t e = ws->huf[c.view() >> (32 - L)]; c.skip((int)(e >> 8)); o[i] = (uint8_t)e; }
    }
    __syncthreads();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// sequences
// ---------------------------------------------------------------------------------------------------
__constant__ uint32_t kLLBase[36] = {0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,18,20,22,24,28,32,40,48,64,
                                     0x80,0x100,0x200,0x400,0x800,0x1000,0x2000,0x4000,0x8000,0x10000};
__constant__ uint8_t kLLBits[36] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,6,7,8,9,10,11,12,13,14,15,16};
__constant__ uint32_t kMLBase[53] = {3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,30,31,32,
                                     33,34,35,37,39,41,43,47,51,59,67,83,99,0x83,0x103,0x203,0x403,0x803,0x1003,0x2003,
                                     0x4003,0x8003,0x10003};
__constant__ uint8_t kMLBits[53] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,
                                    1,1,1,1,2,2,3,3,4,4,5,7,8,9,10,11,12,13,14,15,16};
__constant__ int16_t kLLDefault[36] = {4,3,2,2,2,2,2,2,2,2,2,2,2,1,1,1,2,2,2,2,2,2,2,2,2,3,2,1,1,1,1,1,-1,-1,-1,-1};
__constant__ int16_t kMLDefault[53] = {1,4,3,2,2,2,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,
                                       1,1,-1,-1,-1,-1,-1,-1,-1};
__constant__ int16_t kOFDefault[29] = {1,1,1,1,1,1,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1};

// thread 0: set up one of the three sequence tables. Returns bytes consumed or 0xFFFFFFFF on error.
__device__ inline uint32_t seq_table_setup(ZstdWs* ws, uint32_t* table, int* log, int* valid, int mode, const uint8_t* src, uint32_t size,
                                           const int16_t* defnorm, int defn, int deflog, int max_sym, int max_log)
{
    uint16_t* next = ws->next16;
    switch (mode) {
    case 0: for (int i = 0; i < defn; ++i) ws->norm[i] = defnorm[i];
            fse_build(table, ws->norm, defn, deflog, next); *log = deflog; *valid = 1; return 0;
    case 1: { if (size < 1) return 0xFFFFFFFFu; uint32_t s = ldg8(src); if ((int)s > max_sym) return 0xFFFFFFFFu;
              table[0] = s; *log = 0; *valid = 1; return 1; }
    case 2: { int tl, ns; uint32_t h = fse_read_ncount(src, size, ws->norm, max_sym, max_log, &tl, &ns);
              if (!h) return 0xFFFFFFFFu;
              fse_build(table, ws->norm, ns, tl, next); *log = tl; *valid = 1; return h; }
    default: return *valid ? 0 : 0xFFFFFFFFu;
    }
}

// ---------------------------------------------------------------------------------------------------
// block + frame (CTA)
// ---------------------------------------------------------------------------------------------------
// Cooperative byte copy / fill over generic pointers.
__device__ inline void cta_copy(uint8_t* d, const uint8_t* s, uint32_t n) { for (uint32_t i = threadIdx.x; i < n; i += NT) d[i] = s[i]; }
__device__ inline void cta_fill(uint8_t* d, uint32_t v, uint32_t n) { for (uint32_t i = threadIdx.x; i < n; i += NT) d[i] = (uint8_t)v; }

// Decode one compressed block. `dst` is the frame output (generic pointer), ws->out_pos the bytes already
// produced, `cap` the frame capacity, `lit` a CTA-private scratch of >= 128 KiB + 32 (global).
__device__ inline void zstd_decode_block(ZstdWs* ws, const uint8_t* src, uint32_t size, uint8_t* dst, uint32_t cap, uint8_t* lit)
{
    const int tid = threadIdx.x;
    if (tid == 0) {
        int err = 0;
        if (size < 1) err = 1;
        uint32_t b0 = err ? 0 : ldg8(src);
        uint32_t ltype = b0 & 3, sf = (b0 >> 2) & 3, lhs = 0, regen = 0, comp = 0, nstreams = 1;
        if (!err) {
            if (ltype < 2) {
                if (sf == 0 || sf == 2) { lhs = 1; regen = b0 >> 3; }
                else if (sf == 1) { if (size < 2) err = 1; else { lhs = 2; regen = (b0 >> 4) | (ldg8(src + 1) << 4); } }
                else { if (size < 3) err = 1; else { lhs = 3; regen = (b0 >> 4) | (ldg8(src + 1) << 4) | (ldg8(src + 2) << 12); } }
            } else {
                if (size < 3) err = 1;
                else {
                    uint64_t h = 0; for (uint32_t i = 0; i < 5 && i < size; ++i) h |= (uint64_t)ldg8(src + i) << (8 * i);
                    if (sf == 0 || sf == 1) { lhs = 3; regen = (uint32_t)(h >> 4) & 0x3FF; comp = (uint32_t)(h >> 14) & 0x3FF; nstreams = sf ? 4 : 1; }
                    else if (sf == 2) { lhs = 4; regen = (uint32_t)(h >> 4) & 0x3FFF; comp = (uint32_t)(h >> 18) & 0x3FFF; nstreams = 4; }
                    else { if (size < 5) err = 1; lhs = 5; regen = (uint32_t)(h >> 4) & 0x3FFFF; comp = (uint32_t)(h >> 22) & 0x3FFFF; nstreams = 4; }
                }
            }
        }
        if (!err && regen > ZSTD_BLOCK_MAX) err = 1;
        uint32_t lit_section = lhs + (ltype == 0 ? regen : (ltype == 1 ? 1 : comp));
        if (!err && lit_section > size) err = 1;
        // sequences header
        uint32_t nseq = 0, modes = 0, so = lit_section;
        if (!err) {
            if (so >= size) err = 1;
            else {
                uint32_t n0 = ldg8(src + so);
                if (n0 == 0) { nseq = 0; so += 1; if (so != size) err = 1; }
                else {
                    if (n0 < 128) { nseq = n0; so += 1; }
                    else if (n0 < 255) { if (so + 2 > size) err = 1; else { nseq = ((n0 - 128) << 8) + ldg8(src + so + 1); so += 2; } }
                    else { if (so + 3 > size) err = 1; else { nseq = ldg8(src + so + 1) + (ldg8(src + so + 2) << 8) + 0x7F00; so += 3; } }
                    if (!err) { if (so >= size) err = 1; else { modes = ldg8(src + so); so += 1; if (modes & 3) err = 1; } }
                }
            }
        }
        ws->ltype = ltype; ws->regen = regen; ws->lcomp = comp; ws->nstreams = nstreams; ws->lhs = lhs;
        ws->nseq = nseq; ws->seq_modes = modes; ws->seq_off = so;
        ws->lit_is_dst = (nseq == 0);
        if (!err && nseq == 0 && ws->out_pos + regen > cap) err = 1;
        // Huffman tree + stream geometry
        if (!err && ltype >= 2) {
            uint32_t p = lhs, left = comp;
            if (ltype == 2) { int nw; uint32_t u = huf_read_weights(ws, src + p, left, &nw); if (!u) err = 1; else { p += u; left -= u; ws->huf_valid = 1; } }
            else if (!ws->huf_valid) err = 1;
            if (!err) {
                uint32_t sz[4], cnt[4];
                if (nstreams == 1) { sz[0] = left; cnt[0] = regen; ws->s_beg[0] = p; }
                else {
                    if (left < 6) err = 1;
                    else {
                        sz[0] = ldg8(src + p) | (ldg8(src + p + 1) << 8); sz[1] = ldg8(src + p + 2) | (ldg8(src + p + 3) << 8);
                        sz[2] = ldg8(src + p + 4) | (ldg8(src + p + 5) << 8);
                        if (6 + sz[0] + sz[1] + sz[2] > left) err = 1;
                        else {
                            sz[3] = left - 6 - sz[0] - sz[1] - sz[2];
                            uint32_t q = (regen + 3) / 4;
                            if (3 * q > regen) err = 1;
                            cnt[0] = cnt[1] = cnt[2] = q; cnt[3] = regen - 3 * q;
                            ws->s_beg[0] = p + 6; ws->s_beg[1] = ws->s_beg[0] + sz[0]; ws->s_beg[2] = ws->s_beg[1] + sz[1]; ws->s_beg[3] = ws->s_beg[2] + sz[2];
                        }
                    }
                }
                const uint32_t tps = NT / nstreams;
                for (uint32_t j = 0; j < nstreams && !err; ++j) {
                    ws->s_size[j] = sz[j]; ws->s_count[j] = cnt[j];
                    if (sz[j] == 0) { err = 1; break; }
                    uint32_t last = ldg8(src + ws->s_beg[j] + sz[j] - 1);
                    if (last == 0) { err = 1; break; }
                    uint32_t T = 8 * (sz[j] - 1) + (uint32_t)highbit32(last);
                    uint32_t sub = (T + tps - 1) / tps; if (sub < HUF_SUB_MIN_BITS) sub = HUF_SUB_MIN_BITS;
                    ws->s_tbits[j] = T; ws->s_sub[j] = sub; ws->s_nsub[j] = (T + sub - 1) / sub;
                }
            }
        }
        ws->berr = err;   // not ws->err: the other threads may still be reading it (frame loop) before this barrier
    }
    __syncthreads();
    if (ws->berr) { if (tid == 0) ws->err = RNSC_ERR_ZSTD_DECODE; return; }
    const uint32_t regen = ws->regen, ltype = ws->ltype, nseq = ws->nseq;
    uint8_t* out = dst + ws->out_pos;
    // ---- literals ----
    const uint8_t* litp;
    if (ltype == 0) { litp = src + ws->lhs; if (nseq == 0) cta_copy(out, litp, regen); }
    else if (ltype == 1) { if (nseq == 0) cta_fill(out, ldg8(src + ws->lhs), regen); else cta_fill(lit, ldg8(src + ws->lhs), regen); litp = lit; }
    else {
        huf_build_table(ws);
        uint8_t* ldst = nseq == 0 ? out : lit;
        if (huf_decode_parallel(ws, src, ldst)) { if (tid == 0) ws->err = RNSC_ERR_ZSTD_DECODE; __syncthreads(); return; }
        litp = ldst;
    }
    __syncthreads();
    if (nseq == 0) { if (tid == 0) ws->out_pos += regen; __syncthreads(); return; }
    // ---- sequences: tables (thread 0), then warp 0 decodes (lane 0) and executes (all lanes) in order ----
    if (tid == 0) {
        uint32_t p = ws->seq_off, left = size - p, u; int err = 0; uint32_t modes = ws->seq_modes;
        u = seq_table_setup(ws, ws->fse_ll, &ws->ll_log, &ws->ll_valid, (modes >> 6) & 3, src + p, left, kLLDefault, 36, 6, 35, 9);
        if (u == 0xFFFFFFFFu) err = 1; else { p += u; left -= u; }
        if (!err) { u = seq_table_setup(ws, ws->fse_of, &ws->of_log, &ws->of_valid, (modes >> 4) & 3, src + p, left, kOFDefault, 29, 5, 31, 8);
                    if (u == 0xFFFFFFFFu) err = 1; else { p += u; left -= u; } }
        if (!err) { u = seq_table_setup(ws, ws->fse_ml, &ws->ml_log, &ws->ml_valid, (modes >> 2) & 3, src + p, left, kMLDefault, 53, 6, 52, 9);
                    if (u == 0xFFFFFFFFu) err = 1; else { p += u; left -= u; } }
        ws->seq_off = p;
        if (err) ws->err = RNSC_ERR_ZSTD_DECODE;
    }
    __syncthreads();
    if (ws->err) return;
    if (tid < 32) {
        const int lane = tid;
        BwdBits b; uint32_t sl = 0, so = 0, sm = 0; int err = 0;
        uint32_t rep0 = ws->rep[0], rep1 = ws->rep[1], rep2 = ws->rep[2];
        if (lane == 0) {
            if (bwd_init(b, src + ws->seq_off, size - ws->seq_off)) err = 1;
            else { sl = bwd_read(b, ws->ll_log); so = bwd_read(b, ws->of_log); sm = bwd_read(b, ws->ml_log); }
        }
        uint32_t op = ws->out_pos, lp = 0;
        const uint32_t avail = cap;   // frame capacity
        for (uint32_t base = 0; base < nseq; base += SEQ_BATCH) {
            const uint32_t nb = min((uint32_t)SEQ_BATCH, nseq - base);
            if (lane == 0 && !err) {
                for (uint32_t i = 0; i < nb; ++i) {
                    uint32_t el = ws->fse_ll[sl], eo = ws->fse_of[so], em = ws->fse_ml[sm];
                    uint32_t llc = el & 0xFF, ofc = eo & 0xFF, mlc = em & 0xFF;
                    if (llc > 35 || mlc > 52 || ofc > 31) { err = 1; break; }
                    uint32_t ofv = (1u << ofc) + bwd_read(b, (int)ofc);
                    uint32_t ml = kMLBase[mlc] + bwd_read(b, kMLBits[mlc]);
                    uint32_t ll = kLLBase[llc] + bwd_read(b, kLLBits[llc]);
                    uint32_t offset;
                    if (ofv > 3) { offset = ofv - 3; rep2 = rep1; rep1 = rep0; rep0 = offset; }
                    else {
                        uint32_t idx = ofv - 1 + (ll == 0);
                        if (idx == 0) offset = rep0;
                        else {
                            offset = idx == 1 ? rep1 : (idx == 2 ? rep2 : rep0 - 1);
                            if (offset == 0) { err = 1; break; }
                            if (idx > 1) rep2 = rep1;
                            rep1 = rep0; rep0 = offset;
                        }
                    }
                    if (base + i + 1 < nseq) {
                        sl = (el >> 16) + bwd_read(b, (int)((el >> 8) & 0xFF));
                        sm = (em >> 16) + bwd_read(b, (int)((em >> 8) & 0xFF));
                        so = (eo >> 16) + bwd_read(b, (int)((eo >> 8) & 0xFF));
                    }
                    if (b.overflow) { err = 1; break; }
                    ws->seq_ll[i] = ll; ws->seq_ml[i] = ml; ws->seq_of[i] = offset;
                }
                if (!err && base + nb == nseq && b.bitpos != 0) err = 1;
            }
            err = __shfl_sync(0xffffffffu, err, 0);
            if (err) break;
            __syncwarp();
            for (uint32_t i = 0; i < nb; ++i) {
                const uint32_t ll = ws->seq_ll[i], ml = ws->seq_ml[i], of = ws->seq_of[i];
                if (lp + ll > regen || op + ll + ml > avail || of > op + ll) { err = 1; break; }
                for (uint32_t k = lane; k < ll; k += 32) dst[op + k] = litp[lp + k];
                op += ll; lp += ll;
                __syncwarp();
                if (of >= 32) {
                    for (uint32_t k0 = 0; k0 < ml; k0 += 32) { uint32_t k = k0 + lane; if (k < ml) dst[op + k] = dst[op + k - of]; __syncwarp(); }
                } else {
                    const uint8_t* sp = dst + op - of;
                    for (uint32_t k = lane; k < ml; k += 32) dst[op + k] = sp[k % of];
                }
                op += ml;
                __syncwarp();
            }
            if (err) break;
        }
        if (!err) {
            const uint32_t rest = regen - lp;
            if (op + rest > avail) err = 1;
            else { for (uint32_t k = lane; k < rest; k += 32) dst[op + k] = litp[lp + k]; op += rest; }
        }
        if (lane == 0) {
            ws->rep[0] = rep0; ws->rep[1] = rep1; ws->rep[2] = rep2; ws->out_pos = op;
            if (err) ws->err = RNSC_ERR_ZSTD_DECODE;
        }
    }
    __syncthreads();
}

// Frame content size of the frame at src (thread-local helper). Returns 0 and fills fields, or an rnsc_status.
__device__ inline int zstd_parse_frame_header(const uint8_t* src, uint32_t size, uint32_t* hdr, uint64_t* fcs, uint32_t* known, uint32_t* ck)
{
    if (size < 5) return RNSC_ERR_NOT_ZSTD;
    if (!(ldg8(src) == 0x28 && ldg8(src + 1) == 0xB5 && ldg8(src + 2) == 0x2F && ldg8(src + 3) == 0xFD)) return RNSC_ERR_NOT_ZSTD;
    uint32_t fhd = ldg8(src + 4), flag = fhd >> 6, single = (fhd >> 5) & 1, did = fhd & 3;
    if (fhd & 0x08) return RNSC_ERR_NOT_ZSTD;
    uint32_t pos = 5;
    if (!single) pos += 1;
    const uint32_t did_sz = did == 3 ? 4 : did;
    if (pos + did_sz > size) return RNSC_ERR_NOT_ZSTD;
    for (uint32_t i = 0; i < did_sz; ++i) if (ldg8(src + pos + i)) return RNSC_ERR_ZSTD_DECODE;   // dictionaries unsupported
    pos += did_sz;
    uint32_t fsz = flag == 0 ? (single ? 1 : 0) : (flag == 1 ? 2 : (flag == 2 ? 4 : 8));
    if (pos + fsz > size) return RNSC_ERR_NOT_ZSTD;
    uint64_t v = 0; for (uint32_t i = 0; i < fsz; ++i) v |= (uint64_t)ldg8(src + pos + i) << (8 * i);
    if (fsz == 2) v += 256;
    *fcs = v; *known = fsz != 0; *hdr = pos + fsz; *ck = (fhd >> 2) & 1;
    return 0;
}

// CTA: decode the frame [src, src+size) into dst[0..cap). All threads return the same status; *produced
// (shared memory word) receives the decoded size. ws->err must be 0 on entry.
__device__ inline int zstd_decode_frame(ZstdWs* ws, const uint8_t* src, uint32_t size, uint8_t* dst, uint32_t cap, uint8_t* lit)
{
    const int tid = threadIdx.x;
    if (tid == 0) {
        uint32_t hdr = 0, known = 0, ck = 0; uint64_t fcs = 0;
        int r = zstd_parse_frame_header(src, size, &hdr, &fcs, &known, &ck);
        ws->err = r; ws->pos = hdr; ws->has_ck = ck; ws->fcs_known = known; ws->fcs_lo = (uint32_t)(fcs > 0xFFFFFFFFull ? 0xFFFFFFFFu : fcs);
        ws->rep[0] = 1; ws->rep[1] = 4; ws->rep[2] = 8; ws->huf_valid = 0; ws->ll_valid = ws->of_valid = ws->ml_valid = 0;
        ws->out_pos = 0; ws->b_last = 0;
        if (!r && known && fcs > cap) ws->err = RNSC_ERR_ZSTD_DECODE;
    }
    __syncthreads();
    while (!ws->err && !ws->b_last) {
        __syncthreads();
        if (tid == 0) {
            uint32_t pos = ws->pos;
            if (pos + 3 > size) ws->err = RNSC_ERR_ZSTD_DECODE;
            else {
                uint32_t bh = ldg8(src + pos) | (ldg8(src + pos + 1) << 8) | (ldg8(src + pos + 2) << 16);
                ws->b_last = bh & 1; ws->b_type = (bh >> 1) & 3; ws->b_size = bh >> 3; pos += 3;
                uint32_t need = ws->b_type == 1 ? 1u : ws->b_size;
                if (ws->b_type == 3 || pos + need > size) ws->err = RNSC_ERR_ZSTD_DECODE;
                else if (ws->b_type < 2 && ws->out_pos + ws->b_size > cap) ws->err = RNSC_ERR_ZSTD_DECODE;
                ws->pos = pos;
            }
        }
        __syncthreads();
        if (ws->err) break;
        const uint32_t pos = ws->pos, bsz = ws->b_size, type = ws->b_type;
        if (type == 0) { cta_copy(dst + ws->out_pos, src + pos, bsz); __syncthreads(); if (tid == 0) { ws->out_pos += bsz; ws->pos = pos + bsz; } }
        else if (type == 1) { cta_fill(dst + ws->out_pos, ldg8(src + pos), bsz); __syncthreads(); if (tid == 0) { ws->out_pos += bsz; ws->pos = pos + 1; } }
        else { zstd_decode_block(ws, src + pos, bsz, dst, cap, lit); if (tid == 0) ws->pos = pos + bsz; }
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0 && !ws->err) {
        uint32_t pos = ws->pos + (ws->has_ck ? 4 : 0);
        if (pos != size) ws->err = RNSC_ERR_ZSTD_DECODE;
        else if (ws->fcs_known && ws->fcs_lo != ws->out_pos) ws->err = RNSC_ERR_ZSTD_DECODE;
    }
    __syncthreads();
    return ws->err;
}

}  // namespace rnsc
