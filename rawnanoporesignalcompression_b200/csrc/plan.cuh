// plan.cuh -- the PLANNER: everything in a chunk's zstd frames that is inherently serial, run as one GPU
// thread per chunk so that the serial chains of ~10^5 chunks overlap each other instead of stalling a CTA.
//
// Per chunk it walks the pgnano container (A.1-A.8: u64 size prefixes) and, per frame, the zstd frame /
// block / literals / sequences headers (RFC 8878 3.1.1), decodes the Huffman tree description (direct or
// FSE-compressed weights) into a table *recipe*, and decodes the (few) sequences of the frame when their
// tables are predefined or RLE. The result is a fixed-size ChunkPlan in global memory that the fast decode
// kernel (fast_decode.cuh) consumes with no serial work left. Anything unusual (multi-block frames,
// treeless literals, more Huffman sections than the plan holds, inconsistent sizes) is not an error
// here: the chunk is routed to the general CTA-cooperative decoder (zstd_decode.cuh, k_decode_chunks), which
// owns the exact error semantics. Replaces, with the fast kernel, the reference's ZSTD_getFrameContentSize +
// ZSTD_decompress calls (signal_compression.cpp:103,117; C5.hpp:494-565,588-655).
#pragma once
#include <cstddef>
#include "zstd_decode.cuh"

namespace rnsc {

constexpr int PLAN_MAX_HUF = 3;       // Huffman-coded literal sections per chunk described in the plan
constexpr int PLAN_MAX_STREAMS = 4 * PLAN_MAX_HUF;
#ifndef RNSC_PLAN_THREADS
#define RNSC_PLAN_THREADS 32
#endif
constexpr int PLAN_THREADS = RNSC_PLAN_THREADS;      // planner CTA size (one chunk per thread)
constexpr uint32_t REGION_HEAD = 16;  // readable bytes in front of the staged chunk (bit-window over-read)
constexpr uint32_t PLAN_SUB_MIN = 192;   // minimum Huffman sub-sequence length in bits
constexpr uint32_t PLAN_SLOT_MAX = 63;   // symbol slot words per Huffman thread at most (252 symbols: more than a sub-sequence holds)
constexpr uint32_t FAST_ORDERED_MAX = 64;   // chunks with at most this many sequences run them in order, one warp per frame

constexpr uint32_t KEY_LUT_BYTES = 256 * 16;   // one copy of the splitter inverse's key-byte table (svb_unpack2.cuh)
__host__ __device__ constexpr uint32_t key_lut_copies(uint32_t nth) { return nth >= 1024 ? 4u : (nth >= 512 ? 2u : 1u); }   // copies a chunk of that class gets if it has room

enum : uint32_t { FK_RAW = 0, FK_RLE = 1, FK_COMP = 2 };
// size classes of the fast kernel (threads per CTA x CTAs per SM): S 256x4, A 512x2, B 1024x1; then the general decoder
enum : uint32_t { CLS_S = 0, CLS_A = 1, CLS_B = 2, CLS_SLOW = 3, CLS_DONE = 4 };
__host__ __device__ inline uint32_t class_threads(uint32_t cls) { return cls == CLS_S ? 256u : (cls == CLS_A ? 512u : 1024u); }

struct FramePlan {               // planner-internal description of one frame
    uint32_t kind, fcs;          // block type of the frame's single block; decoded size
    uint32_t data_off;           // RAW: chunk offset of the bytes; RLE: the byte value
    uint32_t dst_off;            // region offset of the decoded output (RLE, COMP unless zero-copy)
    uint32_t ltype, regen;       // literals: 0 raw, 1 RLE, 2 Huffman; count
    uint32_t lit_off;            // raw literals: chunk offset; RLE literals: the byte value
    uint32_t nstreams, huf;      // Huffman: 1 or 4 streams; index of the HufPlan
    uint32_t nseq, seq_base;     // sequences [seq_base, seq_base + nseq) of the sequence pool
    uint32_t nruns;              // number of non-empty literal runs (the tail after the last match included)
    uint32_t hz_off, hz_len;     // chunk offset / size of the Huffman bitstream area (dead once the literals are decoded)
    uint32_t s_beg[4], s_tbits[4];   // Huffman streams: chunk offset of the first byte; payload bits below the end mark
    uint32_t s_runs[4];              // non-empty literal runs that start inside each stream's literals (frames with sequences)
};
// A FramePlan travels through the plan's global scratch between the two frame loops of plan_chunk: only the fields its kind
// of frame uses are stored and read back (raw / RLE frames: 4 words; Huffman literals: + the stream geometry; + the runs per
// stream when LZ sequences follow) -- every word is a 32-sector access for the warp.
__device__ __forceinline__ void frame_store(FramePlan* g, const FramePlan& fp)
{
    g->kind = fp.kind; g->fcs = fp.fcs; g->data_off = fp.data_off; g->dst_off = fp.dst_off;
    if (fp.kind != FK_COMP) return;
    g->ltype = fp.ltype; g->regen = fp.regen; g->lit_off = fp.lit_off; g->nstreams = fp.nstreams; g->huf = fp.huf;
    g->nseq = fp.nseq; g->seq_base = fp.seq_base; g->nruns = fp.nruns;
    if (fp.ltype != 2) return;
    g->hz_off = fp.hz_off; g->hz_len = fp.hz_len;
    for (int j = 0; j < 4; ++j) { g->s_beg[j] = fp.s_beg[j]; g->s_tbits[j] = fp.s_tbits[j]; }
    if (fp.nseq) for (int j = 0; j < 4; ++j) g->s_runs[j] = fp.s_runs[j];
}
__device__ __forceinline__ void frame_load(FramePlan& fp, const FramePlan* g)
{
    fp.kind = g->kind; fp.fcs = g->fcs; fp.data_off = g->data_off; fp.dst_off = g->dst_off;
    fp.ltype = 0; fp.regen = 0; fp.lit_off = 0; fp.nstreams = 0; fp.huf = 0; fp.nseq = 0; fp.seq_base = 0; fp.nruns = 0; fp.hz_off = 0; fp.hz_len = 0;
    for (int j = 0; j < 4; ++j) { fp.s_beg[j] = 0; fp.s_tbits[j] = 0; fp.s_runs[j] = 0; }
    if (fp.kind != FK_COMP) return;
    fp.ltype = g->ltype; fp.regen = g->regen; fp.lit_off = g->lit_off; fp.nstreams = g->nstreams; fp.huf = g->huf;
    fp.nseq = g->nseq; fp.seq_base = g->seq_base; fp.nruns = g->nruns;
    if (fp.ltype != 2) return;
    fp.hz_off = g->hz_off; fp.hz_len = g->hz_len;
    for (int j = 0; j < 4; ++j) { fp.s_beg[j] = g->s_beg[j]; fp.s_tbits[j] = g->s_tbits[j]; }
    if (fp.nseq) for (int j = 0; j < 4; ++j) fp.s_runs[j] = g->s_runs[j];
}
struct HufPlan {                 // recipe of a HUF_readDTableX1-ordered decoding table
    uint32_t log;                // table log (max code length), <= 12
    uint32_t tab_off;            // region offset of the table (2^log entries of 2 bytes)
    uint32_t skip_off;           // region offset of its skip table (2^log bytes), 0 = none (it would push the chunk into a bigger size class)
    uint32_t lmin;               // shortest code length in use
    uint16_t cls_start[14];      // first table entry of weight class w (1..log); cls_start[log+1] = 2^log
    uint16_t cls_sym[14];        // index into sorted[] of the first symbol of class w
    uint8_t sorted[256];         // symbols of weight > 0 ordered by (weight, symbol)
};
static_assert(offsetof(HufPlan, cls_start) % 4 == 0 && offsetof(HufPlan, cls_sym) % 4 == 0 && offsetof(HufPlan, sorted) % 4 == 0 && sizeof(HufPlan) % 4 == 0,
              "the planner stores these arrays as 32-bit words");
// One Huffman bitstream: payload bits, absolute bit address of the end mark in the region, symbols, region offset of
// the frame's output, literal index of its first symbol, its table; and, when LZ sequences follow, where they are in
// ChunkPlan::seq, so that every literal is written straight to its final place between the matches.
struct StreamPlan { uint32_t tbits, abit, count, out, lit0, tab_off, log, sop, regen, pitch, skip_off, sub; };   // sop: index of the frame's SeqOp or ~0; pitch: bytes per table entry; sub: sub-sequence length in bits
struct FillOp { uint32_t dst_off, len, value; };                         // RLE block / RLE literals without sequences
// LZ sequences of a frame: entries [seq_base, seq_base + nseq] of the sequence pool. lit = region offset of raw literals or
// the RLE byte; meta_off = region offset of the run table the kernel copies there: nruns x {first literal index, shift to
// the final position} (non-empty literal runs only); mbytes = matched bytes = fcs - regen.
struct SeqOp { uint32_t dst_off, ltype, lit, regen, nseq, seq_base, nruns, meta_off, mbytes, pad; };
struct ViewPlan { uint32_t off, len; };                                  // stream k of the splitter inverse = region[off .. off+len)
// Sequence pool entry i of a frame: words 0-5 describe sequence i (literal length, match length, offset, literal position,
// output position, matched bytes in front of it); words 6-7 of entry k = the k-th NON-EMPTY literal run (first literal
// index, shift to its final position = matched bytes in front of it; the tail after the last match included).
constexpr int SEQ_WORDS = 8;
constexpr uint32_t ORIGIN_LIT = 0x80000000u;   // origin word of a matched byte: literal flag | output position, or the index of another matched byte

struct PlanHead {                // what the fast kernel copies to shared memory for every chunk
    uint32_t cls, need, phase, stage_bytes;
    uint32_t nhuf, nst, sub, nfill, nsop, nseq;
    int32_t view_status;
    uint32_t early;                // bit k: the sequences of sop[k] (raw / RLE literals) are executed by the CTA's last warp while the others decode Huffman streams
    uint32_t zone_off, zone_len;   // region bytes that are dead after the Huffman stage (the largest bitstream area): match resolver scratch
    // Thread groups of the Huffman stage (whole warps that synchronise on a named barrier of their own): bit s of grp_mask =
    // stream s is the first of a group, whose threads are [st_first[s], st_first[next leading stream]); grp_total = threads
    // in all groups = st_first[nst]. A group is a Huffman-coded frame, or ONE stream of a frame with LZ sequences.
    uint32_t grp_mask, grp_total;
    uint32_t lut_off;                       // region offset of the chunk's replicated key-byte table (C4 / C5, key_lut_copies(threads) copies), 0 = none: the single copy in front of the region
    uint32_t slot_off, slot_w;              // symbol slots of the Huffman stage: region offset, 32-bit words per thread (odd, 0 = none): what is left of the class's region
    uint32_t st_first[PLAN_MAX_STREAMS + 1];
    StreamPlan st[PLAN_MAX_STREAMS];
    FillOp fill[5];
    SeqOp sop[5];
    ViewPlan view[5];
};
struct ChunkPlan {
    PlanHead head;
    HufPlan h[PLAN_MAX_HUF];
    FramePlan f[5];                  // planner scratch
    uint32_t fse[512 + 256 + 512];   // planner scratch: FSE-compressed LL / OF / ML decoding tables of the frame being planned
};

struct PlanParams {
    const uint8_t* comp; const uint64_t* comp_off; const uint32_t* n; uint32_t nchunks; int codec;
    ChunkPlan* plans; int32_t* status;
    uint32_t* lists;             // 4 lists of nchunks chunk indices: class S, A, B, slow
    uint32_t* counts;            // [0..3] list lengths (atomic), zeroed before the launch
    uint32_t region[3];          // bytes of dynamic region available to a class S / A / B CTA
    const uint32_t* deftab;      // predefined FSE decoding tables: LL[64] OF[32] ML[64]
    uint32_t has_origin;         // the fast kernel has global scratch for the match resolver's origin words (else: chunks that need it go to the general decoder)
    uint32_t cost_bit, cost_sym; // Huffman stage cost model: a sub-sequence of b bits holding s symbols costs cost_bit * b + cost_sym * s
    uint32_t cost_run;           // ... + cost_run per literal run that starts in it (the second pass changes its output position at every run)
    uint32_t* seq_pool; uint32_t pool_cap; unsigned long long* pool_cursor;   // 64-bit cursor: a hostile batch can ask for more than 2^32 entries; bump-allocated sequence storage (SEQ_WORDS words per entry)
};

// Per planner thread scratch in shared memory. An odd number of words, so the same field of the 32 threads of a
// warp falls into 32 different banks. The FSE build temporaries (norm, next16) are dead before the first weight is
// written, so they share the weights' bytes: a smaller scratch = more planner CTAs per SM.
// -DRNSC_PLAN_HYBRID: the weights (written once, read once by the sort) and the temporaries of the sequence tables' FSE
// descriptions live in thread-private local memory instead; the shared scratch keeps the small arrays the serial chains
// read (the weights' FSE table and its 13-symbol temporaries, the staged description): 460 instead of 652 bytes per thread.
struct PlanBig { union { uint8_t weights[256]; struct { int16_t norm[64]; uint16_t next16[64]; }; }; };
struct PlanScratch {
#ifndef RNSC_PLAN_HYBRID
    PlanBig big;
#endif
    uint32_t fse_w[64];
    // stage: the bytes of a Huffman tree description (<= 127) as words, word 0 = a zero word below them, two more behind them
    // (plan_read_weights); dead before the weight counts are written
    union { struct { uint32_t rank_start[16]; uint32_t rank_cnt[16]; int huf_log; }; uint32_t stage[35]; };
#ifdef RNSC_PLAN_HYBRID
    int16_t wnorm[16]; uint16_t wnext[16];      // normalised counts / build temporaries of the weights' FSE table (<= 13 symbols)
#endif
};
#ifdef RNSC_PLAN_HYBRID
static_assert(sizeof(PlanScratch) == 460, "PlanScratch layout");      // 115 words: odd
#define RNSC_PLAN_WNORM(sc, bg) ((sc)->wnorm)
#define RNSC_PLAN_WNEXT(sc, bg) ((sc)->wnext)
#else
static_assert(sizeof(PlanScratch) == 652, "PlanScratch layout");      // 163 words: odd
#define RNSC_PLAN_WNORM(sc, bg) ((bg)->norm)
#define RNSC_PLAN_WNEXT(sc, bg) ((bg)->next16)
#endif

__host__ __device__ inline uint32_t align16(uint32_t v) { return (v + 15u) & ~15u; }
__host__ __device__ inline uint32_t huf_tab_bytes(uint32_t log) { return align16(2u << log); }
__host__ __device__ inline uint32_t huf_skip_bytes(uint32_t log) { return align16(1u << log); }
__host__ __device__ inline uint32_t seq_meta_bytes(uint32_t nseq) { return align16((nseq + 1) * 8); }

// The planner's reader of a Huffman tree description (RFC 8878 4.2.1): same results as huf_read_weights (zstd_decode.cuh, kept
// by the general decoder) -- weights[0 .. nw), the symbol count per weight in rank_cnt, the table log in huf_log, bytes
// consumed (0 = error) -- at a third of the instructions and without dependent global loads: the description (<= 127 bytes)
// is copied to the thread's shared scratch first (independent loads), the FSE header and the backward bitstream are read
// from there with two shared loads and a funnel shift per field, and the per-weight counts and the weight total are kept
// while the weights are decoded instead of in a second pass. Single exit; every lane of the warp calls it (size = 0 to skip).
__device__ __forceinline__ uint32_t stage_bits(const uint32_t* stage, uint32_t bit)      // 32 bits from bit `bit` of the staged bytes (word 1 = byte 0)
{
    const uint32_t wi = 1u + (bit >> 5);
    return __funnelshift_r(stage[wi], stage[wi + 1], bit);
}
__device__ inline uint32_t plan_read_weights(PlanScratch* sc, PlanBig* bg, const uint8_t* src, uint32_t size, int* nw_out)
{
    uint8_t* w = bg->weights; int16_t* const wnorm = RNSC_PLAN_WNORM(sc, bg); uint16_t* const wnext = RNSC_PLAN_WNEXT(sc, bg); uint32_t nw = 0;
    bool ok = size >= 1;
    const uint32_t hb = ok ? ldg8(src) : 0;
    const bool direct = hb >= 128;
    uint32_t used = 0;
    if (ok && direct) { nw = hb - 127; used = 1 + (nw + 1) / 2; if (used > size) ok = false; }
    if (ok && !direct) { used = 1 + hb; if (used > size || hb < 1) ok = false; }
    const uint32_t dbytes = ok ? used - 1 : 0;                        // <= 127
    {
        const uint32_t nwords = (dbytes + 3) >> 2;
        for (uint32_t i = 0; i < nwords; ++i) {
            uint32_t v = gbits32(src + 1, 32 * i);
            if (4 * i + 4 > dbytes) v &= 0xFFFFFFFFu >> (8 * (4 * i + 4 - dbytes));      // bits past the description read as zero
            sc->stage[1 + i] = v;
        }
        sc->stage[0] = 0; sc->stage[1 + nwords] = 0; sc->stage[2 + nwords] = 0;
    }
    unsigned long long c_lo = 0, c_hi = 0;                            // symbols per weight, 8-bit fields (weights 0-7, 8-15): at most 255 weights are listed
    uint32_t total = 0; bool big = false;
    auto tally = [&](uint32_t wi) {
        if (wi > HUF_LOG_MAX) { big = true; wi = 0; }
        total += (1u << wi) >> 1;
        const unsigned long long one = 1ull << ((wi & 7u) * 8u);
        if (wi & 8u) c_hi += one; else c_lo += one;
    };
    // ---- direct representation: one 4-bit weight per symbol, first weight in the high nibble ----
    {
        const uint32_t lim = (ok && direct) ? nw : 0;
        for (uint32_t i = 0; i < lim; ++i) { const uint32_t by = (sc->stage[1 + (i >> 3)] >> (8 * ((i >> 1) & 3))) & 0xFFu, wi = (i & 1) ? (by & 15u) : (by >> 4); w[i] = (uint8_t)wi; tally(wi); }
    }
    // ---- FSE-compressed weights: normalised counts (FSE_readNCount, <= 13 symbols, accuracy log <= 6) ----
    int tl = 0, ns = 0; uint32_t h = 0;
    if (ok && !direct) {
        const uint32_t tbits = 8 * hb;
        uint32_t bp = 4;
        tl = (int)(stage_bits(sc->stage, 0) & 15u) + 5;
        bool bad = tl > 6;
        int remaining = (1 << tl) + 1, threshold = 1 << tl, nb = tl + 1, sym = 0; bool prev0 = false;
        while (!bad && remaining > 1 && sym <= 12) {
            if (prev0) {
                int n0 = sym;
                for (;;) { const uint32_t r = bp < tbits ? stage_bits(sc->stage, bp) & 3u : 0u; bp += 2; n0 += (int)r; if (r != 3 || n0 > 13) break; }
                if (n0 > 13) { bad = true; break; }
                while (sym < n0) wnorm[sym++] = 0;
                if (sym > 12) break;
            }
            const int mx = (2 * threshold - 1) - remaining;
            const uint32_t v = bp < tbits ? stage_bits(sc->stage, bp) : 0u;
            int count = (int)(v & ((1u << (nb - 1)) - 1u));
            if (count < mx) bp += nb - 1;
            else { count = (int)(v & ((1u << nb) - 1u)); if (count >= threshold) count -= mx; bp += nb; }
            count--;
            remaining -= count < 0 ? -count : count;
            wnorm[sym++] = (int16_t)count;
            prev0 = count == 0;
            while (remaining < threshold) { nb--; threshold >>= 1; }
        }
        h = (bp + 7) >> 3;
        if (bad || remaining != 1 || h > hb) ok = false;
        ns = sym;
    }
    warp_converge<true>();
    if (ok && !direct) fse_build(sc->fse_w, wnorm, ns, tl, wnext);
    warp_converge<true>();
    // ---- two interleaved FSE states over the backward bitstream, bytes [h, hb) of the description ----
    if (ok && !direct) {
        const uint32_t last = (sc->stage[1 + ((hb - 1) >> 2)] >> (8 * ((hb - 1) & 3))) & 0xFFu;
        if (hb == h || last == 0) ok = false;
        else {
            const int start = 8 * (int)h;
            int pos = 8 * (int)hb - (8 - highbit32(last));            // bit address of the end mark = end of the payload
            // n bits ending at pos (n <= 6); bits below `start` read as zero
            auto rd = [&](uint32_t n) -> uint32_t {
                pos -= (int)n;
                const uint32_t m = (1u << n) - 1u;
                if (pos >= start) return stage_bits(sc->stage, (uint32_t)pos) & m;
                const int avail = (int)n - (start - pos);
                return avail <= 0 ? 0u : ((stage_bits(sc->stage, (uint32_t)start) & ((1u << avail) - 1u)) << (start - pos)) & m;
            };
            uint32_t s1 = rd((uint32_t)tl), s2 = rd((uint32_t)tl);
            for (;;) {
                if (nw >= 254) { ok = false; break; }
                uint32_t e = sc->fse_w[s1], wi = e & 0xFFu; w[nw++] = (uint8_t)wi; tally(wi);
                s1 = (e >> 16) + rd((e >> 8) & 0xFFu);
                if (pos < start) { if (nw >= 255) ok = false; else { wi = sc->fse_w[s2] & 0xFFu; w[nw++] = (uint8_t)wi; tally(wi); } break; }
                if (nw >= 254) { ok = false; break; }
                e = sc->fse_w[s2]; wi = e & 0xFFu; w[nw++] = (uint8_t)wi; tally(wi);
                s2 = (e >> 16) + rd((e >> 8) & 0xFFu);
                if (pos < start) { if (nw >= 255) ok = false; else { wi = sc->fse_w[s1] & 0xFFu; w[nw++] = (uint8_t)wi; tally(wi); } break; }
            }
        }
    }
    warp_converge<true>();
    // ---- the last weight is implied: the weights must complete a power of two ----
    uint32_t* cnt = sc->rank_cnt;                                     // (the stage is dead from here on)
    if (ok) {
        for (int k = 0; k < 8; ++k) { cnt[k] = (uint32_t)(c_lo >> (8 * k)) & 0xFFu; cnt[8 + k] = (uint32_t)(c_hi >> (8 * k)) & 0xFFu; }
        if (big || total == 0) ok = false;
    }
    if (ok) {
        const int maxBits = highbit32(total) + 1;
        const uint32_t rest = (1u << maxBits) - total;
        if (maxBits > HUF_LOG_MAX || (rest & (rest - 1))) ok = false;
        else {
            const uint32_t lastw = (uint32_t)highbit32(rest) + 1;
            w[nw++] = (uint8_t)lastw; cnt[lastw]++;
            if (cnt[1] < 2 || (cnt[1] & 1)) ok = false;
            sc->huf_log = maxBits;
            *nw_out = (int)nw;
        }
    }
    return ok ? used : 0;
}

// The planner's reader of a frame's sequence bitstream (backward, MSB first; same values as BwdBits / bwd_read of
// zstd_decode.cuh, which the general decoder keeps). The next unread bits stay left-aligned in a 64-bit register window: a
// read is a funnel shift, and the aligned word below the window is loaded one refill AHEAD, so that nobody waits for it
// (the word-at-a-time reader was 43 % of the planner's instructions on VBZ frames -- ~2000 sequences, six reads each --
// and its loads a quarter of the stall samples). Bits below the stream's first byte read as zero; bitpos < 0 says that
// some were asked for.
struct PlanBwd {
    const uint32_t* base;        // the aligned word that holds the stream's first byte
    uint32_t hi, lo;             // window: `have` unread bits at the top of hi:lo, zeros below them
    int have, bitpos;            // bitpos: unread bits left in the stream
    int wnext; uint32_t nextw;   // index of the next word to append (< 0: below the stream) and the word itself
    uint32_t low_mask;           // bits of word 0 that belong to the stream
};
__device__ __forceinline__ uint32_t planbwd_word(const PlanBwd& b, int w) { return w > 0 ? ldg32(b.base + w) : (w == 0 ? ldg32(b.base) & b.low_mask : 0u); }
__device__ inline int planbwd_init(PlanBwd& b, const uint8_t* p, uint32_t size)
{
    b.base = (const uint32_t*)((uintptr_t)p & ~(uintptr_t)3); b.hi = b.lo = 0; b.have = 0; b.bitpos = 0; b.wnext = -1; b.nextw = 0; b.low_mask = 0;
    if (size == 0) return -1;
    const uint32_t last = ldg8(p + size - 1);
    if (last == 0) return -1;
    const uint32_t sh0 = 8u * (uint32_t)((uintptr_t)p & 3);
    b.low_mask = 0xFFFFFFFFu << sh0;
    b.bitpos = (int)size * 8 - (8 - highbit32(last));
    const uint32_t top = (uint32_t)b.bitpos + sh0;            // bit index of the end mark, counted from word 0
    if (top == 0) return 0;
    const int w0 = (int)((top - 1) >> 5);                     // the word that holds the highest payload bit
    const uint32_t k = top - 32u * (uint32_t)w0;              // payload bits in it: 1 .. 32
    const uint32_t word = planbwd_word(b, w0);
    b.hi = (k < 32 ? word & ((1u << k) - 1u) : word) << (32 - k);
    b.have = (int)k; b.wnext = w0 - 1; b.nextw = planbwd_word(b, b.wnext);
    return 0;
}
__device__ __forceinline__ uint32_t planbwd_read(PlanBwd& b, uint32_t n)      // n <= 31
{
    if (b.have < 32) {                                        // every unread bit is in hi: append the word below
        b.hi |= b.nextw >> b.have;
        b.lo = b.have ? b.nextw << (32 - b.have) : 0u;
        b.have += 32; --b.wnext;
        b.nextw = planbwd_word(b, b.wnext);                   // wanted at the next refill only
    }
    const uint32_t v = __funnelshift_l(b.hi, 0u, n);          // the top n bits of the window (0 for n == 0)
    b.hi = __funnelshift_l(b.lo, b.hi, n); b.lo <<= n;
    b.have -= (int)n; b.bitpos -= (int)n;
    return v;
}

// Decodes `nseq` sequences given the three decoding tables (predefined, RLE = one entry with log 0, or built from
// an FSE description). Returns 0 on success. Single exit (the planner runs one chunk per lane).
__device__ inline int plan_sequences(const uint8_t* bs, uint32_t bs_size, uint32_t nseq, const uint32_t* tll, int ll_log, const uint32_t* tof, int of_log,
                                     const uint32_t* tml, int ml_log, uint32_t regen, uint32_t fcs, uint32_t* out, uint32_t* nruns, uint32_t lit_per_stream, uint32_t* runs4)
{
    unsigned long long rc = 0;                  // runs per Huffman stream of the literals, four 16-bit counters (saturating use: a cost estimate)
    PlanBwd b;
    bool bad = planbwd_init(b, bs, bs_size) != 0;
    uint32_t sl = 0, so = 0, sm = 0;
    // The table entries of sequence i + 1 are fetched as soon as its states are known -- right behind the extra bits of
    // sequence i, whose offset history, checks and stores then run while the three loads are in flight (the tables of an
    // FSE-described frame live in the plan's global scratch: an L2 round trip, 13 % of the planner's stall samples on VBZ).
    uint32_t el = 0, eo = 0, em = 0;
    if (!bad) {
        sl = planbwd_read(b, (uint32_t)ll_log); so = planbwd_read(b, (uint32_t)of_log); sm = planbwd_read(b, (uint32_t)ml_log);
        if (nseq) { el = tll[sl]; eo = tof[so]; em = tml[sm]; }
    }
    uint32_t rep0 = 1, rep1 = 4, rep2 = 8, op = 0, lp = 0, nr = 0;
    for (uint32_t i = 0; i < nseq && !bad; ++i) {
        const uint32_t llc = el & 0xFF, ofc = eo & 0xFF, mlc = em & 0xFF;
        if (llc > 35 || mlc > 52 || ofc > 31) { bad = true; break; }
        const uint32_t ofv = (1u << ofc) + planbwd_read(b, ofc);
        const uint32_t ml = kMLBase[mlc] + planbwd_read(b, (uint32_t)kMLBits[mlc]);
        const uint32_t ll = kLLBase[llc] + planbwd_read(b, (uint32_t)kLLBits[llc]);
        uint32_t nel = 0, neo = 0, nem = 0;
        if (i + 1 < nseq) {
            sl = (el >> 16) + planbwd_read(b, (el >> 8) & 0xFF);
            sm = (em >> 16) + planbwd_read(b, (em >> 8) & 0xFF);
            so = (eo >> 16) + planbwd_read(b, (eo >> 8) & 0xFF);
            nel = tll[sl]; neo = tof[so]; nem = tml[sm];
        }
        uint32_t offset;
        if (ofv > 3) { offset = ofv - 3; rep2 = rep1; rep1 = rep0; rep0 = offset; }
        else {
            const uint32_t idx = ofv - 1 + (ll == 0);
            if (idx == 0) offset = rep0;
            else {
                offset = idx == 1 ? rep1 : (idx == 2 ? rep2 : rep0 - 1);
                if (offset == 0) { bad = true; break; }
                if (idx > 1) rep2 = rep1;
                rep1 = rep0; rep0 = offset;
            }
        }
        if (b.bitpos < 0 || lp + ll > regen || op + ll + ml > fcs || offset > op + ll) { bad = true; break; }
        out[SEQ_WORDS * i] = ll; out[SEQ_WORDS * i + 1] = ml; out[SEQ_WORDS * i + 2] = offset; out[SEQ_WORDS * i + 3] = lp; out[SEQ_WORDS * i + 4] = op; out[SEQ_WORDS * i + 5] = op - lp;
        if (ll) {      // compact list of the non-empty literal runs (nr <= i)
            out[SEQ_WORDS * nr + 6] = lp; out[SEQ_WORDS * nr + 7] = op - lp; ++nr;
            const uint32_t sidx = (lp >= lit_per_stream) + (lp >= 2 * lit_per_stream) + (lp >= 3 * lit_per_stream);
            rc += 1ull << (16 * sidx);
        }
        op += ll + ml; lp += ll;
        el = nel; eo = neo; em = nem;
    }
    if (!bad && (b.bitpos != 0 || op + (regen - lp) != fcs)) bad = true;
    if (!bad) {
        if (regen > lp) { out[SEQ_WORDS * nr + 6] = lp; out[SEQ_WORDS * nr + 7] = op - lp; ++nr; }   // the tail (the allocation has nseq + 1 entries)
        *nruns = nr;
        for (int j = 0; j < 4; ++j) runs4[j] = (uint32_t)(rc >> (16 * j)) & 0xFFFFu;
    }
    return bad ? 1 : 0;
}

// Plans one compressed block (the only block of its frame). Returns 0 = planned, 1 = route to the slow path (or
// `enable` false). Every lane of the warp calls it for every frame (enable = false when the lane has nothing to plan):
// the body is single-exit so that the warp can reconverge in front of each long loop.
__device__ inline int plan_block(bool enable, const uint8_t* chunk, uint32_t boff, uint32_t size, FramePlan& fp, ChunkPlan* pl, uint32_t& nhuf, uint32_t& nseq_tot,
                                 PlanScratch* sc, PlanBig* bg, const uint32_t* deftab, const PlanParams& P)
{
    const uint8_t* src = chunk + boff;
    bool bad = !enable || size < 1;
    const uint32_t b0 = bad ? 0 : ldg8(src);
    const uint32_t ltype = b0 & 3, sf = (b0 >> 2) & 3;
    uint32_t lhs = 0, regen = 0, comp = 0, nstreams = 1;
    if (ltype == 3) bad = true;                     // treeless literals need a previous block's table
    if (!bad) {
        if (ltype < 2) {
            if (sf == 0 || sf == 2) { lhs = 1; regen = b0 >> 3; }
            else if (sf == 1) { if (size < 2) bad = true; else { lhs = 2; regen = (b0 >> 4) | (ldg8(src + 1) << 4); } }
            else { if (size < 3) bad = true; else { lhs = 3; regen = (b0 >> 4) | (ldg8(src + 1) << 4) | (ldg8(src + 2) << 12); } }
        } else if (size < 3) bad = true;
        else {
            uint64_t h = 0; for (uint32_t i = 0; i < 5 && i < size; ++i) h |= (uint64_t)ldg8(src + i) << (8 * i);
            if (sf == 0 || sf == 1) { lhs = 3; regen = (uint32_t)(h >> 4) & 0x3FF; comp = (uint32_t)(h >> 14) & 0x3FF; nstreams = sf ? 4 : 1; }
            else if (sf == 2) { lhs = 4; regen = (uint32_t)(h >> 4) & 0x3FFF; comp = (uint32_t)(h >> 18) & 0x3FFF; nstreams = 4; }
            else { if (size < 5) bad = true; lhs = 5; regen = (uint32_t)(h >> 4) & 0x3FFFF; comp = (uint32_t)(h >> 22) & 0x3FFFF; nstreams = 4; }
        }
    }
    if (regen > ZSTD_BLOCK_MAX) bad = true;
    const uint32_t lit_section = lhs + (ltype == 0 ? regen : (ltype == 1 ? 1 : comp));
    if (!bad && lit_section >= size) bad = true;    // the sequences header needs at least one byte
    // ---- sequences header ----
    uint32_t nseq = 0, modes = 0, so = lit_section;
    if (!bad) {
        const uint32_t n0 = ldg8(src + so);
        if (n0 == 0) { so += 1; if (so != size) bad = true; }
        else {
            if (n0 < 128) { nseq = n0; so += 1; }
            else if (n0 < 255) { if (so + 2 > size) bad = true; else { nseq = ((n0 - 128) << 8) + ldg8(src + so + 1); so += 2; } }
            else { if (so + 3 > size) bad = true; else { nseq = ldg8(src + so + 1) + (ldg8(src + so + 2) << 8) + 0x7F00; so += 3; } }
            if (!bad && so >= size) bad = true;
            if (!bad) { modes = ldg8(src + so); so += 1; if (modes & 3) bad = true; }
        }
    }
    if (!bad && nseq == 0 && regen != fp.fcs) bad = true;
    if (!bad) {
        fp.ltype = ltype; fp.regen = regen; fp.nstreams = nstreams; fp.nseq = nseq; fp.seq_base = 0; fp.nruns = 0;
        fp.lit_off = ltype == 0 ? boff + lhs : (ltype == 1 ? ldg8(src + lhs) : 0);
    }
    // ---- Huffman tree description + stream geometry (every lane walks through plan_read_weights) ----
    bool huf = !bad && ltype == 2;
    if (huf && nhuf >= PLAN_MAX_HUF) { bad = true; huf = false; }
    int nw = 0;
    const uint32_t u = plan_read_weights(sc, bg, src + lhs, huf ? comp : 0u, &nw);
    if (huf && !u) { bad = true; huf = false; }
    if (huf) {
        uint32_t p = lhs + u, left = comp - u;
        fp.hz_off = boff + p; fp.hz_len = left;
        uint32_t sz[4] = {0, 0, 0, 0};
        if (nstreams == 1) { sz[0] = left; fp.s_beg[0] = boff + p; }
        else if (left < 6) bad = true;
        else {
            sz[0] = ldg8(src + p) | (ldg8(src + p + 1) << 8); sz[1] = ldg8(src + p + 2) | (ldg8(src + p + 3) << 8); sz[2] = ldg8(src + p + 4) | (ldg8(src + p + 5) << 8);
            if (6 + sz[0] + sz[1] + sz[2] > left || 3 * ((regen + 3) / 4) > regen) bad = true;
            else {
                sz[3] = left - 6 - sz[0] - sz[1] - sz[2];
                fp.s_beg[0] = boff + p + 6; fp.s_beg[1] = fp.s_beg[0] + sz[0]; fp.s_beg[2] = fp.s_beg[1] + sz[1]; fp.s_beg[3] = fp.s_beg[2] + sz[2];
            }
        }
        for (uint32_t j = 0; j < nstreams && !bad; ++j) {
            if (sz[j] == 0) { bad = true; break; }
            const uint32_t last = ldg8(chunk + fp.s_beg[j] + sz[j] - 1);
            if (last == 0) { bad = true; break; }
            fp.s_tbits[j] = 8 * (sz[j] - 1) + (uint32_t)highbit32(last);
        }
        if (bad) huf = false;
    }
    warp_converge<true>();
    // table recipe: weight classes in ascending weight (= descending code length), symbols in natural order
    HufPlan* hp = &pl->h[huf ? nhuf : 0];
    unsigned long long p_lo = 0, p_hi = 0;           // running insert position of every class, 8-bit fields (weights 0-7, 8-15)
    uint32_t nsym = 0;                               // symbols of weight > 0
    if (huf) {
        const int maxBits = sc->huf_log;
        uint32_t e = 0, s = 0, lmin = 0;             // plan_read_weights left the per-weight symbol counts in rank_cnt
        uint32_t cs[14], cy[14];                     // class starts / first sorted index: built in registers, stored as 7 + 7 words
        cs[0] = 0; cy[0] = 0;
#pragma unroll
        for (int k = 1; k < 14; ++k) {
            cs[k] = e; cy[k] = s;                    // classes above maxBits: e == 2^maxBits, s == all symbols
            if (k <= maxBits) {
                const uint32_t cnt = sc->rank_cnt[k];
                if (cnt) lmin = (uint32_t)(maxBits + 1 - k);
                if (k & 8) p_hi |= (unsigned long long)(s & 0xFFu) << ((k & 7) * 8); else p_lo |= (unsigned long long)(s & 0xFFu) << ((k & 7) * 8);
                e += cnt << (k - 1); s += cnt;
            }
        }
        uint32_t* const w_start = (uint32_t*)hp->cls_start; uint32_t* const w_sym = (uint32_t*)hp->cls_sym;
#pragma unroll
        for (int i = 0; i < 7; ++i) { w_start[i] = cs[2 * i] | (cs[2 * i + 1] << 16); w_sym[i] = cy[2 * i] | (cy[2 * i + 1] << 16); }
        hp->lmin = lmin; hp->log = (uint32_t)maxBits;
        nsym = s;
    }
    warp_converge<true>();
    // The symbols are scattered into the (dead by now) FSE table of the weights in shared memory and leave as whole words:
    // a byte store per symbol straight into the plan was 512 of the ~800 global store instructions of a chunk, each one 32
    // sectors wide across the warp (the plans of the 32 chunks of a warp lie 7.6 KB apart).
    uint8_t* const srt = (uint8_t*)sc->fse_w;
    {
        const int lim = huf ? nw : 0;
        // a position is at most 255 when it is used; the increment past the last symbol of the heaviest class may carry
        // into the (unused) field above it
        // weights are read a word at a time (PlanBig is 4-byte aligned, weights[] sits at its start and holds 256 bytes)
        for (int i0 = 0; i0 < lim; i0 += 4) {
            const uint32_t w4 = *reinterpret_cast<const uint32_t*>(bg->weights + i0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t w = (w4 >> (8 * j)) & 0xFFu;
                if (w && i0 + j < lim) {
                    const uint32_t sh = (w & 7u) * 8u;
                    const uint32_t pos = (uint32_t)(((w & 8u) ? p_hi : p_lo) >> sh) & 0xFFu;
                    srt[pos] = (uint8_t)(i0 + j);
                    if (w & 8u) p_hi += 1ull << sh; else p_lo += 1ull << sh;
                }
            }
        }
    }
    if (huf) {
        uint32_t* const dst = (uint32_t*)hp->sorted;
        for (uint32_t i = 0; 4 * i < nsym; ++i) dst[i] = sc->fse_w[i];
    }
    if (huf) fp.huf = nhuf++;
    warp_converge<true>();
    // ---- sequences: predefined, RLE or FSE-described tables (repeat mode is invalid in a first block) ----
    if (!bad && nseq) {
        const unsigned long long base64 = atomicAdd(P.pool_cursor, (unsigned long long)nseq + 1);   // bump allocation in the sequence pool (+1: the tail run)
        if (base64 + nseq + 1 > P.pool_cap) bad = true;
        const uint32_t base = bad ? 0u : (uint32_t)base64;
        fp.seq_base = base;
        uint32_t rle[3] = {0, 0, 0};
        const uint32_t* tab[3] = {deftab, deftab, deftab}; int tlog[3] = {0, 0, 0};
        uint32_t p = so;
        for (int t = 0; t < 3 && !bad; ++t) {       // order in the stream: LL, OF, ML
            const uint32_t mode = (modes >> (6 - 2 * t)) & 3;
            const uint32_t max_sym = t == 0 ? 35u : (t == 1 ? 31u : 52u);
            if (mode == 0) { tab[t] = deftab + (t == 0 ? 0 : (t == 1 ? 64 : 96)); tlog[t] = t == 1 ? 5 : 6; }
            else if (mode == 1) {
                if (p >= size) { bad = true; break; }
                const uint32_t sym = ldg8(src + p); p += 1;
                if (sym > max_sym) { bad = true; break; }
                rle[t] = sym; tab[t] = &rle[t]; tlog[t] = 0;
            } else if (mode == 2) {
                int tl = 0, ns = 0;
                const uint32_t h = fse_read_ncount(src + p, size - p, bg->norm, (int)max_sym, t == 1 ? 8 : 9, &tl, &ns);
                if (!h) { bad = true; break; }
                uint32_t* tt = pl->fse + (t == 0 ? 0 : (t == 1 ? 512 : 768));
                fse_build(tt, bg->norm, ns, tl, bg->next16);
                tab[t] = tt; tlog[t] = tl; p += h;
            } else { bad = true; break; }
        }
        if (!bad && p >= size) bad = true;
        if (!bad && plan_sequences(src + p, size - p, nseq, tab[0], tlog[0], tab[1], tlog[1], tab[2], tlog[2], regen, fp.fcs, P.seq_pool + (size_t)SEQ_WORDS * base, &fp.nruns,
                                    nstreams == 4 ? (regen + 3) / 4 : (regen ? regen : 1u), fp.s_runs)) bad = true;
        if (!bad) nseq_tot += nseq;
    }
    return bad ? 1 : 0;
}

__device__ inline void plan_emit(const PlanParams& P, uint32_t c, uint32_t cls)
{
    P.plans[c].head.cls = cls;
    if (cls <= CLS_SLOW) { const uint32_t i = atomicAdd(P.counts + cls, 1u); P.lists[(size_t)cls * P.nchunks + i] = c; }
}

__device__ inline void plan_chunk(const PlanParams& P, uint32_t c, PlanScratch* sc, PlanBig* bg, const uint32_t* deftab)
{
    ChunkPlan* pl = P.plans + c;
    PlanHead* hd = &pl->head;
    const uint8_t* src = P.comp + P.comp_off[c];
    const uint32_t clen = (uint32_t)(P.comp_off[c + 1] - P.comp_off[c]);
    const uint32_t n = P.n[c];
    const int nf = codec_frames(P.codec);
    const uint64_t bound = stream_bound(P.codec, n);
    const uint32_t phase = (uint32_t)((uintptr_t)src & 15u);
    const uint32_t stage = align16(phase + clen);
    const uint32_t cbase = REGION_HEAD + phase;                    // region offset of chunk byte 0
    uint32_t region = REGION_HEAD + stage + 16, pos = 0, nhuf = 0, nseq_tot = 0;
    uint32_t meta_bytes = 0;                                       // run tables of the frames with sequences
    uint32_t hlog[PLAN_MAX_HUF] = {0, 0, 0};                       // table logs: kept here, not read back from the plan in global memory
    int st = RNSC_OK; bool slow = false;
    // Uniform frame loop: every lane runs all nf iterations (a lane whose chunk already failed just idles through
    // them), so that the warp can reconverge at the marked points.
    for (int f = 0; f < nf; ++f) {
        warp_converge<true>();
        uint32_t fl = 0, hdr = 0, known = 0, ck = 0; uint64_t fcs = 0;
        if (st == RNSC_OK) {
            if (f + 1 < nf) {
                if (pos + 8 > clen) st = RNSC_ERR_CONTAINER;
                else {
                    unsigned long long v = 0; for (int i = 0; i < 8; ++i) v |= (unsigned long long)ldg8(src + pos + i) << (8 * i);
                    pos += 8;
                    if (v > clen - pos) st = RNSC_ERR_CONTAINER; else fl = (uint32_t)v;
                }
            } else fl = clen - pos;
        }
        if (st == RNSC_OK) {
            const int r = zstd_parse_frame_header(src + pos, fl, &hdr, &fcs, &known, &ck);
            if (r || !known) st = RNSC_ERR_NOT_ZSTD;                  // ZSTD_getFrameContentSize must succeed
            else if (fcs > bound) st = RNSC_ERR_REMAINING;
        }
        const bool live = st == RNSC_OK && !slow;
        FramePlan fp;
        fp.kind = FK_RAW; fp.fcs = (uint32_t)fcs; fp.data_off = 0; fp.dst_off = 0; fp.ltype = 0; fp.regen = 0; fp.lit_off = 0;
        fp.nstreams = 0; fp.huf = 0; fp.nseq = 0; fp.seq_base = 0; fp.nruns = 0; fp.hz_off = 0; fp.hz_len = 0;
        for (int j = 0; j < 4; ++j) { fp.s_beg[j] = 0; fp.s_tbits[j] = 0; fp.s_runs[j] = 0; }
        bool comp_block = false; uint32_t bsz = 0;
        if (live) {
            if (hdr + 3 > fl) slow = true;
            else {
                const uint32_t bh = ldg8(src + pos + hdr) | (ldg8(src + pos + hdr + 1) << 8) | (ldg8(src + pos + hdr + 2) << 16);
                const uint32_t last = bh & 1, type = (bh >> 1) & 3;
                bsz = bh >> 3;
                const uint32_t body = type == 1 ? 1u : bsz;
                if (!last || type == 3 || hdr + 3 + body + (ck ? 4 : 0) != fl) slow = true;
                else if (type == 0) { if (bsz != (uint32_t)fcs) slow = true; fp.kind = FK_RAW; fp.data_off = pos + hdr + 3; }
                else if (type == 1) { if (bsz != (uint32_t)fcs) slow = true; fp.kind = FK_RLE; fp.data_off = ldg8(src + pos + hdr + 3); }
                else { fp.kind = FK_COMP; fp.data_off = pos + hdr + 3; comp_block = true; }
            }
        }
        comp_block = comp_block && !slow;
        const int pb = plan_block(comp_block, src, pos + hdr + 3, bsz, fp, pl, nhuf, nseq_tot, sc, bg, deftab, P);
        if (comp_block && pb) slow = true;
        if (st == RNSC_OK && !slow) {
            // placement of the decoded output in the CTA's dynamic region (raw data is used in place; the literals of a
            // frame with sequences are written straight to their final positions)
            const bool zero_copy = fp.kind == FK_RAW || (fp.kind == FK_COMP && fp.ltype == 0 && fp.nseq == 0);
            if (!zero_copy) { fp.dst_off = region; region += align16(fp.fcs) + 16; }
            if (fp.kind == FK_COMP && fp.nseq) meta_bytes += seq_meta_bytes(fp.nseq);
            if (fp.kind == FK_COMP && fp.ltype == 2) {
#pragma unroll
                for (int h = 0; h < PLAN_MAX_HUF; ++h) if ((uint32_t)h == fp.huf) hlog[h] = (uint32_t)sc->huf_log;      // plan_read_weights left it there
            }
            frame_store(&pl->f[f], fp);
        }
        if (st == RNSC_OK) pos += fl;
    }
    if (st != RNSC_OK) { P.status[c] = st; plan_emit(P, c, CLS_DONE); return; }
    if (slow) { plan_emit(P, c, CLS_SLOW); return; }
    // ---- Huffman tables live in the region too ----
    uint32_t htab[PLAN_MAX_HUF] = {0, 0, 0}, hskip[PLAN_MAX_HUF] = {0, 0, 0};
#pragma unroll
    for (int h = 0; h < PLAN_MAX_HUF; ++h) if ((uint32_t)h < nhuf) { htab[h] = region; pl->h[h].tab_off = region; region += huf_tab_bytes(hlog[h]); }
    uint32_t meta = region;                                      // run tables of the frames with sequences
    region += meta_bytes;
    const uint32_t cls = region <= P.region[0] ? CLS_S : (region <= P.region[1] ? CLS_A : (region <= P.region[2] ? CLS_B : CLS_SLOW));
    if (cls == CLS_SLOW) { plan_emit(P, c, CLS_SLOW); return; }
    {   // skip tables (pass 1 strides over several symbols per access) go behind everything else, if the class has room for them
        uint32_t r2 = region;
#pragma unroll
        for (int h = 0; h < PLAN_MAX_HUF; ++h) if ((uint32_t)h < nhuf) r2 += huf_skip_bytes(hlog[h]);
        const bool room = r2 <= P.region[cls];
#pragma unroll
        for (int h = 0; h < PLAN_MAX_HUF; ++h) if ((uint32_t)h < nhuf) { if (room) { hskip[h] = region; region += huf_skip_bytes(hlog[h]); } pl->h[h].skip_off = hskip[h]; }
    }
    const uint32_t nth = class_threads(cls);
    hd->lut_off = 0;
    if ((P.codec == RNSC_C4 || P.codec == RNSC_C5) && key_lut_copies(nth) > 1) {      // replicated key-byte table of the splitter inverse, if the class has room for it
        const uint32_t bytes = KEY_LUT_BYTES * key_lut_copies(nth);
        if (region + bytes <= P.region[cls]) { hd->lut_off = region; region += bytes; }
    }
    // ---- the kernel's program: Huffman streams, fills, sequence ops, views ----
    uint32_t nst = 0, ttot = 0, nfill = 0, nsop = 0, zone_off = 0, zone_len = 0;
    uint32_t sruns[PLAN_MAX_STREAMS];                  // literal runs per Huffman stream (0 without sequences)
    uint32_t stb[PLAN_MAX_STREAMS], stc[PLAN_MAX_STREAMS];   // payload bits / symbols per stream (thread-private copies: hd->st lives in global memory)
    for (int s = 0; s < PLAN_MAX_STREAMS; ++s) { sruns[s] = 0; stb[s] = 0; stc[s] = 0; }
    // thread groups of the Huffman stage = the Huffman-coded frames in order (fp.huf counts them): bits, symbols, streams, first stream, runs
    uint32_t gt[PLAN_MAX_HUF] = {0, 0, 0}, gc[PLAN_MAX_HUF] = {0, 0, 0}, gn[PLAN_MAX_HUF] = {0, 0, 0}, gs0[PLAN_MAX_HUF] = {0, 0, 0}, gr[PLAN_MAX_HUF] = {0, 0, 0};
    uint32_t ngrp = 0;
    int vst = RNSC_OK;
    const uint32_t kl = key_len(P.codec, n);
    uint32_t tot_fcs = 0;
    for (int f = 0; f < nf; ++f) {
        FramePlan fp; frame_load(fp, &pl->f[f]);
        uint32_t data;                                         // region offset of the frame's decoded bytes
        if (fp.kind == FK_RAW) data = cbase + fp.data_off;
        else if (fp.kind == FK_COMP && fp.ltype == 0 && fp.nseq == 0) data = cbase + fp.lit_off;
        else data = fp.dst_off;
        if (fp.kind == FK_RLE) hd->fill[nfill++] = FillOp{fp.dst_off, fp.fcs, fp.data_off};
        else if (fp.kind == FK_COMP && fp.nseq == 0 && fp.ltype == 1) hd->fill[nfill++] = FillOp{fp.dst_off, fp.fcs, fp.lit_off};
        const uint32_t my_sop = (fp.kind == FK_COMP && fp.nseq) ? nsop : 0xFFFFFFFFu;
        if (fp.kind == FK_COMP && fp.ltype == 2) {
            if (fp.hz_len > zone_len + 16) { const uint32_t z0 = align16(cbase + fp.hz_off); zone_off = z0; zone_len = cbase + fp.hz_off + fp.hz_len - z0; }
            const uint32_t q = (fp.regen + 3) / 4;
            uint32_t f_tab = 0, f_skip = 0, f_log = 0;
#pragma unroll
            for (int h = 0; h < PLAN_MAX_HUF; ++h) if ((uint32_t)h == fp.huf) { f_tab = htab[h]; f_skip = hskip[h]; f_log = hlog[h]; }
            const uint32_t g = ngrp++;                            // == fp.huf
            gs0[g] = nst;
            for (uint32_t j = 0; j < fp.nstreams; ++j) {
                StreamPlan sp;
                sp.tbits = fp.s_tbits[j]; sp.abit = 8 * (cbase + fp.s_beg[j]) + fp.s_tbits[j];
                sp.count = fp.nstreams == 1 ? fp.regen : (j < 3 ? q : fp.regen - 3 * q);
                sp.out = fp.dst_off; sp.lit0 = fp.nstreams == 1 ? 0 : j * q;
                sp.tab_off = f_tab; sp.skip_off = f_skip; sp.log = f_log;
                sp.sop = my_sop; sp.regen = fp.regen; sp.pitch = 2; sp.sub = 0;
                if (sp.tbits == 0 && sp.count != 0) slow = true;   // nothing to decode from: the general decoder reports it
                sruns[nst] = fp.nseq ? fp.s_runs[j] : 0u;
                stb[nst] = sp.tbits; stc[nst] = sp.count;
                gt[g] += sp.tbits; gc[g] += sp.count; gn[g] += 1; gr[g] += sruns[nst];
                hd->st[nst++] = sp; ttot += sp.tbits;
            }
        }
        if (fp.kind == FK_COMP && fp.nseq) { hd->sop[nsop++] = SeqOp{fp.dst_off, fp.ltype, fp.ltype == 0 ? cbase + fp.lit_off : fp.lit_off, fp.regen, fp.nseq, fp.seq_base, fp.nruns, meta, fp.fcs - fp.regen, 0}; meta += seq_meta_bytes(fp.nseq); }
        // A.8: stream views = slices of the decompressed buffer(s)
        if (nf == 1) {
            if (fp.fcs < kl) vst = RNSC_ERR_REMAINING;
            else { hd->view[0] = ViewPlan{data, kl}; hd->view[1] = ViewPlan{data + kl, fp.fcs - kl}; }
            for (int k = 2; k < 5; ++k) hd->view[k] = ViewPlan{cbase, 0};
        } else {
            if (f == 0 && fp.fcs != kl) vst = RNSC_ERR_REMAINING;
            hd->view[f] = ViewPlan{data, fp.fcs};
        }
        tot_fcs += fp.fcs;
    }
    for (int f = nf; f < 5 && nf > 1; ++f) hd->view[f] = ViewPlan{cbase, 0};
    if (n == 0 && tot_fcs) vst = RNSC_ERR_REMAINING;
    if (!P.has_origin && nseq_tot > FAST_ORDERED_MAX) {     // the pointer-jumping resolver keeps its origin words in the dead bitstream area or in global scratch
        uint32_t mb = 0;
        for (uint32_t k = 0; k < nsop; ++k) mb += hd->sop[k].mbytes;
        if ((unsigned long long)mb * 4 > zone_len) slow = true;
    }
    if (slow) { plan_emit(P, c, CLS_SLOW); return; }
    // Frames whose sequences do not wait for Huffman literals (C5: Lhigh) get the last warp to themselves during the Huffman stage
    uint32_t early = 0;
    if (nst && nseq_tot <= FAST_ORDERED_MAX) for (uint32_t k = 0; k < nsop; ++k) if (hd->sop[k].ltype != 2) early |= 1u << k;
    const uint32_t nth_h = (early || cls == CLS_B) ? nth - 32 : nth;      // class B: the last warp also fetches the CTA's next chunk index meanwhile
    // Thread groups: the streams of one Huffman-coded frame (same table, same statistics) share a group of whole warps, so
    // that the chain check and the scan of a frame synchronise on a named barrier of their own and frames with short codes
    // (many symbols per bit: long second pass) do not hold up the others. Warps are dealt out in proportion to each
    // frame's estimated cost, cost_bit per payload bit (first pass) + cost_sym per symbol (second pass); the warm-up is the
    // same for every thread and drops out.
    uint32_t gw[PLAN_MAX_HUF] = {0, 0, 0};
    {
        const uint32_t nw = nth_h / 32;
        unsigned long long wsum = 0, wg[PLAN_MAX_HUF];
        for (uint32_t g = 0; g < ngrp; ++g) { wg[g] = (unsigned long long)P.cost_bit * gt[g] + (unsigned long long)P.cost_sym * gc[g] + (unsigned long long)P.cost_run * gr[g] + 1; wsum += wg[g]; }
        uint32_t used = 0, big = 0;
        for (uint32_t g = 0; g < ngrp; ++g) { gw[g] = (uint32_t)((wg[g] * nw + wsum / 2) / wsum); if (gw[g] < 1) gw[g] = 1; used += gw[g]; if (wg[g] > wg[big]) big = g; }
        while (used > nw) { uint32_t m = 0; for (uint32_t g = 1; g < ngrp; ++g) if (gw[g] > gw[m]) m = g; --gw[m]; --used; }     // nw >= 7 > ngrp: never below one warp
        if (ngrp) gw[big] += nw - used;
    }
    uint32_t first = 0, sub_max = PLAN_SUB_MIN, grp_mask = 0;
    for (uint32_t g = 0; g < ngrp; ++g) {
        const uint32_t nt = gw[g] * 32;
        uint32_t sub = nt > gn[g] ? (gt[g] + (nt - gn[g]) - 1) / (nt - gn[g]) : gt[g];
        if (sub < PLAN_SUB_MIN) sub = PLAN_SUB_MIN;
        if (cls == CLS_B) sub = ((sub + 31u) >> 5 | 1u) << 5;   // an odd number of 32-bit words: the lanes of a warp start in 32 different banks of the staged
                                                                // bitstream (worth its rounding loss only where one CTA has the SM to itself)
        if (sub > sub_max) sub_max = sub;
        uint32_t t = first;
        grp_mask |= 1u << gs0[g];
        if (gr[g] && gn[g] > 1 && nt > gn[g]) {
            // a frame with LZ sequences: the streams whose literals are cut into many runs (VBZ: the key region, at the start of
            // the first stream) decode slower in the second pass; threads are dealt out to the streams by cost, not by bits.
            // With a warp or more per stream every stream is a group of its own: a run-dense stream has many threads, i.e. a
            // short first pass and a long second one, the others the opposite -- behind a common barrier between the passes each
            // pass lasted as long as its slowest stream (measured on VBZ: the threads of the first stream waited as long again as
            // their first pass took); alone, a stream's threads go on as soon as their own stream is checked and scanned.
            unsigned long long wsm = 0, wst[4] = {0, 0, 0, 0};
            for (uint32_t k = 0; k < gn[g] && k < 4; ++k) { const uint32_t s = gs0[g] + k;
                wst[k] = (unsigned long long)P.cost_bit * stb[s] + (unsigned long long)P.cost_sym * stc[s] + (unsigned long long)P.cost_run * sruns[s] + 1; wsm += wst[k]; }
            const uint32_t nwg = gw[g], ns = gn[g] < 4 ? gn[g] : 4;
            uint32_t ws4[4] = {0, 0, 0, 0};             // whole warps per stream (own groups), if there are enough
            const bool split = nwg >= 2 * ns;
            if (split) {
                uint32_t used = 0, big = 0;
                for (uint32_t k = 0; k < ns; ++k) { ws4[k] = (uint32_t)((wst[k] * nwg + wsm / 2) / wsm); if (ws4[k] < 1) ws4[k] = 1; used += ws4[k]; if (wst[k] > wst[big]) big = k; }
                while (used > nwg) { uint32_t m = 0; for (uint32_t k = 1; k < ns; ++k) if (ws4[k] > ws4[m]) m = k; --ws4[m]; --used; }     // nwg >= 2 ns: never below one warp
                ws4[big] += nwg - used;
            }
            for (uint32_t k = 0; k < ns; ++k) {
                const uint32_t s = gs0[g] + k;
                uint32_t ts = split ? 32u * ws4[k] : (uint32_t)((wst[k] * nt) / wsm); if (ts < 1) ts = 1;
                uint32_t ss = (stb[s] + ts - 1) / ts;
                if (ss < PLAN_SUB_MIN) ss = PLAN_SUB_MIN;
                if (cls == CLS_B) ss = ((ss + 31u) >> 5 | 1u) << 5;
                if (ss > sub_max) sub_max = ss;
                hd->st[s].sub = ss; hd->st_first[s] = t;
                if (split) { grp_mask |= 1u << s; t += ts; }      // the stream's warps, used or not: the next stream starts on a warp of its own
                else t += (stb[s] + ss - 1) / ss;
            }
        } else
            for (uint32_t s = gs0[g]; s < gs0[g] + gn[g]; ++s) { hd->st[s].sub = sub; hd->st_first[s] = t; t += (stb[s] + sub - 1) / sub; }
        first += nt;
    }
    hd->grp_mask = grp_mask; hd->grp_total = first;
    for (uint32_t s = nst; s <= PLAN_MAX_STREAMS; ++s) hd->st_first[s] = first;
    const uint32_t sub = sub_max;
    // Symbol slots: the part of the class's region that this chunk does not need is dealt out to the Huffman threads; the
    // first pass keeps the first symbols of every sub-sequence there (decoded once, copied by the second pass) -- an odd
    // number of words per thread, so that the lanes of a warp write to 32 different banks.
    {
        const uint32_t so = (region + 3u) & ~3u;
        uint32_t w = (first && P.region[cls] > so + 8) ? (P.region[cls] - so - 8) / 4 / first : 0;
        if (w > PLAN_SLOT_MAX) w = PLAN_SLOT_MAX;
        if (!(w & 1u) && w) --w;
        hd->slot_off = so; hd->slot_w = w >= 3 ? w : 0;
    }
    hd->need = region; hd->phase = phase; hd->stage_bytes = stage;
    hd->nhuf = nhuf; hd->nst = nst; hd->sub = sub; hd->nfill = nfill; hd->nsop = nsop; hd->nseq = nseq_tot; hd->view_status = vst; hd->early = early; hd->zone_off = zone_off; hd->zone_len = zone_len;
    plan_emit(P, c, cls);
}

#ifndef RNSC_PLAN_MINCTAS
#define RNSC_PLAN_MINCTAS 1
#endif
__global__ void __launch_bounds__(PLAN_THREADS, RNSC_PLAN_MINCTAS) k_plan_chunks(PlanParams P)
{
#ifdef RNSC_PLAN_LOCAL
    PlanScratch local_scratch;                       // thread-private (local memory, L1): occupancy is no longer bound by shared memory
    PlanScratch* const sc = &local_scratch;
#else
    __shared__ PlanScratch scratch[PLAN_THREADS];
    PlanScratch* const sc = &scratch[threadIdx.x];
#endif
#ifdef RNSC_PLAN_HYBRID
    PlanBig big;
    PlanBig* const bg = &big;
#else
    PlanBig* const bg = &sc->big;
#endif
    __shared__ uint32_t deftab[160];
    for (int i = threadIdx.x; i < 160; i += PLAN_THREADS) deftab[i] = P.deftab[i];
    __syncthreads();
    const uint32_t c = blockIdx.x * PLAN_THREADS + threadIdx.x;
    if (c < P.nchunks) plan_chunk(P, c, sc, bg, deftab);
}

}  // namespace rnsc
