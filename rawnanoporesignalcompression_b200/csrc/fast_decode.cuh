// fast_decode.cuh -- the fast whole-chunk decoder: executes a ChunkPlan (plan.cuh), so nothing serial is left.
//
// One persistent CTA owns one chunk at a time (three size classes: 256 threads x 4 CTAs per SM, 512 x 2, 1024 x 1):
//   1. one elected thread starts a bulk asynchronous copy (TMA, cp.async.bulk + mbarrier) of the whole compressed
//      chunk into shared memory while all threads copy the plan and build the Huffman decoding tables;
//   2. every Huffman-coded literal section of the chunk (typically keys and M of a C5 chunk: 8 backward
//      bitstreams) is cut into sub-sequences, one per thread, decoded at once from shared memory: speculative
//      start WARM bits early (self-synchronisation), chain check + CTA scan of the symbol counts, second pass writes;
//   3. raw frames are used in place, RLE frames filled, the few LZ sequences executed (literal runs by all
//      threads, matches in order by one warp);
//   4. the splitter inverse (svb_unpack2.cuh) turns the streams, all in shared memory, into int16 samples.
#pragma once
#include "plan.cuh"
#include "svb_codec.cuh"
#include "svb_unpack2.cuh"

namespace rnsc {

// (Tried: symbol slots in GLOBAL memory -- an L2-resident scratch row per slot word, the whole sub-sequence decoded once and
// the second pass a copy -- for the chunks that leave little shared memory: the copy back is bound by the SM's outstanding
// L1 misses, ~3 k cycles per round trip of 12 words per thread with every warp asking at once, 6 k cycles per chunk against
// the 8 k of decoding the symbols again; no gain, 20k-read C5 decode 508 -> 501 GB/s.)
// (Tried: using a 64-bit window up in the counting walk -- strides until fewer than 12 of its bits are left, 7-10 per three
// window loads instead of 4: the data-dependent inner loop costs more than the loads it saves, 20k-read C5 decode 511 -> 504.)
// (Tried: the Huffman stage as a __noinline__ function of its own, to take its registers out of the chunk loop's
// allocation -- no spills, but 3 % slower overall on the B200: 483.7 vs 500.6 GB/s on the 20k-read C5 set.)
#ifndef RNSC_NOINLINE
#define RNSC_NOINLINE
#endif

constexpr uint32_t FAST_WARM_BITS = 192;   // speculative warm-up before a sub-sequence
constexpr uint32_t FAST_NONE = 0xFFFFFFFFu;

struct FastParams {
    const uint8_t* comp; const uint64_t* comp_off; const uint32_t* n; const uint64_t* sample_off;
    int16_t* samples; int32_t* status; int codec;
    const ChunkPlan* plans; const uint32_t* list; const uint32_t* count; uint32_t* counter; uint32_t nchunks;
    const uint32_t* seq_pool;    // the planner's sequence pool (SEQ_WORDS words per entry)
    uint32_t* origin; unsigned long long origin_stride;   // per-CTA global scratch (words) of the match resolver
    unsigned long long* span;    // optional (same switch): per CTA {start, end in ns, chunks decoded}
    unsigned long long* prof;    // optional (rnsc_set_profiling): 8 phase cycle counters of this size class, summed over CTAs
    uint32_t prof_warp;          // phases are timed by the first thread of this warp counted from the end (0 = thread 0 keeps the time)
};

__device__ __forceinline__ unsigned long long global_ns()
{
#ifdef RNSC_EMUL
    return 0;
#else
    unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
#endif
}

// Phase timing for tuning: thread 0 of every CTA adds the cycles since the previous mark to prof[phase].
enum { PH_STAGE = 0, PH_HUF1 = 1, PH_HUFSCAN = 2, PH_HUF2 = 3, PH_SEQ = 4, PH_UNPACK = 5, PH_OTHER = 6, PH_CHUNKS = 7 };
struct PhaseClock {
    unsigned long long* prof; long long t; uint32_t who;      // who: the thread that keeps the time (0; tuning: the first thread of warp RNSC_PROF_WARP from the end)
    __device__ __forceinline__ void start(unsigned long long* p, uint32_t w) { prof = p; who = w; if (prof && threadIdx.x == who) t = clock64(); }
    __device__ __forceinline__ void mark(int phase) { if (prof && threadIdx.x == who) { const long long n = clock64(); atomicAdd(prof + phase, (unsigned long long)(n - t)); t = n; } }
};

template <int NTH>
struct FastWs {
    uint32_t sub_end[NTH], sub_inc[NTH];       // end position of each sub-sequence; in-warp inclusive scan of the symbol counts
    uint32_t warp_tot[NTH / 32];
    uint32_t skip_on;                          // this chunk's Huffman tables have skip tables
    uint32_t stash_tag;                        // chunk whose plan words wait in the stash (sub_end/sub_inc, idle between two Huffman stages)
    uint32_t stash_bad;                        // a thread fetched its share of the stash for another chunk than stash_tag (see fast_huffman): not usable
    Unpack2Ws<NTH> u2;
    PlanHead head;
    HufPlan h[PLAN_MAX_HUF];
    // Long matches of the frames being resolved (a warp each): six arrays of LONG_CAP words. The classes with 512 and 1024
    // threads keep them in the part of the sub-sequence arrays that is idle during the sequence stage (behind the
    // ST_WORDS <= 512 words of the parked plan) -- the largest chunks of the one-CTA-per-SM class need every byte of shared
    // memory; the 256-thread class has no such room and a short queue of its own (its chunks are small).
    static constexpr uint32_t LONG_CAP = NTH >= 1024 ? 128u : (NTH >= 512 ? 64u : 32u);
    uint32_t nlong, long_own[NTH >= 512 ? 1 : 6 * LONG_CAP];
    __device__ __forceinline__ uint32_t* long_arr(int k) { return (NTH >= 512 ? sub_end + 512 : long_own) + LONG_CAP * k; }
    int status; uint32_t next;
    uint32_t n; unsigned long long sample_off;   // the chunk's sample count and offset in the output
    alignas(8) unsigned long long mbar;
};

// ---- bulk asynchronous copy global -> shared (TMA 1-D) ------------------------------------------------
#ifndef RNSC_EMUL
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // earlier generic accesses of dst vs the async proxy
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#else
__device__ inline void mbar_init(unsigned long long*, uint32_t) {}
__device__ inline void bulk_load(void* dst, const void* src, uint32_t bytes, unsigned long long*) { memcpy(dst, src, bytes); }
__device__ inline void mbar_wait(unsigned long long*, uint32_t) {}
#endif

// ---- named barriers: a group of whole warps (the threads of one Huffman-coded frame) synchronises on barrier `id` -------
#ifndef RNSC_EMUL
__device__ __forceinline__ void bar_named(uint32_t id, uint32_t nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ bool bar_named_or(uint32_t id, uint32_t nthreads, bool pred)
{
    uint32_t r;
    asm volatile("{\n.reg .pred q, o;\nsetp.ne.u32 q, %3, 0;\nbar.red.or.pred o, %1, %2, q;\nselp.u32 %0, 1, 0, o;\n}" : "=r"(r) : "r"(id), "r"(nthreads), "r"((uint32_t)pred) : "memory");
    return r != 0;
}
#else
__device__ inline void bar_named(uint32_t id, uint32_t nthreads) { emul::named_barrier((int)id, (int)nthreads); }
__device__ inline bool bar_named_or(uint32_t id, uint32_t nthreads, bool pred)
{
    static int acc[16];
    acc[id] = 0; emul::named_barrier((int)id, (int)nthreads);
    if (pred) acc[id] = 1;
    emul::named_barrier((int)id, (int)nthreads);
    const bool r = acc[id] != 0;
    emul::named_barrier((int)id, (int)nthreads);
    return r;
}
#endif

// ---- Huffman table from its recipe: each thread fills 8 consecutive entries with one 16-byte store --------
// `nthr` threads (t = 0 .. nthr-1) build this table. The weight class of an entry is found by counting class starts
// <= the entry in registers (the starts are 14 halfwords, fetched as 7 broadcast words); a group of 8 entries that
// lies in one class of code length <= L - 3 is one symbol (runs are 2^(w-1) >= 8 entries long and aligned).
// (Tried: groups inside ONE light class -- 1, 2 or 4 entries per symbol -- without the class search per entry: no change.)
__device__ __forceinline__ void fast_build_table(uint16_t* tab, const HufPlan* hp, int t, int nthr)
{
    const uint32_t L = hp->log, size = 1u << L;
    uint32_t cs[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) cs[k] = ((const uint32_t*)hp->cls_start)[k];      // (start[2k], start[2k+1])
    auto cls_of = [&](uint32_t u) -> uint32_t {       // largest w with cls_start[w] <= u (classes above L start at 2^L > u)
        uint32_t w = 0;
#pragma unroll
        for (int k = 0; k < 7; ++k) { w += (cs[k] >> 16) <= u; if (k) w += (cs[k] & 0xFFFFu) <= u; }
        return w;
    };
    for (uint32_t u0 = (uint32_t)t * 8; u0 < size; u0 += (uint32_t)nthr * 8) {
        uint32_t packed[4];
        const uint32_t w0 = cls_of(u0), w7 = cls_of(u0 + 7);
        if (w0 == w7 && w0 >= 4) {
            const uint32_t sym = hp->sorted[hp->cls_sym[w0] + ((u0 - hp->cls_start[w0]) >> (w0 - 1))];
            const uint32_t e = (sym << 8) | (L + 1 - w0);        // low byte: code length, high byte: symbol
            packed[0] = packed[1] = packed[2] = packed[3] = e | (e << 16);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t u = u0 + k;
                uint32_t e = 0;
                if (u < size) {
                    const uint32_t w = cls_of(u);
                    const uint32_t sym = hp->sorted[hp->cls_sym[w] + ((u - hp->cls_start[w]) >> (w - 1))];
                    e = (sym << 8) | (L + 1 - w);
                }
                if (k & 1) packed[k >> 1] |= e << 16; else packed[k >> 1] = e;
            }
        }
        if (size >= 8) *(V16*)(tab + u0) = V16{packed[0], packed[1], packed[2], packed[3]};
        else for (uint32_t k = 0; k < size; ++k) tab[k] = (uint16_t)(packed[k >> 1] >> (16 * (k & 1)));
    }
}

// Skip table of a finished decoding table: entry u = sum of the code lengths | number << 6 of the whole symbols (one
// or two) that the L-bit peek u determines. A second symbol counts only if its code ends inside the peek: then the
// bits behind the peek cannot change it. Built in two steps so that no step has a bank-conflicting access pattern
// of any size: (1) aux: for every code length l in use and every value r of the L - l bits behind such a code, the
// length of the symbol that r starts with, if it fits (0 otherwise) -- 2^(L-l) entries at offset A - 2^(L-l+1),
// A = 2^(L-lmin+1); (2) skip[u] from table[u] and aux, four entries per thread and store.
__device__ __forceinline__ uint32_t skip_aux_bytes(uint32_t L, uint32_t lmin) { return 2u << (L - lmin); }
__device__ __forceinline__ void fast_build_aux(uint8_t* aux, const uint16_t* tab, uint32_t L, uint32_t lmin, int t, int nthr)
{
    const uint32_t A = skip_aux_bytes(L, lmin);
    for (uint32_t idx = (uint32_t)t; idx + 1 < A; idx += (uint32_t)nthr) {
        const uint32_t k = 31u - (uint32_t)__clz((int)(A - idx - 1u));   // bits behind the first code
        const uint32_t l = L - k, r = idx - (A - (2u << k));
        const uint32_t len = tab[r << l] & 0xFFu;
        aux[idx] = (uint8_t)(len <= k ? len : 0u);
    }
}
__device__ __forceinline__ void fast_build_skip(uint8_t* skip, const uint8_t* aux, const uint16_t* tab, uint32_t L, uint32_t lmin, int t, int nthr)
{
    const uint32_t size = 1u << L, A = skip_aux_bytes(L, lmin);
    for (uint32_t u0 = (uint32_t)t * 4; u0 < size; u0 += (uint32_t)nthr * 4) {
        uint32_t w = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t u = (u0 + q) & (size - 1u);
            const uint32_t l1 = tab[u] & 0xFFu, k = L - l1;
            const uint32_t a = aux[A - (2u << k) + (u & ((1u << k) - 1u))];
            w |= ((l1 + a) | (a ? 0x80u : 0x40u)) << (8 * q);
        }
        if (size >= 4) *(uint32_t*)(skip + u0) = w;
        else for (uint32_t q = 0; q < size; ++q) skip[q] = (uint8_t)(w >> (8 * q));
    }
}

// 64 bits of a backward bitstream ending (exclusive) at absolute bit address A of the word array w, left aligned.
__device__ __forceinline__ void bits_ending_at(const uint32_t* w, uint32_t A, uint32_t& hi, uint32_t& lo)
{
    const uint32_t wi = (A - 1) >> 5, sh = (32u - (A & 31u)) & 31u;
    const uint32_t a = w[wi], b = w[wi - 1], c = w[wi - 2];
    hi = __funnelshift_l(b, a, sh); lo = __funnelshift_l(c, b, sh);
}

// Position-based reader of a backward bitstream in the staged chunk. `ap` = bit address of the next unread bit (the
// highest one) + 8 x the shared-memory byte address of the region: (ap >> 3) & ~3 is then the shared address of the word
// that holds the bit and ~ap (mod 32) the left shift that brings it to bit 31 -- a window costs two logic operations,
// three loads and two funnel shifts, and consuming n bits is one subtraction. Reads down to two words below the bit.
struct BitPos {
    uint32_t ap;
#ifdef RNSC_EMUL
    const uint32_t* w;
    __device__ __forceinline__ BitPos(const uint8_t* region, uint32_t A) : ap(A - 1u), w((const uint32_t*)region) {}
    __device__ __forceinline__ void window(uint32_t& hi, uint32_t& lo) const
    {
        const uint32_t wi = ap >> 5, a = w[wi], b = w[wi - 1], c = w[wi - 2];
        hi = __funnelshift_l(b, a, ~ap); lo = __funnelshift_l(c, b, ~ap);
    }
#else
    __device__ __forceinline__ BitPos(const uint8_t* region, uint32_t A) : ap(A - 1u + 8u * (uint32_t)__cvta_generic_to_shared(region)) {}
    __device__ __forceinline__ void window(uint32_t& hi, uint32_t& lo) const
    {
        const uint32_t ad = (ap >> 3) & ~3u;
        uint32_t a, b, c;
        asm volatile("ld.shared.u32 %0, [%3];\nld.shared.u32 %1, [%3+-4];\nld.shared.u32 %2, [%3+-8];" : "=r"(a), "=r"(b), "=r"(c) : "r"(ad));
        hi = __funnelshift_l(b, a, ~ap); lo = __funnelshift_l(c, b, ~ap);
    }
#endif
    __device__ __forceinline__ void consume(uint32_t nbits) { ap -= nbits; }
};

// One symbol: table entry `ev` = symbol << 8 | code length. The funnel shifts take the entry itself as the shift
// amount (they use its low 5 bits; a code is at most 12 bits), so the serial chain is peek -> load -> shift.
// The integer ALU pipe (shifts, logic, 3-input adds) issues at half rate and is what bounds these loops, so the peek
// and the address go through the multiplier pipe instead: index = mulhi(window, 2^L), address = index * pitch + base,
// with the pitch (2 bytes) kept as a run-time value so that it stays a multiply-add.
struct HufTab {
    const uint16_t* tab; const uint8_t* skip; uint32_t L;
    uint32_t addr, saddr, pow, pitch;
    __device__ __forceinline__ HufTab(const uint16_t* t, const uint8_t* sk, uint32_t log, uint32_t pitch_) : tab(t), skip(sk), L(log), pow(1u << log), pitch(pitch_)
    {
#ifdef RNSC_EMUL
        addr = 0; saddr = 0;
#else
        addr = (uint32_t)__cvta_generic_to_shared(t); saddr = (uint32_t)__cvta_generic_to_shared(sk);
#endif
    }
    // skip table: the same peek -> code length sum | symbol count << 6 of the (one or two) whole symbols inside the peek
    __device__ __forceinline__ uint32_t lookup_skip(uint32_t hi) const
    {
#ifdef RNSC_EMUL
        return skip[hi >> (32 - L)];
#else
        uint32_t e;
        asm("{\n.reg .b32 t;\nmul.hi.u32 t, %1, %2;\nadd.u32 t, t, %3;\nld.shared.u8 %0, [t];\n}" : "=r"(e) : "r"(hi), "r"(pow), "r"(saddr));
        return e;
#endif
    }
    __device__ __forceinline__ uint32_t lookup(uint32_t hi) const
    {
#ifdef RNSC_EMUL
        return tab[hi >> (32 - L)];
#else
        uint32_t e;
        asm("{\n.reg .b32 t;\nmul.hi.u32 t, %1, %2;\nmad.lo.u32 t, t, %3, %4;\nld.shared.u16 %0, [t];\n}" : "=r"(e) : "r"(hi), "r"(pow), "r"(pitch), "r"(addr));
        return e;
#endif
    }
};
#define RNSC_HUF_STEP(hi, lo, ht, ev)                                            \
    do { ev = (ht).lookup(hi);                                                   \
         hi = __funnelshift_l(lo, hi, ev); lo = __funnelshift_l(0u, lo, ev); } while (0)

// Advances from consumed-bit position p to the first symbol boundary >= lim; counts the symbols decoded. Nothing is
// written, so while the limit is far the walk strides over up to two symbols per table access (the skip table). The
// last < 48 bits go one symbol at a time out of ONE window (48 + a 12-bit peek fit its 64 bits).
__device__ __forceinline__ void huf_walk(const uint8_t* region, uint32_t abit, const HufTab& tab, uint32_t& p, uint32_t lim, uint32_t& cnt)
{
    uint32_t hi, lo, e;
    BitPos bp(region, abit - p);
    if (tab.skip) {
        while (p + 4 * HUF_LOG_MAX <= lim) {
            bp.window(hi, lo);
            uint32_t sum = 0;                       // bits 0-5: the four length sums (<= 48); above: the symbol counts
#pragma unroll
            for (int k = 0; k < 4; ++k) { e = tab.lookup_skip(hi); hi = __funnelshift_l(lo, hi, e); lo = __funnelshift_l(0u, lo, e); sum += e; }
            p += sum & 63u; bp.consume(sum & 63u); cnt += sum >> 6;
        }
    } else {
        while (p + 4 * HUF_LOG_MAX <= lim) {
            bp.window(hi, lo);
            uint32_t adv = 0;                       // the low byte sums the four code lengths (<= 48)
#pragma unroll
            for (int k = 0; k < 4; ++k) { RNSC_HUF_STEP(hi, lo, tab, e); adv += e; }
            p += adv & 0xFFu; bp.consume(adv & 0xFFu); cnt += 4;
        }
    }
    if (p < lim) {
        bp.window(hi, lo);
        do { RNSC_HUF_STEP(hi, lo, tab, e); p += e & 0xFFu; cnt += 1; } while (p < lim);
    }
}

// `nbytes` symbols kept in a thread's slot (32-bit words, 4 symbols each) go to their final place o (any alignment): byte
// stores up to the next word boundary, then whole words shifted into place, byte stores for what is left. Reads one word
// past the last one that holds a kept symbol.
__device__ __forceinline__ void put_kept(uint8_t* o, const uint32_t* slot, uint32_t nbytes)
{
    const uint32_t h = min(nbytes, (4u - ((uint32_t)(uintptr_t)o & 3u)) & 3u);
    const uint32_t w0 = slot[0];
#pragma unroll
    for (uint32_t q = 0; q < 3; ++q) if (q < h) o[q] = (uint8_t)(w0 >> (8 * q));
    uint32_t* ow = (uint32_t*)(o + h);
    const uint32_t nmid = (nbytes - h) >> 2, sh = 8 * h;
    uint32_t lo = w0;
    for (uint32_t j = 0; j < nmid; ++j) { const uint32_t hi = slot[j + 1]; ow[j] = __funnelshift_r(lo, hi, sh); lo = hi; }
    const uint32_t done = h + 4 * nmid;
    if (done < nbytes) {
        const uint32_t t = __funnelshift_r(lo, slot[nmid + 1], sh);
#pragma unroll
        for (uint32_t q = 0; q < 3; ++q) if (done + q < nbytes) o[done + q] = (uint8_t)(t >> (8 * q));
    }
}

// Plan prefetch: the words the next chunk's start needs (plan head, Huffman table recipes, offset of its compressed
// bytes) are fetched from global memory during this chunk's Huffman stage and parked in the sub-sequence arrays, so
// that a chunk does not begin with a chain of dependent global loads.
constexpr uint32_t ST_HEAD_W = sizeof(PlanHead) / 4, ST_HUF_W = PLAN_MAX_HUF * sizeof(HufPlan) / 4;
constexpr uint32_t ST_COFF = ST_HEAD_W + ST_HUF_W, ST_SOFF = ST_COFF + 2, ST_N = ST_SOFF + 2, ST_WORDS = ST_N + 1;   // + compressed offset, sample offset, sample count
static_assert(ST_WORDS <= 2 * 256, "the stash must fit the smallest CTA's sub-sequence arrays");
static_assert(512 + 6 * 64 <= 2 * 512 && 512 + 6 * 128 <= 2 * 1024, "the long-match queue must fit behind the stash (FastWs::long_arr)");
__device__ __forceinline__ uint32_t stash_fetch(const FastParams& P, uint32_t c, uint32_t idx)
{
    if (idx < ST_HEAD_W) return ((const uint32_t*)&P.plans[c].head)[idx];
    if (idx < ST_HEAD_W + ST_HUF_W) return ((const uint32_t*)P.plans[c].h)[idx - ST_HEAD_W];
    if (idx < ST_SOFF) return ((const uint32_t*)(P.comp_off + c))[idx - ST_COFF];
    if (idx < ST_N) return ((const uint32_t*)(P.sample_off + c))[idx - ST_SOFF];
    return P.n[c];
}

// Decodes every Huffman stream of the chunk (ws->head.st[]) from the staged chunk into the region. The threads of one
// Huffman-coded frame (or of one stream of a frame with LZ sequences) form a group of whole warps (PlanHead::grp_mask) that synchronises on its own named barrier
// between the passes: a frame of short codes (many symbols per bit, long second pass) does not hold up the others, and
// the CTA meets again only at the end. Returns 0 on success (same value in all threads).
template <int NTH>
__device__ RNSC_NOINLINE int fast_huffman(FastWs<NTH>* ws, uint8_t* region, const FastParams& P, PhaseClock& pc)
{
    constexpr int NW = NTH / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* words = (const uint32_t*)region;
    const PlanHead& hd = ws->head;
    const uint32_t nst = hd.nst;
    uint32_t j = 0;
    while (j + 1 < nst && (uint32_t)tid >= hd.st_first[j + 1]) ++j;
    // my group: from the last leading stream at or below mine to the next leading stream (PlanHead::grp_mask); the padding
    // threads behind a group's last sub-sequence were counted into its last stream above
    const uint32_t gmask = hd.grp_mask;
    const uint32_t lead = 31u - (uint32_t)__clz((int)((gmask & ((2u << j) - 1u)) | 1u));
    const uint32_t above = gmask >> (j + 1);
    const uint32_t nlead = above ? j + 1u + (uint32_t)(__ffs((int)above) - 1) : nst;
    const bool member = (uint32_t)tid < hd.grp_total;                     // takes part in its group's barriers
    const uint32_t bar_id = 1u + lead, bar_n = hd.st_first[nlead] - hd.st_first[lead];     // ids 1 .. PLAN_MAX_STREAMS (the hardware has 16)
    const uint32_t t0 = hd.st_first[j], k = (uint32_t)tid - t0;
    const StreamPlan sp = hd.st[j];
    const uint32_t T = sp.tbits, abit = sp.abit, sub = sp.sub;
    const HufTab tab((const uint16_t*)(region + sp.tab_off), ws->skip_on ? region + sp.skip_off : nullptr, sp.log, sp.pitch);
    const uint32_t s0 = k * sub, lim = min(s0 + sub, T);
    const bool active = member && s0 < T;                                  // the padding threads of a group only keep its barriers company
    uint32_t my_start = 0, my_end = 0, my_cnt = 0;
    // Symbol slot of this thread: the first symbols of the sub-sequence are decoded for good in the first pass and kept
    // here (frames without LZ sequences: their literals are one run); the second pass copies them instead of decoding them
    // again. The rest of the sub-sequence is only counted (skip table).
    uint32_t* const slot = (uint32_t*)(region + hd.slot_off) + (uint32_t)tid * hd.slot_w;
    const uint32_t keep_max = sp.sop >= 5 ? 4u * hd.slot_w : 0u;
    uint32_t kept = 0, p_keep = 0;
    if (active) {
        uint32_t p = s0 > FAST_WARM_BITS ? s0 - FAST_WARM_BITS : 0, dummy = 0;
        huf_walk(region, abit, tab, p, s0, dummy);
        my_start = p;
        if (keep_max) {
            BitPos bp(region, abit - p);
            uint32_t hi, lo, e;
            while (kept < keep_max && p + 4 * HUF_LOG_MAX <= lim) {
                bp.window(hi, lo);
                uint32_t word = 0, adv = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    RNSC_HUF_STEP(hi, lo, tab, e); adv += e;
                    word = __byte_perm(word, e, q == 0 ? 0x3215u : (q == 1 ? 0x3250u : (q == 2 ? 0x3510u : 0x5210u)));   // byte q <- symbol
                }
                slot[kept >> 2] = word; kept += 4; p += adv & 0xFFu; bp.consume(adv & 0xFFu);
            }
        }
        p_keep = p;
        huf_walk(region, abit, tab, p, lim, my_cnt);
        my_cnt += kept;
        my_end = p;
    }
    pc.mark(PH_HUF1);
    uint32_t inc = 0;
    bool fail = false;
    if (member) for (int round = 0;; ++round) {
        inc = warp_inclusive_scan(active ? my_cnt : 0u);
        ws->sub_end[tid] = my_end; ws->sub_inc[tid] = inc;
        if (lane == 31) ws->warp_tot[warp] = inc;
        bar_named(bar_id, bar_n);
        // a sub-sequence must start where its predecessor ended; repair the (rare) ones that did not synchronise
        const bool bad = active && k > 0 && ws->sub_end[tid - 1] != my_start;
        const uint32_t want = bad ? ws->sub_end[tid - 1] : 0;
        if (!bar_named_or(bar_id, bar_n, bad)) break;
        if (round >= NTH) { fail = true; break; }    // cannot happen for a consistent stream: every round fixes at least one more sub-sequence
        if (bad) {
            uint32_t p = want; my_start = p; my_cnt = 0; kept = 0; p_keep = p;      // what the slot holds was decoded from a wrong start
            if (p < lim) huf_walk(region, abit, tab, p, lim, my_cnt);
            my_end = p;
        }
    }
    // next chunk's plan words: requested now (ws->next was written before the barriers above), parked after the second pass
    constexpr int PF = (ST_WORDS + NTH - 1) / NTH;
    uint32_t pf[PF];
    const uint32_t nx = ws->next;
#pragma unroll
    for (int q = 0; q < PF; ++q) { const uint32_t idx = (uint32_t)tid + (uint32_t)q * NTH; pf[q] = (nx != FAST_NONE && idx < ST_WORDS) ? stash_fetch(P, nx, idx) : 0u; }
    // output offset inside my stream = exclusive scan of the counts over the warps minus the scan value at the stream's first
    // thread: only warps of my own group lie between the two, the (possibly unfinished) totals of other groups cancel
    const uint32_t wt = lane < NW ? ws->warp_tot[lane] : 0u;
    const uint32_t wex = warp_inclusive_scan(wt) - wt;                        // lane w: symbols before warp w
    const uint32_t ex = __shfl_sync(0xffffffffu, wex, warp) + inc - (active ? my_cnt : 0u);
    const uint32_t base = __shfl_sync(0xffffffffu, wex, (int)(t0 >> 5)) + ((t0 & 31u) ? ws->sub_inc[t0 - 1] : 0u);
    const uint32_t off = ex - base;
    pc.mark(PH_HUFSCAN);
    bool bad = fail;
    if (active && s0 + sub >= T) bad = bad || (my_end != T) || (off + my_cnt != sp.count);      // the stream's last sub-sequence
    if (active && my_cnt && off < sp.count) {
        uint32_t p = my_start, rem = min(my_cnt, sp.count - off), hi, lo, e;      // clamped: a corrupt stream cannot write outside its output
        uint32_t li = sp.lit0 + off;                                                   // literal index of my first symbol
        // Literal runs: without sequences the whole frame is one run at shift 0; with sequences run i (the literals in
        // front of match i; run nseq = the tail) starts at literal index run[i].x and lands run[i].y bytes further
        // (the match bytes in front of it). The run table was built in the region at chunk start.
        const uint32_t nruns = sp.sop < 5 ? hd.sop[sp.sop].nruns : 0u;
        const uint2* run = (const uint2*)(region + (sp.sop < 5 ? hd.sop[sp.sop].meta_off : 0u));
        uint32_t seg = 0, run_end = sp.regen, shift = 0;
        if (nruns) {   // the run that holds literal li: the last one whose first literal index is <= li
            uint32_t lo_i = 0, hi_i = nruns;
            while (lo_i + 1 < hi_i) { const uint32_t mid = (lo_i + hi_i) >> 1; if (run[mid].x <= li) lo_i = mid; else hi_i = mid; }
            seg = lo_i; shift = run[seg].y;
            if (seg + 1 < nruns) run_end = run[seg + 1].x;
        }
        if (li + rem <= run_end) {      // all my literals belong to one run (always, without sequences)
            uint8_t* o = region + sp.out + li + shift;
            if (kept) {                 // symbols kept by the first pass: copy, then go on behind them
                const uint32_t k = min(kept, rem);
                put_kept(o, slot, k);
                o += k; rem -= k; p = p_keep;
            }
            BitPos bp(region, abit - p);
            // up to three symbols bring the output to a word boundary: one window, predicated steps
            const uint32_t head = min(rem, (4u - ((uint32_t)(uintptr_t)o & 3u)) & 3u);
            if (head) {
                bp.window(hi, lo);
                uint32_t adv = 0;
#pragma unroll
                for (uint32_t q = 0; q < 3; ++q) if (q < head) { RNSC_HUF_STEP(hi, lo, tab, e); adv += e; o[q] = (uint8_t)(e >> 8); }
                o += head; rem -= head; bp.consume(adv & 0xFFu);
            }
            for (; rem >= 4; rem -= 4) {
                bp.window(hi, lo);
                uint32_t word = 0, adv = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    RNSC_HUF_STEP(hi, lo, tab, e); adv += e;
                    word = __byte_perm(word, e, q == 0 ? 0x3215u : (q == 1 ? 0x3250u : (q == 2 ? 0x3510u : 0x5210u)));   // byte q <- symbol
                }
                *(uint32_t*)o = word; o += 4; bp.consume(adv & 0xFFu);
            }
            if (rem) {
                bp.window(hi, lo);
#pragma unroll
                for (uint32_t q = 0; q < 3; ++q) if (q < rem) { RNSC_HUF_STEP(hi, lo, tab, e); o[q] = (uint8_t)(e >> 8); }
            }
        } else {
            // runs can be a few literals long (the key region of a VBZ frame): one symbol at a time from a buffered
            // 64-bit window, switching the output position whenever a run ends (the next run is fetched ahead).
            // (Tried: pieces of 16 and more literals through the four-symbols-per-window loop with word stores, the short ones
            // as here -- the lanes of a warp sit in different kinds of piece at any time and the warp pays for both bodies:
            // second pass of a large VBZ chunk 38 k -> 76 k cycles.)
            uint8_t* o = region + sp.out + li + shift;
            uint2 nxt = seg + 1 < nruns ? run[seg + 1] : make_uint2(sp.regen, 0u);
            int avail = 0;
            while (rem) {
                if (avail < (int)HUF_LOG_MAX) { bits_ending_at(words, abit - p, hi, lo); avail = 64; }
                RNSC_HUF_STEP(hi, lo, tab, e);
                p += e & 0xFFu; avail -= (int)(e & 0xFFu);
                *o++ = (uint8_t)(e >> 8); ++li; --rem;
                if (li == run_end) {
                    ++seg; o = region + sp.out + li + nxt.y;
                    nxt = seg + 1 < nruns ? run[seg + 1] : make_uint2(sp.regen, 0u);
                    run_end = nxt.x;
                }
            }
        }
    }
    const int r = __syncthreads_or(bad);
    // The one-CTA-per-SM class writes ws->next from inside the Huffman stage (fast_decode_chunk), ordered before the reads
    // above only for the threads of the writer's own group: a thread of another group may have fetched for the previous value.
    // Behind the CTA-wide barrier the word is stable: a thread that fetched for another chunk parks nothing and marks the stash
    // unusable (the next chunk then reads its plan from global memory: slower, never wrong).
    const uint32_t nx_now = ws->next;
    if (nx != nx_now) ws->stash_bad = 1u;
    else if (nx != FAST_NONE) {
#pragma unroll
        for (int q = 0; q < PF; ++q) { const uint32_t idx = (uint32_t)tid + (uint32_t)q * NTH; if (idx < ST_WORDS) ws->sub_end[idx] = pf[q]; }   // sub_end and sub_inc are contiguous
    }
    if (tid == 0 && nx_now != FAST_NONE) ws->stash_tag = nx_now;
    pc.mark(PH_HUF2);
    return r;
}

// Origin words of bytes [q0, q1) of one match (stored at origin[mb + q]; matched-byte indices are offset by `obase`,
// the frame's first word in the shared origin array). Bytes of the first period (q < offset) read output positions
// ms - of + q: located once in the run table (binary search), then walked -- a position is inside a run's literals
// or in the matched span behind them. Later bytes of an overlapping match repeat the first period: they point at
// their own match (one extra jump round instead of a chain as long as the match).
__device__ inline void emit_origins(const uint2* run, uint32_t nruns, uint32_t regen, uint32_t* origin, uint32_t obase,
                                    uint32_t ms, uint32_t ml, uint32_t of, uint32_t mb, uint32_t q0, uint32_t q1)
{
    const uint32_t qe = min(q1, min(of, ml));
    mb += obase;
    if (q0 < qe) {
        uint32_t x = ms - of + q0;
        int k = -1;                               // last run whose first output position is <= x (-1: none)
        if (nruns && x >= run[0].x + run[0].y) {
            uint32_t lo = 0, hi = nruns;
            while (lo + 1 < hi) { const uint32_t mid = (lo + hi) >> 1; const uint2 r = run[mid]; if (r.x + r.y <= x) lo = mid; else hi = mid; }
            k = (int)lo;
        }
        uint32_t next_lit = k + 1 < (int)nruns ? run[k + 1].x : regen;          // first literal index of the next run
        uint32_t lit_end = k >= 0 ? next_lit + run[k].y : 0u;                    // end of run k's literals in the output
        uint32_t next_start = k + 1 < (int)nruns ? next_lit + run[k + 1].y : 0xFFFFFFFFu;   // first output position of the next run
        for (uint32_t q = q0; q < qe; ++q, ++x) {
            while (x >= next_start) {
                ++k;
                next_lit = k + 1 < (int)nruns ? run[k + 1].x : regen;
                lit_end = next_lit + run[k].y;
                next_start = k + 1 < (int)nruns ? next_lit + run[k + 1].y : 0xFFFFFFFFu;
            }
            origin[mb + q] = x < lit_end ? (ORIGIN_LIT | x) : obase + x - next_lit;
        }
    }
    uint32_t q = max(q0, of);
    for (uint32_t r = q % of; q < q1; ++q) { origin[mb + q] = mb + r; if (++r == of) r = 0; }
}

// Frames with a handful of sequences (C5: keys, Lhigh): ONE warp executes the frame's matches in sequence order, the
// lanes copying the bytes of a match together; different frames of the chunk are handled by different warps at once.
__device__ inline void fast_exec_sequences_ordered(const SeqOp& op, const uint32_t* sq, uint8_t* region)
{
    const uint32_t lane = threadIdx.x & 31u;
    uint8_t* dst = region + op.dst_off;
    const uint32_t nseq = op.nseq;
    if (op.ltype != 2) {
        const uint2* run = (const uint2*)(region + op.meta_off);
        for (uint32_t i = 0; i < op.nruns; ++i) {
            const uint32_t lp = run[i].x, len = (i + 1 < op.nruns ? run[i + 1].x : op.regen) - lp;
            uint8_t* d = dst + lp + run[i].y;
            if (op.ltype == 1) for (uint32_t q = lane; q < len; q += 32) d[q] = (uint8_t)op.lit;
            else { const uint8_t* s = region + op.lit + lp; for (uint32_t q = lane; q < len; q += 32) d[q] = s[q]; }
        }
    }
    uint32_t m_ms[2] = {0, 0}, m_ml[2] = {0, 0}, m_of[2] = {1, 1};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const uint32_t i = lane + k * 32;
        if (i < nseq) { const uint32_t* e = sq + (size_t)SEQ_WORDS * i; m_ms[k] = e[4] + e[0]; m_ml[k] = e[1]; m_of[k] = e[2]; }
    }
    __syncwarp();
    for (uint32_t i = 0; i < nseq; ++i) {
        const int src = (int)(i & 31u);
        const uint32_t ms = __shfl_sync(0xffffffffu, i < 32 ? m_ms[0] : m_ms[1], src), ml = __shfl_sync(0xffffffffu, i < 32 ? m_ml[0] : m_ml[1], src);
        const uint32_t of = __shfl_sync(0xffffffffu, i < 32 ? m_of[0] : m_of[1], src);
        uint8_t* d = dst + ms; const uint8_t* sp = d - of;
        if (of >= ml) { for (uint32_t q = lane; q < ml; q += 32) d[q] = sp[q]; }                      // no overlap
        else if (of == 1) { const uint8_t v = sp[0]; for (uint32_t q = lane; q < ml; q += 32) d[q] = v; }   // a run of one byte
        else if (of >= 32) { for (uint32_t k0 = 0; k0 < ml; k0 += 32) { const uint32_t q = k0 + lane; if (q < ml) d[q] = sp[q]; __syncwarp(); } }
        else { for (uint32_t q = lane; q < ml; q += 32) d[q] = sp[q % of]; }
        __syncwarp();
    }
}

// Executes the LZ sequences of ALL frames of the chunk in one pass.
// Literal runs: Huffman literals were already written in place by fast_huffman; raw / RLE literal runs are copied /
// filled here. Matches: resolved byte-parallel by POINTER JUMPING instead of in sequence order. Every matched byte gets
// an origin word: the output position of a literal, or the index of the matched byte it copies (classified through the
// run table). Chains (a match copying an earlier match) halve every round; then every matched byte is read from its
// literal. One thread per short match, a whole warp per long one. The origin words live in the dead bitstream area
// of the staged chunk when they fit, else in global scratch.
template <int NTH>
__device__ inline void fast_exec_sequences(FastWs<NTH>* ws, const PlanHead& hd, const uint32_t* pool, uint8_t* region, uint32_t* gorigin)
{
    const int tid = threadIdx.x;
    const uint32_t nsop = hd.nsop;
    uint32_t sbase[6], obase[6];                       // first sequence / first origin word of every frame
    sbase[0] = 0; obase[0] = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) { const bool on = (uint32_t)k < nsop; sbase[k + 1] = sbase[k] + (on ? hd.sop[k].nseq : 0u); obase[k + 1] = obase[k] + (on ? hd.sop[k].mbytes : 0u); }
    const uint32_t total_seq = sbase[5], total_mb = obase[5];
    uint32_t* origin = (size_t)total_mb * 4 <= hd.zone_len ? (uint32_t*)(region + hd.zone_off) : gorigin;
    for (uint32_t k = 0; k < nsop; ++k) {
        const SeqOp& op = hd.sop[k];
        if (op.ltype == 2) continue;
        const uint2* run = (const uint2*)(region + op.meta_off);
        uint8_t* dst = region + op.dst_off;
        for (uint32_t i = 0; i < op.nruns; ++i) {
            const uint32_t lp = run[i].x, len = (i + 1 < op.nruns ? run[i + 1].x : op.regen) - lp;
            uint8_t* d = dst + lp + run[i].y;
            if (op.ltype == 1) for (uint32_t q = tid; q < len; q += NTH) d[q] = (uint8_t)op.lit;
            else { const uint8_t* s = region + op.lit + lp; for (uint32_t q = tid; q < len; q += NTH) d[q] = s[q]; }
        }
    }
    constexpr uint32_t LONG = 48, QCAP = FastWs<NTH>::LONG_CAP;
    if (tid == 0) ws->nlong = 0;
    // my (up to two) sequences, fetched once from the global pool: frame, match start, length, offset, matched-byte index
    uint32_t m_k[2] = {0, 0}, m_ms[2] = {0, 0}, m_ml[2] = {0, 0}, m_of[2] = {1, 1}, m_mb[2] = {0, 0};
    auto fetch = [&](uint32_t idx, uint32_t& k, uint32_t& ms, uint32_t& ml, uint32_t& of, uint32_t& mb) {
        k = 0;
#pragma unroll
        for (int j = 1; j < 5; ++j) if (idx >= sbase[j]) k = j;
        const uint32_t* e = pool + (size_t)SEQ_WORDS * (hd.sop[k].seq_base + idx - sbase[k]);
        ms = e[4] + e[0]; ml = e[1]; of = e[2]; mb = e[5];
    };
#pragma unroll
    for (int j = 0; j < 2; ++j) { const uint32_t idx = tid + j * NTH; if (idx < total_seq) fetch(idx, m_k[j], m_ms[j], m_ml[j], m_of[j], m_mb[j]); }
    __syncthreads();
    // ---- round 0: origin of every matched byte ----
    for (uint32_t idx = tid, j = 0; idx < total_seq; idx += NTH, ++j) {
        uint32_t k, ms, ml, of, mb;
        if (j < 2) { k = m_k[j & 1]; ms = m_ms[j & 1]; ml = m_ml[j & 1]; of = m_of[j & 1]; mb = m_mb[j & 1]; } else fetch(idx, k, ms, ml, of, mb);
        if (ml > LONG) {
            const uint32_t q = atomicAdd(&ws->nlong, 1u);
            if (q < QCAP) { ws->long_arr(0)[q] = idx; ws->long_arr(1)[q] = k; ws->long_arr(2)[q] = ms; ws->long_arr(3)[q] = ml; ws->long_arr(4)[q] = of; ws->long_arr(5)[q] = mb; continue; }
        }
        const SeqOp& op = hd.sop[k];
        emit_origins((const uint2*)(region + op.meta_off), op.nruns, op.regen, origin, obase[k], ms, ml, of, mb, 0, ml);
    }
    __syncthreads();
    const uint32_t nl = min(ws->nlong, QCAP), lane = tid & 31u;
    for (uint32_t w = tid >> 5; w < nl; w += NTH / 32) {
        const uint32_t k = ws->long_arr(1)[w], ms = ws->long_arr(2)[w], ml = ws->long_arr(3)[w], of = ws->long_arr(4)[w], mb = ws->long_arr(5)[w];
        const SeqOp& op = hd.sop[k];
        const uint32_t c = (ml + 31) / 32;                                  // a contiguous piece per lane
        emit_origins((const uint2*)(region + op.meta_off), op.nruns, op.regen, origin, obase[k], ms, ml, of, mb, min(ml, lane * c), min(ml, (lane + 1) * c));
    }
    __syncthreads();
    // ---- pointer jumping ----
    for (int round = 0; round < 32; ++round) {
        bool open = false;
        for (uint32_t b = tid; b < total_mb; b += NTH) {
            const uint32_t o = origin[b];
            if (!(o & ORIGIN_LIT)) { const uint32_t o2 = origin[o]; origin[b] = o2; open |= !(o2 & ORIGIN_LIT); }
        }
        if (!__syncthreads_or(open)) break;
    }
    // ---- every matched byte from its literal ----
    for (uint32_t idx = tid, j = 0; idx < total_seq; idx += NTH, ++j) {
        uint32_t k, ms, ml, of, mb;
        if (j < 2) { k = m_k[j & 1]; ms = m_ms[j & 1]; ml = m_ml[j & 1]; mb = m_mb[j & 1]; } else fetch(idx, k, ms, ml, of, mb);
        if (ml > LONG) { bool queued = false; for (uint32_t q = 0; q < nl; ++q) queued |= ws->long_arr(0)[q] == idx; if (queued) continue; }
        uint8_t* dst = region + hd.sop[k].dst_off;
        const uint32_t* o = origin + obase[k] + mb;
        for (uint32_t q = 0; q < ml; ++q) dst[ms + q] = dst[o[q] & ~ORIGIN_LIT];
    }
    for (uint32_t w = tid >> 5; w < nl; w += NTH / 32) {
        const uint32_t k = ws->long_arr(1)[w], ms = ws->long_arr(2)[w], ml = ws->long_arr(3)[w], mb = ws->long_arr(5)[w];
        uint8_t* dst = region + hd.sop[k].dst_off;
        const uint32_t* o = origin + obase[k] + mb;
        for (uint32_t q = lane; q < ml; q += 32) dst[ms + q] = dst[o[q] & ~ORIGIN_LIT];
    }
    __syncthreads();
}

// Decodes chunk c: the stream views are ws->head.view (offsets into the region), the result ws->status. Must be entered by the whole CTA after a barrier.
__device__ __forceinline__ void copy_run_tables(const PlanHead* hd, const uint32_t* pool, uint8_t* region, uint32_t first, uint32_t step)
{
    for (uint32_t k = 0; k < hd->nsop; ++k) {
        const SeqOp op = hd->sop[k];
        const uint32_t* g = pool + (size_t)SEQ_WORDS * op.seq_base;
        uint2* run = (uint2*)(region + op.meta_off);
        for (uint32_t i = first; i < op.nruns; i += step) run[i] = make_uint2(g[SEQ_WORDS * i + 6], g[SEQ_WORDS * i + 7]);
    }
}

template <int NTH>
__device__ inline void fast_decode_chunk(FastWs<NTH>* ws, uint8_t* region, const FastParams& P, uint32_t c, uint32_t parity, PhaseClock& pc)
{
    const int tid = threadIdx.x;
    const ChunkPlan* pl = P.plans + c;
    // ---- stage the compressed chunk (async) while the plan is copied and the tables are built ----
    const bool parked = ws->stash_tag == c && !ws->stash_bad;   // the previous chunk's Huffman stage fetched this chunk's plan
    const uint32_t* stash = ws->sub_end;
    const PlanHead* hsrc = parked ? (const PlanHead*)stash : &pl->head;
    if (tid == 0) {
        const unsigned long long coff = parked ? ((unsigned long long)stash[ST_COFF + 1] << 32 | stash[ST_COFF]) : P.comp_off[c];
        const uint8_t* gsrc = P.comp + coff;
        const uint32_t phase = (uint32_t)((uintptr_t)gsrc & 15u);
        ws->status = RNSC_OK;
        bulk_load(region + REGION_HEAD, gsrc - phase, hsrc->stage_bytes, &ws->mbar);
    }
    // Next chunk index (classes with several CTAs per SM; the one-CTA class leaves it to a spare warp, below): three dependent
    // global round trips -- the ticket, the list length, the list entry. Thread 0 takes the ticket here and reads the entry
    // behind the barrier; the value is only stored once its share of the tables is built, so nobody waits for the round trips
    // (they used to sit between the barrier and thread 0's table entries, with the rest of the CTA waiting at the next barrier).
    constexpr bool CHAIN_LATE = NTH == 1024;
    uint32_t ticket = 0, next_c = FAST_NONE;
    if (!CHAIN_LATE && tid == 0) ticket = atomicAdd(P.counter, 1u);
    if (tid == 32) {      // where the samples go and how many (read by the splitter inverse)
        ws->n = parked ? stash[ST_N] : P.n[c];
        ws->sample_off = parked ? ((unsigned long long)stash[ST_SOFF + 1] << 32 | stash[ST_SOFF]) : P.sample_off[c];
    }
    uint2 run_v = make_uint2(0u, 0u); uint32_t run_off = FAST_NONE;      // this thread's run-table entry on its way from the sequence pool
    {
        const uint32_t* s = (const uint32_t*)hsrc; uint32_t* d = (uint32_t*)&ws->head;
        for (uint32_t i = tid; i < sizeof(PlanHead) / 4; i += NTH) d[i] = s[i];
        const uint32_t nhuf = hsrc->nhuf;
        s = parked ? stash + ST_HEAD_W : (const uint32_t*)pl->h; d = (uint32_t*)ws->h;
        for (uint32_t i = tid; i < (uint32_t)sizeof(HufPlan) / 4 * nhuf; i += NTH) d[i] = s[i];
        // run tables of the frames with sequences: {first literal index, shift} of every non-empty literal run. They are
        // first read in the second Huffman pass: with few sequences the one-CTA-per-SM class leaves the (global memory
        // latency of the) copy to its spare last warp, below
        // ... the other classes fetch a handful of runs (C5: the large values' high bytes) into registers here and store them
        // when the tables are built: the round trip to the sequence pool is not waited for at the barrier below
        if (hsrc->nseq > FAST_ORDERED_MAX) copy_run_tables(hsrc, P.seq_pool, region, tid, NTH);
        else if (NTH != 1024) {
            uint32_t idx = (uint32_t)tid;                         // at most FAST_ORDERED_MAX + 5 runs in all: one per thread
            for (uint32_t k = 0; k < hsrc->nsop; ++k) {
                const uint32_t nr = hsrc->sop[k].nruns;
                if (idx < nr) {
                    const uint32_t* g = P.seq_pool + (size_t)SEQ_WORDS * hsrc->sop[k].seq_base + SEQ_WORDS * idx;
                    run_v = make_uint2(g[6], g[7]); run_off = hsrc->sop[k].meta_off + 8u * idx;
                    break;
                }
                idx -= nr;
            }
        }
    }
    __syncthreads();
    if (tid == 0) ws->stash_bad = 0u;                // every thread has read it (`parked`); set again only behind later barriers
    // prefetch the next chunk index while this one is decoded
    // Three dependent global round trips. With four or two CTAs per SM thread 0 takes them here and the other CTAs cover
    // the stall; the one-CTA-per-SM class gives them to its last warp, which the planner keeps free of Huffman sub-sequences
    // (see below), so that nobody waits for them.
    if (!CHAIN_LATE && tid == 0) next_c = ticket < *P.count ? P.list[ticket] : FAST_NONE;
    const PlanHead& hd = ws->head;
    {   // all tables at once: the CTA's threads are split evenly over them
        const uint32_t nhuf = hd.nhuf;
        if (nhuf) {
            const int per = NTH / (int)nhuf, t = tid / per;
            if (t < (int)nhuf) fast_build_table((uint16_t*)(region + ws->h[t].tab_off), &ws->h[t], tid - t * per, per);
            // skip tables (all tables of a chunk have room for one, or none has); the scratch of the build lives in the
            // sub-sequence arrays, which the Huffman stage only uses later
            bool skip_on = ws->h[0].skip_off != 0;
            const uint32_t aux_cap = ((uint32_t)sizeof(ws->sub_end) + (uint32_t)sizeof(ws->sub_inc)) / nhuf & ~15u;
            for (uint32_t h = 0; h < nhuf; ++h) if (skip_aux_bytes(ws->h[h].log, ws->h[h].lmin) > aux_cap) skip_on = false;
            if (tid == 0) ws->skip_on = skip_on;
            if (skip_on) {
                uint8_t* aux = (uint8_t*)ws->sub_end + (size_t)aux_cap * (t < (int)nhuf ? t : 0);
                const HufPlan* hp = &ws->h[t < (int)nhuf ? t : 0];
                const uint16_t* tb = (const uint16_t*)(region + hp->tab_off);
                const uint32_t lmin = hp->lmin;
                __syncthreads();
                if (t < (int)nhuf) fast_build_aux(aux, tb, hp->log, lmin, tid - t * per, per);
                __syncthreads();
                if (t < (int)nhuf) fast_build_skip(region + hp->skip_off, aux, tb, hp->log, lmin, tid - t * per, per);
            }
        }
    }
    if (hd.lut_off) {      // this chunk has room for the replicated key-byte table: entry i of copy r in slot i * R + r
        constexpr uint32_t LR = key_lut_copies(NTH);
        const KeyLut* one = (const KeyLut*)(region - unpack2_lut_bytes(P.codec));
        KeyLut* rep = (KeyLut*)(region + hd.lut_off);
        for (uint32_t i = tid; i < 256u * LR; i += NTH) rep[i] = one[i / LR];
    }
    for (uint32_t i = 0; i < hd.nfill; ++i) { uint8_t* d = region + hd.fill[i].dst_off; for (uint32_t q = tid; q < hd.fill[i].len; q += NTH) d[q] = (uint8_t)hd.fill[i].value; }
    if (run_off != FAST_NONE) *(uint2*)(region + run_off) = run_v;
    if (!CHAIN_LATE && tid == 0) ws->next = next_c;
    mbar_wait(&ws->mbar, parity);
    __syncthreads();
    pc.mark(PH_STAGE);
    if (CHAIN_LATE && tid >= NTH - 32) {
        if (hd.nseq <= FAST_ORDERED_MAX) { copy_run_tables(&hd, P.seq_pool, region, (uint32_t)tid & 31u, 32u); __syncwarp(); }
    }
    // (the round trips for the next chunk index go to the warp in front of it: the planner's rounding of the sub-sequence
    // length leaves the last warps with Huffman sub-sequences mostly idle)
    if (CHAIN_LATE && tid == NTH - 64) { const uint32_t i = atomicAdd(P.counter, 1u); ws->next = i < *P.count ? P.list[i] : FAST_NONE; }
    // frames with a few sequences over raw / RLE literals: the last warp (which the planner kept free of Huffman
    // sub-sequences) executes them now, off the critical path
    if (hd.early && (tid >> 5) == NTH / 32 - 1)
        for (uint32_t k = 0; k < hd.nsop; ++k)
            if ((hd.early >> k) & 1u) fast_exec_sequences_ordered(hd.sop[k], P.seq_pool + (size_t)SEQ_WORDS * hd.sop[k].seq_base, region);
    if (hd.nst && fast_huffman<NTH>(ws, region, P, pc)) { if (tid == 0) ws->status = RNSC_ERR_ZSTD_DECODE; __syncthreads(); return; }
    if (hd.nsop) {
        if (hd.nseq <= FAST_ORDERED_MAX) {      // few sequences: one warp per remaining frame, in order
            const uint32_t late = ((1u << hd.nsop) - 1u) & ~hd.early, w = tid >> 5;
            if (late) {
                if (CHAIN_LATE && !hd.nst) __syncthreads();      // the run tables were copied by the last warp just above
                if ((late >> w) & 1u) fast_exec_sequences_ordered(hd.sop[w], P.seq_pool + (size_t)SEQ_WORDS * hd.sop[w].seq_base, region);
                __syncthreads();
            }
        } else fast_exec_sequences<NTH>(ws, hd, P.seq_pool, region, P.origin + (size_t)blockIdx.x * P.origin_stride);
    }
    if (tid == 0 && hd.view_status != RNSC_OK) ws->status = hd.view_status;
    __syncthreads();
    pc.mark(PH_SEQ);
}

}  // namespace rnsc
