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

constexpr uint32_t FAST_WARM_BITS = 128;   // speculative warm-up before a sub-sequence
constexpr uint32_t FAST_NONE = 0xFFFFFFFFu;

struct FastParams {
    const uint8_t* comp; const uint64_t* comp_off; const uint32_t* n; const uint64_t* sample_off;
    int16_t* samples; int32_t* status; int codec;
    const ChunkPlan* plans; const uint32_t* list; const uint32_t* count; uint32_t* counter;
    const uint32_t* seq_pool;    // the planner's sequence pool (SEQ_WORDS words per entry)
    unsigned long long* prof;    // optional (rnsc_set_profiling): 8 phase cycle counters of this size class, summed over CTAs
};

// Phase timing for tuning: thread 0 of every CTA adds the cycles since the previous mark to prof[phase].
enum { PH_STAGE = 0, PH_HUF1 = 1, PH_HUFSCAN = 2, PH_HUF2 = 3, PH_SEQ = 4, PH_UNPACK = 5, PH_OTHER = 6, PH_CHUNKS = 7 };
struct PhaseClock {
    unsigned long long* prof; long long t;
    __device__ __forceinline__ void start(unsigned long long* p) { prof = p; if (prof && threadIdx.x == 0) t = clock64(); }
    __device__ __forceinline__ void mark(int phase) { if (prof && threadIdx.x == 0) { const long long n = clock64(); atomicAdd(prof + phase, (unsigned long long)(n - t)); t = n; } }
};

template <int NTH>
struct FastWs {
    uint32_t sub_end[NTH], sub_inc[NTH];       // end position of each sub-sequence; in-warp inclusive scan of the symbol counts
    uint32_t warp_tot[NTH / 32];
    Unpack2Ws<NTH> u2;
    PlanHead head;
    HufPlan h[PLAN_MAX_HUF];
    uint32_t seq[PLAN_MAX_SEQ * SEQ_WORDS];    // copy of the chunk's sequences when there are at most PLAN_MAX_SEQ of them
    const uint32_t* sq[5];                     // sequences of SeqOp i: in ws->seq or in the global pool
    uint32_t nlong[2], longq[128];             // long matches of the current dependency level, copied warp-wide
    View view[5];
    int status; uint32_t next;
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

// ---- Huffman table from its recipe: each thread fills 8 consecutive entries with one 16-byte store --------
// `nthr` threads (t = 0 .. nthr-1) build this table.
__device__ inline void fast_build_table(uint16_t* tab, const HufPlan* hp, int t, int nthr)
{
    const uint32_t L = hp->log, size = 1u << L;
    for (uint32_t u0 = (uint32_t)t * 8; u0 < size; u0 += (uint32_t)nthr * 8) {
        uint32_t w = 1;
        while (hp->cls_start[w + 1] <= u0) ++w;
        uint32_t packed[4];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t u = u0 + k;
            uint32_t e = 0;
            if (u < size) {
                while (hp->cls_start[w + 1] <= u) ++w;
                const uint32_t sym = hp->sorted[hp->cls_sym[w] + ((u - hp->cls_start[w]) >> (w - 1))];
                e = ((L + 1 - w) << 8) | sym;
            }
            if (k & 1) packed[k >> 1] |= e << 16; else packed[k >> 1] = e;
        }
        if (size >= 8) *(V16*)(tab + u0) = V16{packed[0], packed[1], packed[2], packed[3]};
        else for (uint32_t k = 0; k < size; ++k) tab[k] = (uint16_t)(packed[k >> 1] >> (16 * (k & 1)));
    }
}

// 64 bits of a backward bitstream ending (exclusive) at absolute bit address A of the word array w, left aligned.
__device__ __forceinline__ void bits_ending_at(const uint32_t* w, uint32_t A, uint32_t& hi, uint32_t& lo)
{
    const uint32_t wi = (A - 1) >> 5, sh = (32u - (A & 31u)) & 31u;
    const uint32_t a = w[wi], b = w[wi - 1], c = w[wi - 2];
    hi = __funnelshift_l(b, a, sh); lo = __funnelshift_l(c, b, sh);
}

#define RNSC_HUF_STEP(hi, lo, tab, L, nbv, ev)                                   \
    do { ev = (tab)[(hi) >> (32 - (L))]; nbv = ev >> 8;                          \
         hi = __funnelshift_l(lo, hi, nbv); lo <<= nbv; } while (0)

// Advances from consumed-bit position p to the first symbol boundary >= lim; counts the symbols decoded.
__device__ __forceinline__ void huf_walk(const uint32_t* words, uint32_t abit, const uint16_t* tab, uint32_t L, uint32_t& p, uint32_t lim, uint32_t& cnt)
{
    uint32_t hi, lo, nb, e;
    while (p + 4 * HUF_LOG_MAX <= lim) {
        bits_ending_at(words, abit - p, hi, lo);
        uint32_t adv = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { RNSC_HUF_STEP(hi, lo, tab, L, nb, e); adv += nb; }
        p += adv; cnt += 4;
    }
    while (p < lim) {
        bits_ending_at(words, abit - p, hi, lo);
        RNSC_HUF_STEP(hi, lo, tab, L, nb, e);
        p += nb; cnt += 1;
    }
}

// Decodes every Huffman stream of the chunk (ws->head.st[]) from the staged chunk into the region. Three barriers
// in the common case. Returns 0 on success (same value in all threads).
template <int NTH>
__device__ inline int fast_huffman(FastWs<NTH>* ws, uint8_t* region, PhaseClock& pc)
{
    constexpr int NW = NTH / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* words = (const uint32_t*)region;
    const PlanHead& hd = ws->head;
    const uint32_t nst = hd.nst, sub = hd.sub;
    uint32_t j = 0;
    while (j + 1 < nst && (uint32_t)tid >= hd.st_first[j + 1]) ++j;
    const bool active = (uint32_t)tid < hd.st_first[nst];
    const uint32_t t0 = hd.st_first[j], k = (uint32_t)tid - t0;
    const StreamPlan sp = hd.st[j];
    const uint32_t T = sp.tbits, abit = sp.abit, L = sp.log;
    const uint16_t* tab = (const uint16_t*)(region + sp.tab_off);
    const uint32_t s0 = k * sub, lim = min(s0 + sub, T);
    uint32_t my_start = 0, my_end = 0, my_cnt = 0;
    if (active) {
        uint32_t p = s0 > FAST_WARM_BITS ? s0 - FAST_WARM_BITS : 0, dummy = 0;
        huf_walk(words, abit, tab, L, p, s0, dummy);
        my_start = p;
        huf_walk(words, abit, tab, L, p, lim, my_cnt);
        my_end = p;
    }
    pc.mark(PH_HUF1);
    uint32_t inc;
    for (int round = 0;; ++round) {
        inc = warp_inclusive_scan(active ? my_cnt : 0u);
        ws->sub_end[tid] = my_end; ws->sub_inc[tid] = inc;
        if (lane == 31) ws->warp_tot[warp] = inc;
        __syncthreads();
        // a sub-sequence must start where its predecessor ended; repair the (rare) ones that did not synchronise
        const bool bad = active && k > 0 && ws->sub_end[tid - 1] != my_start;
        const uint32_t want = bad ? ws->sub_end[tid - 1] : 0;
        if (!__syncthreads_or(bad)) break;
        if (round >= NTH) return 1;                  // cannot happen for a consistent stream: every round fixes at least one more sub-sequence
        if (bad) {
            uint32_t p = want; my_start = p; my_cnt = 0;
            if (p < lim) huf_walk(words, abit, tab, L, p, lim, my_cnt);
            my_end = p;
        }
    }
    // output offset inside my stream = CTA exclusive scan of the counts minus the scan value at the stream's first thread
    const uint32_t wt = lane < NW ? ws->warp_tot[lane] : 0u;
    const uint32_t wex = warp_inclusive_scan(wt) - wt;                        // lane w: symbols before warp w
    const uint32_t ex = __shfl_sync(0xffffffffu, wex, warp) + inc - (active ? my_cnt : 0u);
    const uint32_t base = __shfl_sync(0xffffffffu, wex, (int)(t0 >> 5)) + ((t0 & 31u) ? ws->sub_inc[t0 - 1] : 0u);
    const uint32_t off = ex - base;
    pc.mark(PH_HUFSCAN);
    bool bad = false;
    if (active && (uint32_t)tid + 1 == hd.st_first[j + 1]) bad = (my_end != T) || (off + my_cnt != sp.count);
    if (active && my_cnt && off < sp.count) {
        uint32_t p = my_start, rem = min(my_cnt, sp.count - off), hi, lo, nb, e;      // clamped: a corrupt stream cannot write outside its output
        uint32_t li = sp.lit0 + off;                                                   // literal index of my first symbol
        // Literal runs: without sequences the whole frame is one run at shift 0; with sequences run i (the literals in
        // front of match i, then the tail) lands at its final position out_pos[i] = lit_pos[i] + sum of earlier matches.
        const uint32_t* sq = sp.sop < 5 ? ws->sq[sp.sop] : ws->seq;
        const uint32_t nseq = sp.sop < 5 ? hd.sop[sp.sop].nseq : 0u;
        uint32_t seg = 0;
        if (nseq) {   // last run whose first literal index is <= li (runs may be empty)
            uint32_t lo_i = 0, hi_i = nseq;                                            // run nseq = tail, starts at lit_end
            const uint32_t lit_end = sq[SEQ_WORDS * (nseq - 1) + 3] + sq[SEQ_WORDS * (nseq - 1)];
            if (li >= lit_end) seg = nseq;
            else { while (lo_i + 1 < hi_i) { const uint32_t mid = (lo_i + hi_i) >> 1; if (sq[SEQ_WORDS * mid + 3] <= li) lo_i = mid; else hi_i = mid; } seg = lo_i; }
        }
        while (rem) {
            uint32_t take = rem, shift = 0;
            if (nseq) {
                const uint32_t last = nseq - 1;
                const uint32_t lit_end = sq[SEQ_WORDS * last + 3] + sq[SEQ_WORDS * last], out_end = sq[SEQ_WORDS * last + 4] + sq[SEQ_WORDS * last] + sq[SEQ_WORDS * last + 1];
                const uint32_t run_lit = seg < nseq ? sq[SEQ_WORDS * seg + 3] : lit_end, run_out = seg < nseq ? sq[SEQ_WORDS * seg + 4] : out_end;
                const uint32_t run_end = seg < nseq ? run_lit + sq[SEQ_WORDS * seg] : sp.regen;
                shift = run_out - run_lit;
                take = min(rem, run_end > li ? run_end - li : 0u);
                if (++seg > nseq + 1) break;                                        // cannot happen for a consistent plan
            }
            uint8_t* o = region + sp.out + li + shift;
            li += take; rem -= take;
            while (((uintptr_t)o & 3u) && take) { bits_ending_at(words, abit - p, hi, lo); RNSC_HUF_STEP(hi, lo, tab, L, nb, e); p += nb; *o++ = (uint8_t)e; --take; }
            while (take >= 4) {
                bits_ending_at(words, abit - p, hi, lo);
                uint32_t word = 0, adv = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) { RNSC_HUF_STEP(hi, lo, tab, L, nb, e); adv += nb; word |= (e & 0xFFu) << (8 * q); }
                *(uint32_t*)o = word; o += 4; p += adv; take -= 4;
            }
            while (take) { bits_ending_at(words, abit - p, hi, lo); RNSC_HUF_STEP(hi, lo, tab, L, nb, e); p += nb; *o++ = (uint8_t)e; --take; }
        }
    }
    const int r = __syncthreads_or(bad);
    pc.mark(PH_HUF2);
    return r;
}

// Executes the LZ sequences of one frame. Literal runs: Huffman literals were already written in place by
// fast_huffman; raw / RLE literal runs are copied / filled by all threads. Matches: level by level of the planner's
// dependency depth -- all matches of a level are independent of each other, one thread copies one match.
template <int NTH>
__device__ inline void fast_exec_sequences(FastWs<NTH>* ws, const SeqOp& op, const uint32_t* sq, uint8_t* region)
{
    const int tid = threadIdx.x;
    uint8_t* dst = region + op.dst_off;
    const uint32_t nseq = op.nseq;
    if (op.ltype != 2) {
        const uint32_t lit_end = sq[SEQ_WORDS * (nseq - 1) + 3] + sq[SEQ_WORDS * (nseq - 1)];                  // literals consumed by the sequences
        const uint32_t out_end = sq[SEQ_WORDS * (nseq - 1) + 4] + sq[SEQ_WORDS * (nseq - 1)] + sq[SEQ_WORDS * (nseq - 1) + 1];
        for (uint32_t i = 0; i <= nseq; ++i) {
            const uint32_t lp = i < nseq ? sq[SEQ_WORDS * i + 3] : lit_end, opos = i < nseq ? sq[SEQ_WORDS * i + 4] : out_end;
            const uint32_t len = i < nseq ? sq[SEQ_WORDS * i] : op.regen - lit_end;
            uint8_t* d = dst + opos;
            if (op.ltype == 1) for (uint32_t q = tid; q < len; q += NTH) d[q] = (uint8_t)op.lit;
            else { const uint8_t* s = region + op.lit + lp; for (uint32_t q = tid; q < len; q += NTH) d[q] = s[q]; }
        }
        __syncthreads();
    }
    // short matches: one thread each; long ones are queued and copied by a whole warp each
    constexpr uint32_t LONG = 24, QCAP = 128;
    if (tid == 0) ws->nlong[0] = 0;
    __syncthreads();
    for (uint32_t lvl = 0; lvl <= op.max_depth; ++lvl) {
        const uint32_t par = lvl & 1u;
        if (tid == 0) ws->nlong[par ^ 1u] = 0;
        for (uint32_t i = tid; i < nseq; i += NTH) {
            if (sq[SEQ_WORDS * i + 5] != lvl) continue;
            const uint32_t ml = sq[SEQ_WORDS * i + 1], of = sq[SEQ_WORDS * i + 2];
            if (ml > LONG) { const uint32_t k = atomicAdd(&ws->nlong[par], 1u); if (k < QCAP) { ws->longq[k] = i; continue; } }
            uint8_t* d = dst + sq[SEQ_WORDS * i + 4] + sq[SEQ_WORDS * i];
            const uint8_t* s = d - of;
            for (uint32_t q = 0; q < ml; ++q) d[q] = s[q];             // ascending byte order: correct for overlapping (offset < length) matches
        }
        __syncthreads();
        const uint32_t nl = min(ws->nlong[par], QCAP);
        if (nl) {
            const uint32_t lane = tid & 31u;
            for (uint32_t k = tid >> 5; k < nl; k += NTH / 32) {
                const uint32_t i = ws->longq[k];
                const uint32_t ml = sq[SEQ_WORDS * i + 1], of = sq[SEQ_WORDS * i + 2];
                uint8_t* d = dst + sq[SEQ_WORDS * i + 4] + sq[SEQ_WORDS * i];
                const uint8_t* sp = d - of;
                if (of >= 32) { for (uint32_t k0 = 0; k0 < ml; k0 += 32) { const uint32_t q = k0 + lane; if (q < ml) d[q] = sp[q]; __syncwarp(); } }
                else { for (uint32_t q = lane; q < ml; q += 32) d[q] = sp[q % of]; }
            }
            __syncthreads();
        }
    }
}

// Decodes chunk c into stream views (ws->view) and ws->status. Must be entered by the whole CTA after a barrier.
template <int NTH>
__device__ inline void fast_decode_chunk(FastWs<NTH>* ws, uint8_t* region, const FastParams& P, uint32_t c, uint32_t parity, PhaseClock& pc)
{
    const int tid = threadIdx.x;
    const ChunkPlan* pl = P.plans + c;
    // ---- stage the compressed chunk (async) while the plan is copied and the tables are built ----
    if (tid == 0) {
        const uint8_t* gsrc = P.comp + P.comp_off[c];
        const uint32_t phase = (uint32_t)((uintptr_t)gsrc & 15u);
        ws->status = RNSC_OK;
        bulk_load(region + REGION_HEAD, gsrc - phase, pl->head.stage_bytes, &ws->mbar);
    }
    {
        const uint32_t* s = (const uint32_t*)&pl->head; uint32_t* d = (uint32_t*)&ws->head;
        for (uint32_t i = tid; i < sizeof(PlanHead) / 4; i += NTH) d[i] = s[i];
        const uint32_t nhuf = pl->head.nhuf, nseq = pl->head.nseq;
        s = (const uint32_t*)pl->h; d = (uint32_t*)ws->h;
        for (uint32_t i = tid; i < (uint32_t)sizeof(HufPlan) / 4 * nhuf; i += NTH) d[i] = s[i];
        // sequences: copied to shared memory when few (the usual case), read from the global pool otherwise
        if (nseq <= PLAN_MAX_SEQ) {
            uint32_t base = 0;
            for (uint32_t k = 0; k < pl->head.nsop; ++k) {
                const uint32_t cnt = pl->head.sop[k].nseq;
                const uint32_t* g = P.seq_pool + (size_t)SEQ_WORDS * pl->head.sop[k].seq_base;
                for (uint32_t i = tid; i < SEQ_WORDS * cnt; i += NTH) ws->seq[SEQ_WORDS * base + i] = g[i];
                if (tid == 0) ws->sq[k] = ws->seq + SEQ_WORDS * base;
                base += cnt;
            }
        } else if (tid < 5 && (uint32_t)tid < pl->head.nsop) ws->sq[tid] = P.seq_pool + (size_t)SEQ_WORDS * pl->head.sop[tid].seq_base;
    }
    __syncthreads();
    // prefetch the next chunk index while this one is decoded
    if (tid == 0) { const uint32_t i = atomicAdd(P.counter, 1u); ws->next = i < *P.count ? P.list[i] : FAST_NONE; }
    const PlanHead& hd = ws->head;
    {   // all tables at once: the CTA's threads are split evenly over them
        const uint32_t nhuf = hd.nhuf;
        if (nhuf) {
            const int per = NTH / (int)nhuf, t = tid / per;
            if (t < (int)nhuf) fast_build_table((uint16_t*)(region + ws->h[t].tab_off), &ws->h[t], tid - t * per, per);
        }
    }
    for (uint32_t i = 0; i < hd.nfill; ++i) { uint8_t* d = region + hd.fill[i].dst_off; for (uint32_t q = tid; q < hd.fill[i].len; q += NTH) d[q] = (uint8_t)hd.fill[i].value; }
    mbar_wait(&ws->mbar, parity);
    __syncthreads();
    pc.mark(PH_STAGE);
    if (hd.nst && fast_huffman<NTH>(ws, region, pc)) { if (tid == 0) ws->status = RNSC_ERR_ZSTD_DECODE; __syncthreads(); return; }
    for (uint32_t i = 0; i < hd.nsop; ++i) fast_exec_sequences<NTH>(ws, hd.sop[i], ws->sq[i], region);
    if (tid < 5) { ws->view[tid].p = region + hd.view[tid].off; ws->view[tid].len = hd.view[tid].len; }
    if (tid == 0 && hd.view_status != RNSC_OK) ws->status = hd.view_status;
    __syncthreads();
    pc.mark(PH_SEQ);
}

}  // namespace rnsc
