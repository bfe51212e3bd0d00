// fast_decode.cuh -- the fast whole-chunk decoder: consumes a ChunkPlan (plan.cuh), so nothing serial is left.
//
// One persistent CTA owns one chunk at a time:
//   1. one elected thread starts a bulk asynchronous copy (TMA, cp.async.bulk + mbarrier) of the whole compressed
//      chunk into shared memory while all threads copy the plan and build the Huffman decoding tables;
//   2. every Huffman-coded literal section of the chunk (typically keys and M of a C5 chunk: 8 backward
//      bitstreams) is cut into NT sub-sequences that are decoded at once from shared memory: speculative start
//      WARM bits early (self-synchronisation), chain check, CTA scan of the symbol counts, second pass writes;
//   3. raw frames are used in place, RLE frames filled, the few LZ sequences executed (literal runs by all
//      threads, matches in order by one warp);
//   4. the splitter inverse (svb_codec.cuh) turns the streams, all in shared memory, into int16 samples.
#pragma once
#include "plan.cuh"
#include "svb_codec.cuh"
#include "svb_unpack2.cuh"

namespace rnsc {

constexpr uint32_t FAST_WARM_BITS = 128;   // speculative warm-up before a sub-sequence
constexpr uint32_t FAST_SUB_MIN = 384;     // minimum sub-sequence length in bits
constexpr int FAST_MAX_STREAMS = 8;        // two Huffman sections x 4 streams per round

struct FastParams {
    const uint8_t* comp; const uint64_t* comp_off; const uint32_t* n; const uint64_t* sample_off;
    int16_t* samples; int32_t* status; int codec;
    const ChunkPlan* plans; const uint32_t* list; const uint32_t* count; uint32_t* counter;
    uint32_t region_bytes;
};

template <int NTH>
struct FastWs {
    uint16_t huf[2][1 << HUF_LOG_MAX];                     // (nbBits << 8) | symbol
    uint32_t sub_start[NTH], sub_end[NTH], sub_cnt[NTH];
    uint32_t scan_a[NTH / 32 + 1];
    Unpack2Ws<NTH> u2;
    FramePlan f[5];
    HufPlan h[2];
    uint32_t seq[PLAN_MAX_SEQ * 3];
    // chain geometry of the current Huffman round
    uint32_t st_first[FAST_MAX_STREAMS + 1], st_tbits[FAST_MAX_STREAMS], st_abit[FAST_MAX_STREAMS], st_count[FAST_MAX_STREAMS],
             st_out[FAST_MAX_STREAMS], st_tab[FAST_MAX_STREAMS], st_log[FAST_MAX_STREAMS], st_base[FAST_MAX_STREAMS];
    uint32_t nst, sub;
    View view[5];
    int status; uint32_t chunk;
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
template <int NTH>
__device__ inline void fast_build_table(uint16_t* tab, const HufPlan* hp)
{
    const uint32_t L = hp->log, size = 1u << L;
    for (uint32_t u0 = threadIdx.x * 8; u0 < size; u0 += NTH * 8) {
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

// One round: decodes the streams described in ws->st_* from the staged chunk into the region. Returns 0 on
// success (same value in all threads).
template <int NTH>
__device__ inline int fast_huffman_round(FastWs<NTH>* ws, const uint32_t* words, uint8_t* region)
{
    const int tid = threadIdx.x;
    const uint32_t nst = ws->nst, sub = ws->sub;
    // my stream / sub-sequence
    uint32_t j = 0;
    while (j + 1 < nst && (uint32_t)tid >= ws->st_first[j + 1]) ++j;
    const bool active = (uint32_t)tid < ws->st_first[nst];
    const uint32_t k = (uint32_t)tid - ws->st_first[j];
    const uint32_t T = ws->st_tbits[j], abit = ws->st_abit[j], L = ws->st_log[j];
    const uint16_t* tab = ws->huf[ws->st_tab[j]];
    uint32_t my_start = 0, my_end = 0, my_cnt = 0;
    const uint32_t s0 = k * sub, lim = min(s0 + sub, T);
    if (active) {
        uint32_t p = s0 > FAST_WARM_BITS ? s0 - FAST_WARM_BITS : 0, dummy = 0;
        huf_walk(words, abit, tab, L, p, s0, dummy);
        my_start = p;
        huf_walk(words, abit, tab, L, p, lim, my_cnt);
        my_end = p;
    }
    ws->sub_start[tid] = my_start; ws->sub_end[tid] = my_end; ws->sub_cnt[tid] = my_cnt;
    __syncthreads();
    // verify / repair the chain: a sub-sequence must start where its predecessor ended
    for (int round = 0; round < NTH; ++round) {
        const bool bad = active && k > 0 && ws->sub_end[tid - 1] != my_start;
        const uint32_t want = bad ? ws->sub_end[tid - 1] : 0;
        if (!__syncthreads_or(bad)) break;
        if (bad) {
            uint32_t p = want; my_start = p; my_cnt = 0;
            if (p < lim) huf_walk(words, abit, tab, L, p, lim, my_cnt);
            my_end = p;
            ws->sub_start[tid] = my_start; ws->sub_end[tid] = my_end; ws->sub_cnt[tid] = my_cnt;
        }
        __syncthreads();
    }
    uint32_t total;
    const uint32_t ex = block_scan_1b<uint32_t, NTH>(active ? my_cnt : 0u, ws->scan_a, &total);
    if (active && k == 0) ws->st_base[j] = ex;
    __syncthreads();
    const uint32_t off = ex - ws->st_base[j];
    bool bad = false;
    if (active && (uint32_t)tid + 1 == ws->st_first[j + 1]) bad = (my_end != T) || (off + my_cnt != ws->st_count[j]);
    if (__syncthreads_or(bad)) return 1;
    if (active && my_cnt) {
        uint8_t* o = region + ws->st_out[j] + off;
        uint32_t p = my_start, rem = my_cnt, hi, lo, nb, e;
        while (((uintptr_t)o & 3u) && rem) { bits_ending_at(words, abit - p, hi, lo); RNSC_HUF_STEP(hi, lo, tab, L, nb, e); p += nb; *o++ = (uint8_t)e; --rem; }
        while (rem >= 4) {
            bits_ending_at(words, abit - p, hi, lo);
            uint32_t word = 0, adv = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) { RNSC_HUF_STEP(hi, lo, tab, L, nb, e); adv += nb; word |= (e & 0xFFu) << (8 * q); }
            *(uint32_t*)o = word; o += 4; p += adv; rem -= 4;
        }
        while (rem) { bits_ending_at(words, abit - p, hi, lo); RNSC_HUF_STEP(hi, lo, tab, L, nb, e); p += nb; *o++ = (uint8_t)e; --rem; }
    }
    __syncthreads();
    return 0;
}

// Executes the LZ sequences of one frame: literal runs by all threads, matches in order by warp 0.
template <int NTH>
__device__ inline void fast_exec_sequences(FastWs<NTH>* ws, const FramePlan& fp, const uint8_t* chunk, uint8_t* region)
{
    const int tid = threadIdx.x;
    const uint32_t* sq = ws->seq + 3 * fp.seq_base;
    uint8_t* dst = region + fp.dst_off;
    uint32_t* lit_pos = ws->sub_start; uint32_t* out_pos = ws->sub_end;      // free between Huffman rounds
    if (tid == 0) {
        uint32_t lp = 0, op = 0;
        for (uint32_t i = 0; i < fp.nseq; ++i) { lit_pos[i] = lp; out_pos[i] = op; lp += sq[3 * i]; op += sq[3 * i] + sq[3 * i + 1]; }
        lit_pos[fp.nseq] = lp; out_pos[fp.nseq] = op;
    }
    __syncthreads();
    if (fp.ltype == 2) {
        // Huffman literals sit at dst[0 .. regen): move run i up by the match bytes in front of it, last run first.
        // Source and destination overlap, so every block is read to registers, then written, between barriers.
        constexpr uint32_t B = NTH * 16;
        for (uint32_t i = fp.nseq; i >= 1; --i) {
            const uint32_t len = i < fp.nseq ? sq[3 * i] : fp.regen - lit_pos[fp.nseq];
            const uint8_t* s = dst + lit_pos[i]; uint8_t* d = dst + out_pos[i];
            if (s == d) continue;
            for (uint32_t done = 0; done < len; done += B) {
                const uint32_t blen = min(B, len - done), b0 = len - done - blen;     // blocks from the top down
                uint8_t t[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) { const uint32_t q = tid + k * NTH; t[k] = q < blen ? s[b0 + q] : (uint8_t)0; }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < 16; ++k) { const uint32_t q = tid + k * NTH; if (q < blen) d[b0 + q] = t[k]; }
                __syncthreads();
            }
        }
    } else {
        for (uint32_t i = 0; i <= fp.nseq; ++i) {
            const uint32_t len = i < fp.nseq ? sq[3 * i] : fp.regen - lit_pos[fp.nseq];
            uint8_t* d = dst + out_pos[i];
            if (fp.ltype == 1) for (uint32_t q = tid; q < len; q += NTH) d[q] = (uint8_t)fp.lit_off;
            else { const uint8_t* s = chunk + fp.lit_off + lit_pos[i]; for (uint32_t q = tid; q < len; q += NTH) d[q] = s[q]; }
        }
    }
    __syncthreads();
    if (tid < 32) {
        for (uint32_t i = 0; i < fp.nseq; ++i) {
            const uint32_t ml = sq[3 * i + 1], of = sq[3 * i + 2], op = out_pos[i] + sq[3 * i];
            if (of >= 32) { for (uint32_t k0 = 0; k0 < ml; k0 += 32) { const uint32_t q = k0 + tid; if (q < ml) dst[op + q] = dst[op + q - of]; __syncwarp(); } }
            else { const uint8_t* sp = dst + op - of; for (uint32_t q = tid; q < ml; q += 32) dst[op + q] = sp[q % of]; }
            __syncwarp();
        }
    }
    __syncthreads();
}

template <int NTH>
__device__ inline void fast_decode_chunk(FastWs<NTH>* ws, uint8_t* region, const FastParams& P, uint32_t c, uint32_t parity)
{
    const int tid = threadIdx.x;
    const ChunkPlan* pl = P.plans + c;
    const int nf = codec_frames(P.codec);
    const uint8_t* gsrc = P.comp + P.comp_off[c];
    const uint32_t n = P.n[c];
    const uint32_t phase = pl->phase, nhuf = pl->nhuf;
    uint8_t* chunk = region + REGION_HEAD + phase;             // chunk byte i lives at chunk[i]
    // ---- stage the compressed chunk (async) while the plan is copied and the tables are built ----
    if (tid == 0) { ws->status = RNSC_OK; bulk_load(region + REGION_HEAD, gsrc - phase, pl->stage_bytes, &ws->mbar); }
    {
        const uint32_t* s = (const uint32_t*)pl->f; uint32_t* d = (uint32_t*)ws->f;
        for (uint32_t i = tid; i < sizeof(FramePlan) * 5 / 4; i += NTH) d[i] = s[i];
        const uint32_t hw = (uint32_t)sizeof(HufPlan) / 4 * min(nhuf, 2u);
        s = (const uint32_t*)pl->h; d = (uint32_t*)ws->h;
        for (uint32_t i = tid; i < hw; i += NTH) d[i] = s[i];
        const uint32_t sw = 3 * pl->nseq;
        for (uint32_t i = tid; i < sw; i += NTH) ws->seq[i] = pl->seq[i];
    }
    __syncthreads();
    for (uint32_t r0 = 0; r0 < nhuf; r0 += 2) {
        const uint32_t nt = min(2u, nhuf - r0);
        if (r0) {     // later rounds (rare): fetch their recipes
            __syncthreads();
            const uint32_t* s = (const uint32_t*)(pl->h + r0); uint32_t* d = (uint32_t*)ws->h;
            for (uint32_t i = tid; i < (uint32_t)sizeof(HufPlan) / 4 * nt; i += NTH) d[i] = s[i];
            __syncthreads();
        }
        for (uint32_t t = 0; t < nt; ++t) fast_build_table<NTH>(ws->huf[t], &ws->h[t]);
        if (tid == 0) {
            // streams of this round's Huffman sections -> sub-sequence geometry
            uint32_t nst = 0, ttot = 0;
            for (int f = 0; f < nf; ++f) {
                const FramePlan& fp = ws->f[f];
                if (fp.kind != FK_COMP || fp.ltype != 2 || fp.huf < r0 || fp.huf >= r0 + nt) continue;
                const uint32_t q = (fp.regen + 3) / 4, out = fp.dst_off;
                for (uint32_t j = 0; j < fp.nstreams; ++j) {
                    ws->st_tbits[nst] = fp.s_tbits[j];
                    ws->st_abit[nst] = 8 * (REGION_HEAD + phase + fp.s_beg[j]) + fp.s_tbits[j];
                    ws->st_count[nst] = fp.nstreams == 1 ? fp.regen : (j < 3 ? q : fp.regen - 3 * q);
                    ws->st_out[nst] = out + (fp.nstreams == 1 ? 0 : j * q);
                    ws->st_tab[nst] = fp.huf - r0; ws->st_log[nst] = ws->h[fp.huf - r0].log;
                    if (fp.s_tbits[j] == 0 && ws->st_count[nst] != 0) ws->status = RNSC_ERR_ZSTD_DECODE;   // nothing to decode from
                    ttot += fp.s_tbits[j]; ++nst;
                }
            }
            uint32_t sub = (ttot + (NTH - FAST_MAX_STREAMS) - 1) / (NTH - FAST_MAX_STREAMS);
            if (sub < FAST_SUB_MIN) sub = FAST_SUB_MIN;
            uint32_t first = 0;
            for (uint32_t s = 0; s < nst; ++s) { ws->st_first[s] = first; first += (ws->st_tbits[s] + sub - 1) / sub; }
            ws->st_first[nst] = first; ws->nst = nst; ws->sub = sub;
        }
        __syncthreads();
        if (r0 == 0) mbar_wait(&ws->mbar, parity);
        if (ws->nst && fast_huffman_round<NTH>(ws, (const uint32_t*)region, region)) { if (tid == 0) ws->status = RNSC_ERR_ZSTD_DECODE; }
        __syncthreads();
        if (ws->status != RNSC_OK) break;
    }
    if (nhuf == 0) mbar_wait(&ws->mbar, parity);
    // ---- RLE frames, RLE literals, sequences ----
    if (ws->status == RNSC_OK) {
        for (int f = 0; f < nf; ++f) {
            const FramePlan& fp = ws->f[f];
            if (fp.kind == FK_RLE) { uint8_t* d = region + fp.dst_off; for (uint32_t q = tid; q < fp.fcs; q += NTH) d[q] = (uint8_t)fp.data_off; }
            else if (fp.kind == FK_COMP && fp.nseq == 0 && fp.ltype == 1) { uint8_t* d = region + fp.dst_off; for (uint32_t q = tid; q < fp.fcs; q += NTH) d[q] = (uint8_t)fp.lit_off; }
            else if (fp.kind == FK_COMP && fp.nseq) fast_exec_sequences<NTH>(ws, fp, chunk, region);
        }
    }
    __syncthreads();
    // ---- stream views (A.8) ----
    if (tid == 0 && ws->status == RNSC_OK) {
        const uint32_t kl = key_len(P.codec, n);
        for (int k = 0; k < 5; ++k) { ws->view[k].p = chunk; ws->view[k].len = 0; }
        auto fptr = [&](int f) -> const uint8_t* {
            const FramePlan& fp = ws->f[f];
            if (fp.kind == FK_RAW) return chunk + fp.data_off;
            if (fp.kind == FK_COMP && fp.ltype == 0 && fp.nseq == 0) return chunk + fp.lit_off;
            return region + fp.dst_off;
        };
        if (nf == 1) {
            if (ws->f[0].fcs < kl) ws->status = RNSC_ERR_REMAINING;
            else { ws->view[0].p = fptr(0); ws->view[0].len = kl; ws->view[1].p = fptr(0) + kl; ws->view[1].len = ws->f[0].fcs - kl; }
        } else {
            if (ws->f[0].fcs != kl) ws->status = RNSC_ERR_REMAINING;
            for (int f = 0; f < nf; ++f) { ws->view[f].p = fptr(f); ws->view[f].len = ws->f[f].fcs; }
        }
        if (n == 0) { uint32_t tot = 0; for (int f = 0; f < nf; ++f) tot += ws->f[f].fcs; if (tot) ws->status = RNSC_ERR_REMAINING; }
    }
    __syncthreads();
}

}  // namespace rnsc
