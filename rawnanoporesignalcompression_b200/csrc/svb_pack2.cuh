// svb_pack2.cuh -- splitter (int16 samples -> streams), second generation.
//
// Same arithmetic as svb_pack (svb_codec.cuh; reference encode_scalar_* loops cited there), organised for
// instruction count and for wide memory operations:
//   * a thread owns P2_SPT = 16 consecutive samples of the chunk; they come from global memory as three 16-byte
//     aligned vector loads (the third overlaps the neighbour's first and hits L1) that are shifted into place in
//     registers whatever the chunk's alignment -- no shared-memory sample tile;
//   * pass 1 (stream sizes) has no scan and no barrier per tile: class counts ride in registers and are reduced once
//     per chunk;
//   * pass 2 scans the per-thread unit counts with ONE barrier (two 32-bit words: nibble|byte units, large values),
//     scatters the units into a staging area that is laid out with the same 16-byte phase as the global destination
//     of each stream, and flushes it with 128-bit stores (bytes only at the ragged ends); the next tile's scan barrier
//     doubles as the "staging may be overwritten" barrier: two barriers per tile in all;
//   * the nibble stream (VBZ0/C4/C5) is staged one nibble per byte and packed while it is flushed; an odd trailing
//     nibble travels to the next tile through a double-buffered shared word.
#pragma once
#include "block_prims.cuh"
#include "svb_codec.cuh"
#include "svb_unpack2.cuh"

namespace rnsc {

constexpr int P2_SPT = 16;
#ifndef RNSC_P2_NT
#define RNSC_P2_NT 128
#endif
constexpr int P2_NT = RNSC_P2_NT;                // threads per CTA of the second-generation splitter kernels (128: 721 GB/s C5, 256: 701, 512: 632, 64: 589)
constexpr int P2_TILE = P2_NT * P2_SPT;
constexpr int P2_STAGE_BYTES = 4 * P2_TILE + 128;   // VBZ0: up to 4 nibbles (bytes here) per sample; C4/C5: 4 streams of <= TILE units

template <int NTH>
struct Pack2Ws {
    uint32_t sa[2][NTH / 32], sb[2][NTH / 32];   // warp totals of the two scan words, double-buffered
    uint32_t carry[2];                           // odd trailing nibble of a tile, double-buffered
    uint32_t red[3][NTH / 32];                   // pass 1: final reduction
    uint8_t* dst[5];
};

// The thread's 16 samples [i0, i0 + 16) of the chunk at `in` (any 2-byte alignment), zero-extended to 32 bits, and the
// sample before them (0 at the chunk start: svb16/encode_scalar.hpp:20). Samples past the chunk end repeat the last
// valid one so that their delta is zero. Only 16-byte blocks that hold at least one sample of the chunk are read.
__device__ __forceinline__ void load16(const int16_t* in, uint32_t n, uint32_t i0, uint32_t cnt, uint32_t s[P2_SPT], uint32_t& prev)
{
    const uint32_t ph = (uint32_t)(((uintptr_t)in >> 1) & 7);
    const V16* va = (const V16*)(in - ph) + (i0 >> 3);
    const uint32_t vend = ph + n;                                   // virtual index one past the last sample
    const uint32_t v0 = i0;                                         // virtual index of the first vector
    V16 a = {0, 0, 0, 0}, b = a, c = a;
    if (v0 < vend) a = va[0];
    if (v0 + 8 < vend) b = va[1];
    if (v0 + 16 < vend && ph) c = va[2];
    const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    uint32_t x[8];
    const uint32_t sh = (ph & 1u) << 4;
    switch (ph >> 1) {
    case 0:
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __funnelshift_r(w[i], w[i + 1], sh);
        break;
    case 1:
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __funnelshift_r(w[i + 1], w[i + 2], sh);
        break;
    case 2:
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __funnelshift_r(w[i + 2], w[i + 3], sh);
        break;
    default:
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __funnelshift_r(w[i + 3], w[i + 4], sh);
        break;
    }
    prev = (i0 && cnt) ? (uint32_t)(uint16_t)in[i0 - 1] : 0u;
#pragma unroll
    for (int j = 0; j < P2_SPT; ++j) s[j] = (j & 1) ? (x[j >> 1] >> 16) : (x[j >> 1] & 0xFFFFu);
    if (cnt < (uint32_t)P2_SPT) {
        uint32_t last = prev;
#pragma unroll
        for (int j = 0; j < P2_SPT; ++j) { if ((uint32_t)j < cnt) last = s[j]; else s[j] = last; }
    }
}

__device__ __forceinline__ void prefetch_l1(const void* p)
{
#ifndef RNSC_EMUL
    asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
#else
    (void)p;
#endif
}

// zigzag of the 16-bit delta between two samples (zero-extended int16 bit patterns): svb16/encode_scalar.hpp:11-14,41-43
__device__ __forceinline__ uint32_t delta_zigzag(uint32_t cur, uint32_t prev)
{
    const int32_t d = (int32_t)(int16_t)(uint16_t)(cur - prev);
    return (uint32_t)((d << 1) ^ (d >> 31));                      // in [0, 65535]: no mask needed on a sign-extended delta
}

// One sample's class into the key bits (one = 1 << j or 1 << 2j) and, for VBZ0/C5 with WANT_V, the class offset out of zz.
// Three compares, then predicated adds only: the key field receives +1 per threshold passed (C5.hpp:73-127, C4.hpp:69-121).
template <int CODEC, bool WANT_V>
__device__ __forceinline__ void classify_one(uint32_t& zz, uint32_t& kb, uint32_t one)
{
    constexpr bool TWO = (CODEC == RNSC_VBZ0 || CODEC == RNSC_C4 || CODEC == RNSC_C5);
#ifdef RNSC_EMUL
    if (!TWO) { if (zz >= 256u) kb += one; }
    else if (CODEC == RNSC_C4) kb += one * ((zz != 0u) + (zz >= 16u) + (zz >= 256u));
    else { const uint32_t a = zz != 0u, b = zz > 16u, c = zz > 272u; kb += one * (a + b + c); if (WANT_V) zz -= a + 16u * b + 256u * c; }
#else
    if (!TWO) asm("{\n.reg .pred c;\nsetp.gt.u32 c, %1, 255;\n@c add.u32 %0, %0, %2;\n}" : "+r"(kb) : "r"(zz), "r"(one));
    else if (CODEC == RNSC_C4)
        asm("{\n.reg .pred a, b, c;\nsetp.ne.u32 a, %1, 0;\nsetp.gt.u32 b, %1, 15;\nsetp.gt.u32 c, %1, 255;\n"
            "@a add.u32 %0, %0, %2;\n@b add.u32 %0, %0, %2;\n@c add.u32 %0, %0, %2;\n}" : "+r"(kb) : "r"(zz), "r"(one));
    else if (!WANT_V)
        asm("{\n.reg .pred a, b, c;\nsetp.ne.u32 a, %1, 0;\nsetp.gt.u32 b, %1, 16;\nsetp.gt.u32 c, %1, 272;\n"
            "@a add.u32 %0, %0, %2;\n@b add.u32 %0, %0, %2;\n@c add.u32 %0, %0, %2;\n}" : "+r"(kb) : "r"(zz), "r"(one));
    else
        asm("{\n.reg .pred a, b, c;\nsetp.ne.u32 a, %1, 0;\nsetp.gt.u32 b, %1, 16;\nsetp.gt.u32 c, %1, 272;\n"
            "@a add.u32 %0, %0, %2;\n@b add.u32 %0, %0, %2;\n@c add.u32 %0, %0, %2;\n"
            "@a sub.u32 %1, %1, 1;\n@b sub.u32 %1, %1, 16;\n@c sub.u32 %1, %1, 256;\n}" : "+r"(kb), "+r"(zz) : "r"(one));
#endif
}

// Pass 1: thresholds passed by one sample, counted straight into na/nb/nc (no key bits needed).
template <int CODEC>
__device__ __forceinline__ void count_one(uint32_t zz, uint32_t& na, uint32_t& nb, uint32_t& nc)
{
    constexpr bool TWO = (CODEC == RNSC_VBZ0 || CODEC == RNSC_C4 || CODEC == RNSC_C5);
    constexpr uint32_t TB = CODEC == RNSC_C4 ? 15u : 16u, TC = CODEC == RNSC_C4 ? 255u : 272u;
#ifdef RNSC_EMUL
    if (!TWO) nc += zz > 255u;
    else { na += zz != 0u; nb += zz > TB; nc += zz > TC; }
#else
    if (!TWO) asm("{\n.reg .pred c;\nsetp.gt.u32 c, %1, 255;\n@c add.u32 %0, %0, 1;\n}" : "+r"(nc) : "r"(zz));
    else asm("{\n.reg .pred a, b, c;\nsetp.ne.u32 a, %3, 0;\nsetp.gt.u32 b, %3, %4;\nsetp.gt.u32 c, %3, %5;\n"
             "@a add.u32 %0, %0, 1;\n@b add.u32 %1, %1, 1;\n@c add.u32 %2, %2, 1;\n}" : "+r"(na), "+r"(nb), "+r"(nc) : "r"(zz), "n"(TB), "n"(TC));
#endif
}

// delta + zigzag + classify of the thread's 16 samples. Returns the key bits (1 or 2 per sample, sample j at bit j or
// 2j); v[j] = the payload value of sample j (class offset removed for VBZ0/C5).
template <int CODEC>
__device__ __forceinline__ uint32_t classify16(const uint32_t s[P2_SPT], uint32_t prev, uint32_t v[P2_SPT])
{
    constexpr bool TWO = (CODEC == RNSC_VBZ0 || CODEC == RNSC_C4 || CODEC == RNSC_C5);
    uint32_t kb = 0;
#pragma unroll
    for (int j = 0; j < P2_SPT; ++j) {
        uint32_t zz = delta_zigzag(s[j], prev); prev = s[j];
        classify_one<CODEC, true>(zz, kb, TWO ? 1u << (2 * j) : 1u << j);
        v[j] = zz;
    }
    return kb;
}

// unit counts of the thread from its key bits (cnt = valid samples of the thread)
template <int CODEC>
__device__ __forceinline__ void unit_counts(uint32_t kb, uint32_t cnt, uint32_t& c1, uint32_t& c2, uint32_t& c3)
{
    constexpr bool TWO = (CODEC == RNSC_VBZ0 || CODEC == RNSC_C4 || CODEC == RNSC_C5);
    c1 = c2 = c3 = 0;
    if (TWO) {
        const uint32_t lo = kb & 0x55555555u, hi = (kb >> 1) & 0x55555555u;
        const uint32_t n1 = __popc(lo & ~hi), n2 = __popc(hi & ~lo), n3 = __popc(lo & hi);
        if (CODEC == RNSC_VBZ0) c1 = n1 + 2 * n2 + 4 * n3; else { c1 = n1; c2 = n2; c3 = n3; }
    } else {
        const uint32_t nl = __popc(kb);
        if (CODEC == RNSC_VBZ || CODEC == RNSC_C1) c1 = cnt + nl;
        else if (CODEC == RNSC_C2) { c1 = cnt; c2 = nl; }
        else { c2 = nl; c1 = cnt - nl; }
    }
}

__device__ __forceinline__ void stream_sizes(int codec, uint32_t n, uint32_t run1, uint32_t run2, uint32_t run3, uint32_t* sizes)
{
    uint32_t s[5] = {key_len(codec, n), 0, 0, 0, 0};
    if (codec == RNSC_VBZ || codec == RNSC_C1) s[1] = run1;
    else if (codec == RNSC_VBZ0) s[1] = (run1 + 1) >> 1;
    else if (codec == RNSC_C2) { s[1] = run1; s[2] = run2; }
    else if (codec == RNSC_C3) { s[1] = run1; s[2] = run2; s[3] = run2; }
    else { s[1] = (run1 + 1) >> 1; s[2] = run2; s[3] = run3; s[4] = run3; }
    for (int k = 0; k < 5; ++k) sizes[k] = s[k];
}

// ---- pass 1: the five stream sizes of a chunk (thread 0 writes sizes[0..4]) ----------------------------------------
template <int CODEC, int NTH>
__device__ __forceinline__ void svb_sizes2(const int16_t* in, uint32_t n, uint32_t* sizes, Pack2Ws<NTH>* ws)
{
    constexpr int NW = NTH / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr bool TWO = (CODEC == RNSC_VBZ0 || CODEC == RNSC_C4 || CODEC == RNSC_C5);
    uint32_t na = 0, nb = 0, nc = 0;                                  // samples past each class threshold
    for (uint32_t tb = 0; tb < n; tb += NTH * P2_SPT) {
        const uint32_t i0 = tb + (uint32_t)tid * P2_SPT;
        const uint32_t cnt = i0 < n ? min((uint32_t)P2_SPT, n - i0) : 0u;
        if (cnt == 0) continue;
        uint32_t s[P2_SPT], prev;
        load16(in, n, i0, cnt, s, prev);
#pragma unroll
        for (int j = 0; j < P2_SPT; ++j) { count_one<CODEC>(delta_zigzag(s[j], prev), na, nb, nc); prev = s[j]; }   // past the end: zz = 0
    }
    uint32_t r1, r2, r3;                                              // per-thread unit counts; the n valid samples are added at the end
    if (TWO) { const uint32_t n1 = na - nb, n2 = nb - nc; if (CODEC == RNSC_VBZ0) { r1 = n1 + 2 * n2 + 4 * nc; r2 = r3 = 0; } else { r1 = n1; r2 = n2; r3 = nc; } }
    else { r1 = 0; r2 = nc; r3 = 0; }
    r1 = __reduce_add_sync(0xffffffffu, r1); r2 = __reduce_add_sync(0xffffffffu, r2); r3 = __reduce_add_sync(0xffffffffu, r3);
    if (lane == 0) { ws->red[0][warp] = r1; ws->red[1][warp] = r2; ws->red[2][warp] = r3; }
    __syncthreads();
    if (tid == 0) {
        uint32_t t1 = 0, t2 = 0, t3 = 0;
        for (int w = 0; w < NW; ++w) { t1 += ws->red[0][w]; t2 += ws->red[1][w]; t3 += ws->red[2][w]; }
        if (CODEC == RNSC_VBZ || CODEC == RNSC_C1) { t1 = n + t2; t2 = 0; }           // one byte per sample + one per large value
        else if (CODEC == RNSC_C2) t1 = n;
        else if (CODEC == RNSC_C3) t1 = n - t2;
        stream_sizes(CODEC, n, t1, t2, t3, sizes);
    }
    __syncthreads();
}

// ---- staging -> global ---------------------------------------------------------------------------------------------
// st[phase + i] <-> g[i] for i < count, where g - phase is 16-byte aligned (phase < 16): 128-bit stores for the full
// 16-byte groups; the ragged head and tail (< 16 bytes each) go out one byte per lane of the first warp (a byte loop on
// one thread would hold the whole warp's issue slots for up to 15 iterations per end and stream). Thread t of nt
// cooperating threads (nt a multiple of 32, t < nt).
__device__ __forceinline__ void flush_bytes(uint8_t* g, const uint8_t* st, uint32_t phase, uint32_t count, uint32_t t, uint32_t nt)
{
    uint8_t* g0 = g - phase;
    const uint32_t end = phase + count, body_lo = (phase + 15u) & ~15u, body_hi = end & ~15u;
    for (uint32_t lo = body_lo + t * 16u; lo < body_hi; lo += nt * 16u) *(V16*)(g0 + lo) = *(const V16*)(st + lo);
    if (t < 32) {
        const uint32_t l = t & 15u;
        const bool tail = t >= 16;
        const uint32_t i = (tail ? max(body_hi, body_lo) : phase) + l, lim = tail ? end : min(body_lo, end);
        if (i < lim) g0[i] = st[i];
    }
}
// nibble stream: st[2 * (phase + i)] | st[2 * (phase + i) + 1] << 4 <-> g[i] for i < count
__device__ __forceinline__ uint32_t pack_nibbles8(uint32_t a, uint32_t b)   // 8 staged nibbles (one per byte) -> 4 bytes
{
    const uint32_t ta = (a | (a >> 4)) & 0x00FF00FFu, tb = (b | (b >> 4)) & 0x00FF00FFu;
    return __byte_perm(ta, tb, 0x6420);
}
__device__ __forceinline__ void flush_nibbles(uint8_t* g, const uint8_t* st, uint32_t phase, uint32_t count, uint32_t t, uint32_t nt)
{
    uint8_t* g0 = g - phase;
    const uint32_t end = phase + count, body_lo = (phase + 15u) & ~15u, body_hi = end & ~15u;
    for (uint32_t lo = body_lo + t * 16u; lo < body_hi; lo += nt * 16u) {
        const V16 p = *(const V16*)(st + 2 * lo), q = *(const V16*)(st + 2 * lo + 16);
        *(V16*)(g0 + lo) = V16{pack_nibbles8(p.x, p.y), pack_nibbles8(p.z, p.w), pack_nibbles8(q.x, q.y), pack_nibbles8(q.z, q.w)};
    }
    if (t < 32) {
        const uint32_t l = t & 15u;
        const bool tail = t >= 16;
        const uint32_t i = (tail ? max(body_hi, body_lo) : phase) + l, lim = tail ? end : min(body_lo, end);
        if (i < lim) g0[i] = (uint8_t)(st[2 * i] | (st[2 * i + 1] << 4));
    }
}

// C4/C5: one sample's unit into the staging areas without a branch. t = the sample's two key bits in place, one = the
// value of code 1 at that place; q1/q2/q3/q4 = running shared-memory addresses in the nibble, byte, low and high areas.
__device__ __forceinline__ void scatter_c45(uint32_t t, uint32_t one, uint32_t v, saddr_t& q1, saddr_t& q2, saddr_t& q3, saddr_t& q4)
{
#ifdef RNSC_EMUL
    if (t == one) { *(uint8_t*)q1 = (uint8_t)v; ++q1; }
    else if (t == 2 * one) { *(uint8_t*)q2 = (uint8_t)v; ++q2; }
    else if (t == 3 * one) { *(uint8_t*)q3 = (uint8_t)v; *(uint8_t*)q4 = (uint8_t)(v >> 8); ++q3; ++q4; }
#else
    asm volatile("{\n.reg .pred c1, c2, c3, c12;\n.reg .b32 a, h;\n"
                 "setp.eq.u32 c1, %4, %5;\nsetp.eq.u32 c2, %4, %6;\nsetp.eq.u32 c3, %4, %7;\nor.pred c12, c1, c2;\n"
                 "selp.b32 a, %0, %1, c1;\n@c12 st.shared.u8 [a], %8;\n@c1 add.u32 %0, %0, 1;\n@c2 add.u32 %1, %1, 1;\n"
                 "@c3 shr.u32 h, %8, 8;\n@c3 st.shared.u8 [%2], %8;\n@c3 st.shared.u8 [%3], h;\n@c3 add.u32 %2, %2, 1;\n@c3 add.u32 %3, %3, 1;\n}"
                 : "+r"(q1), "+r"(q2), "+r"(q3), "+r"(q4) : "r"(t), "r"(one), "r"(2u * one), "r"(3u * one), "r"(v) : "memory");
#endif
}

// ---- pass 2: write the streams of a chunk at dst[0..4] (global; stream k gets exactly the size pass 1 reported) ------
// `stage` = P2_STAGE_BYTES of shared memory, 16-byte aligned.
template <int CODEC, int NTH>
__device__ __forceinline__ void svb_pack2(const int16_t* in, uint32_t n, uint8_t* const* dst, Pack2Ws<NTH>* ws, uint8_t* stage, uint32_t* runs_out = nullptr)
{
    constexpr bool TWO = (CODEC == RNSC_VBZ0 || CODEC == RNSC_C4 || CODEC == RNSC_C5);
    constexpr bool NIB1 = TWO;                                        // stream 1 is a nibble stream
    constexpr bool WORD_B = (CODEC == RNSC_C4 || CODEC == RNSC_C5);   // a third unit count to scan
    constexpr int NW = NTH / 32;
    constexpr int T = NTH * P2_SPT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* const d0 = dst[0]; uint8_t* const d1 = dst[1]; uint8_t* const d2 = dst[2]; uint8_t* const d3 = dst[3]; uint8_t* const d4 = dst[4];
    // staging areas (each with room for a 16-byte phase; the nibble area for a 32-byte one)
    uint8_t* const st1 = stage;
    uint8_t* const st2 = stage + ((CODEC == RNSC_VBZ || CODEC == RNSC_C1) ? 2 * T + 32 : T + 48);
    uint8_t* const st3 = st2 + T + 16;
    uint8_t* const st4 = st3 + T + 16;
    uint32_t run1 = 0, run2 = 0, run3 = 0;                            // units written so far (stream 1: nibbles when NIB1)
    int flip = 0;
    for (uint32_t tb = 0; tb < n; tb += T, flip ^= 1) {
        const uint32_t i0 = tb + (uint32_t)tid * P2_SPT;
        const uint32_t cnt = i0 < n ? min((uint32_t)P2_SPT, n - i0) : 0u;
        uint32_t v[P2_SPT], kb = 0, c1 = 0, c2 = 0, c3 = 0;
        if (cnt) {
            uint32_t s[P2_SPT], prev;
            load16(in, n, i0, cnt, s, prev);
            if (i0 + T < n) prefetch_l1(in + i0 + T);                 // the thread's samples of the next tile: L1 hit after the two barriers
            kb = classify16<CODEC>(s, prev, v);
            unit_counts<CODEC>(kb, cnt, c1, c2, c3);
        }
        // ---- one barrier: scan of the unit counts; it also separates the previous tile's flush from this tile's scatter
        const uint32_t wa = c1 | (c2 << 16);
        const uint32_t ia = warp_inclusive_scan(wa), ib = WORD_B ? warp_inclusive_scan(c3) : 0u;
        if (lane == 31) { ws->sa[flip][warp] = ia; if (WORD_B) ws->sb[flip][warp] = ib; }
        __syncthreads();
        const uint32_t ta = lane < NW ? ws->sa[flip][lane] : 0u, tb3 = (WORD_B && lane < NW) ? ws->sb[flip][lane] : 0u;
        const uint32_t totA = __reduce_add_sync(0xffffffffu, ta), exA = __reduce_add_sync(0xffffffffu, lane < warp ? ta : 0u) + ia - wa;
        const uint32_t t3 = WORD_B ? __reduce_add_sync(0xffffffffu, tb3) : 0u;
        const uint32_t ex3 = WORD_B ? __reduce_add_sync(0xffffffffu, lane < warp ? tb3 : 0u) + ib - c3 : 0u;
        const uint32_t t1 = totA & 0xFFFFu, t2 = totA >> 16;
        // ---- phases of the destinations of this tile ----
        const uint32_t par = NIB1 ? (run1 & 1u) : 0u;                 // a pending low nibble from the previous tile
        uint8_t* const g1 = d1 + (NIB1 ? (run1 >> 1) : run1);
        const uint32_t ph1 = (uint32_t)((uintptr_t)g1 & 15), ph2 = (uint32_t)((uintptr_t)(d2 + run2) & 15);
        const uint32_t ph3 = (uint32_t)((uintptr_t)(d3 + ((CODEC == RNSC_C3) ? run2 : run3)) & 15), ph4 = (uint32_t)((uintptr_t)(d4 + run3) & 15);
        uint32_t p1 = (NIB1 ? 2 * ph1 + par : ph1) + (exA & 0xFFFFu), p2 = ph2 + (exA >> 16), p3 = ex3;
        if (NIB1) {
            if (tid == 0) {
                const uint32_t pend = ws->carry[flip ^ 1];
                if (par) st1[2 * ph1] = (uint8_t)pend;
                if (t1 == 0) ws->carry[flip] = pend;
            }
            if (c1 && (exA & 0xFFFFu) + c1 == t1) {                   // owner of the tile's last nibble
                uint32_t lastv = 0;
                if (CODEC == RNSC_VBZ0) {
#pragma unroll
                    for (int j = 0; j < P2_SPT; ++j) { const uint32_t code = (kb >> (2 * j)) & 3u; if (code) lastv = (v[j] >> (code == 3 ? 12 : (code == 2 ? 4 : 0))) & 15u; }
                } else {
#pragma unroll
                    for (int j = 0; j < P2_SPT; ++j) if (((kb >> (2 * j)) & 3u) == 1u) lastv = v[j] & 15u;
                }
                ws->carry[flip] = lastv;
            }
        }
        // ---- keys straight to global: one aligned word / half word per full thread, bytes otherwise ----
        if (cnt) {
            if (TWO) {
                uint8_t* kp = d0 + (i0 >> 2);
                if (cnt == (uint32_t)P2_SPT) {
                    const uint32_t al = (uint32_t)((uintptr_t)kp & 3);               // same for every thread of the chunk
                    if (al == 0) *(uint32_t*)kp = kb;
                    else if (al == 2) { *(uint16_t*)kp = (uint16_t)kb; *(uint16_t*)(kp + 2) = (uint16_t)(kb >> 16); }
                    else { kp[0] = (uint8_t)kb; kp[1] = (uint8_t)(kb >> 8); kp[2] = (uint8_t)(kb >> 16); kp[3] = (uint8_t)(kb >> 24); }
                } else for (uint32_t q = 0; q < (cnt + 3) >> 2; ++q) kp[q] = (uint8_t)(kb >> (8 * q));
            } else {
                uint8_t* kp = d0 + (i0 >> 3);
                if (cnt == (uint32_t)P2_SPT) {
                    if (((uintptr_t)kp & 1) == 0) *(uint16_t*)kp = (uint16_t)kb;
                    else { kp[0] = (uint8_t)kb; kp[1] = (uint8_t)(kb >> 8); }
                } else for (uint32_t q = 0; q < (cnt + 7) >> 3; ++q) kp[q] = (uint8_t)(kb >> (8 * q));
            }
        }
        // ---- scatter the units into the staging areas ----
        saddr_t q1 = to_saddr(st1) + p1, q2 = to_saddr(st2) + p2, q3 = to_saddr(st3) + ph3 + p3, q4 = to_saddr(st4) + ph4 + p3;   // C4/C5
#pragma unroll
        for (int j = 0; j < P2_SPT; ++j) {
            const uint32_t vv = v[j];
            if (CODEC == RNSC_VBZ || CODEC == RNSC_C1) {
                if ((uint32_t)j < cnt) { st1[p1++] = (uint8_t)vv; if ((kb >> j) & 1u) st1[p1++] = (uint8_t)(vv >> 8); }
            } else if (CODEC == RNSC_C2) {
                if ((uint32_t)j < cnt) { st1[p1++] = (uint8_t)vv; if ((kb >> j) & 1u) st2[p2++] = (uint8_t)(vv >> 8); }
            } else if (CODEC == RNSC_C3) {
                if ((uint32_t)j < cnt) { if ((kb >> j) & 1u) { st2[p2] = (uint8_t)vv; st3[ph3 - ph2 + p2] = (uint8_t)(vv >> 8); p2++; } else st1[p1++] = (uint8_t)vv; }
            } else {
                const uint32_t code = (kb >> (2 * j)) & 3u;           // samples past the end have code 0
                if (CODEC == RNSC_VBZ0) {                              // 0 / 1 / 2 / 4 nibbles, straight-line
                    if (code >= 1u) st1[p1] = (uint8_t)(vv & 15u);
                    if (code >= 2u) st1[p1 + 1] = (uint8_t)((vv >> 4) & 15u);
                    if (code == 3u) { st1[p1 + 2] = (uint8_t)((vv >> 8) & 15u); st1[p1 + 3] = (uint8_t)(vv >> 12); }
                    p1 += nib_units(code);
                }
                else scatter_c45(kb & (3u << (2 * j)), 1u << (2 * j), vv, q1, q2, q3, q4);
            }
        }
        __syncthreads();
        // ---- flush, 16-byte vectors in the aligned body ----
        if (WORD_B && NW >= 4) {
            // C4/C5: the streams are flushed by different warps (nibbles: warp 0; large values: warp 1; bytes: the rest),
            // so a thread runs one flush's set-up instead of four
            if (warp == 0) flush_nibbles(g1, st1, ph1, (par + t1) >> 1, (uint32_t)lane, 32u);
            else if (warp == 1) { if (t3) { flush_bytes(d3 + run3, st3, ph3, t3, (uint32_t)lane, 32u); flush_bytes(d4 + run3, st4, ph4, t3, (uint32_t)lane, 32u); } }
            else flush_bytes(d2 + run2, st2, ph2, t2, (uint32_t)tid - 64u, (uint32_t)NTH - 64u);
        } else {
            if (NIB1) flush_nibbles(g1, st1, ph1, (par + t1) >> 1, (uint32_t)tid, (uint32_t)NTH);
            else flush_bytes(g1, st1, ph1, t1, (uint32_t)tid, (uint32_t)NTH);
            if (CODEC == RNSC_C2 || CODEC == RNSC_C4 || CODEC == RNSC_C5) flush_bytes(d2 + run2, st2, ph2, t2, (uint32_t)tid, (uint32_t)NTH);
            if (CODEC == RNSC_C3) { flush_bytes(d2 + run2, st2, ph2, t2, (uint32_t)tid, (uint32_t)NTH); flush_bytes(d3 + run2, st3, ph3, t2, (uint32_t)tid, (uint32_t)NTH); }
            if (WORD_B && t3) { flush_bytes(d3 + run3, st3, ph3, t3, (uint32_t)tid, (uint32_t)NTH); flush_bytes(d4 + run3, st4, ph4, t3, (uint32_t)tid, (uint32_t)NTH); }
        }
        run1 += t1; run2 += t2; run3 += t3;
    }
    __syncthreads();
    if (NIB1 && tid == 0 && (run1 & 1u)) d1[run1 >> 1] = (uint8_t)ws->carry[flip ^ 1];   // trailing half byte
    if (runs_out && tid == 0) { runs_out[0] = run1; runs_out[1] = run2; runs_out[2] = run3; }        // units written per stream group (for callers that did not run the size pass)
    __syncthreads();
}

}  // namespace rnsc
