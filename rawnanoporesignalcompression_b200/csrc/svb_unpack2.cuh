// svb_unpack2.cuh -- splitter inverse (streams -> int16 samples), second generation.
//
// Same arithmetic as svb_unpack (svb_codec.cuh; reference decode_scalar_* loops cited there), organised for
// instruction count:
//   * a thread owns U2_SPT = 16 consecutive samples chosen so that they are two 16-byte aligned segments of the
//     OUTPUT: results go from registers to global memory with two 128-bit stores, no shared-memory tile;
//   * key bits and the thread's payload bytes/nibbles are fetched as aligned 32-bit words and funnel-shifted
//     into register windows (C4/C5), so the per-sample work is predicates, shifts and adds, no loads;
//   * stream positions and the delta^-1 prefix sum are two 32-bit CTA scans behind ONE barrier per tile: the unit
//     counts of two streams of the NEXT tile share a word, and the count of large values of the next tile rides in
//     the low half of the word whose high half carries this tile's delta sums (mod 2^16);
//   * C4/C5 go four samples at a time: a 256-entry table indexed by the key byte gives the byte-permute selectors that
//     pull the samples' units out of the stream windows; zigzag^-1 runs on two 16-bit lanes per register.
// Works on generic pointers; every stream must be readable up to 8 bytes past its end (the fast decoder's
// region pads each stream with 16 bytes).
#pragma once
#include "block_prims.cuh"
#include "plan.cuh"

namespace rnsc {

constexpr int U2_SPT = 16;
// (Tried: 32 samples per thread and tile, decoded as two 16-sample halves, to pay a tile's scans and barrier once per 32
// samples -- slower: the halves are one dependent chain per thread and the last tile gets coarser; 20k-read C5 decode 508 ->
// 491 GB/s. Tried: warps whose samples all have key 0 or lie past the end skip the arithmetic -- slower too, 508 -> 501.)

// C4/C5: what one key byte (four 2-bit codes) says about its four samples. x = byte-permute selectors that pull the
// samples' units out of the next four nibbles (low half) / bytes (high half) of the stream windows, a zero byte for
// samples of another class; y = bits to drop from the nibble window | large values << 5 | bits to drop from the byte
// window << 8 | selectors of the samples' large values (bytes of the thread's four-byte Llow / Lhigh windows, counted from
// the group's first one) << 16; z, w = the class offsets (C5: 1 / 17 / 273) of samples 0,1 and 2,3 as 16-bit lanes.
struct alignas(16) KeyLut { uint32_t x, y, z, w; };
// The table is read with one 128-bit shared load per four samples, eight lanes per wavefront, and an entry covers four of
// the 32 banks: lanes whose key bytes differ collide unless their entries lie in different bank groups (measured: 2.7
// wavefronts per quarter warp with one copy, a third of all shared-memory traffic of the decode). Where a chunk leaves the
// shared memory for it the table is therefore REPLICATED behind the chunk's streams (the planner decides, PlanHead::lut_off):
// lane l reads copy l mod R, entry i of copy r sits in slot i * R + r, so that the lanes of a quarter warp share a bank group
// only with the lanes of their own copy. A chunk without the room (the largest, noisiest ones) reads the single copy in
// front of the region.
static_assert(sizeof(KeyLut) * 256 == KEY_LUT_BYTES, "key-byte table size (plan.cuh)");

template <int CODEC>
__device__ inline KeyLut make_key_lut(uint32_t kbyte)
{
    uint32_t selS = 0, selM = 0, selL = 0, nS = 0, nM = 0, nL = 0, off[4];
    for (int j = 0; j < 4; ++j) {
        const uint32_t c = (kbyte >> (2 * j)) & 3u;
        selS |= (c == 1 ? nS : 4u) << (4 * j);
        selM |= (c == 2 ? nM : 4u) << (4 * j);
        selL |= (c == 3 ? nL : 4u) << (4 * j);
        off[j] = CODEC == RNSC_C5 ? (c == 1 ? 1u : (c == 2 ? 17u : (c == 3 ? 273u : 0u))) : 0u;
        nS += c == 1; nM += c == 2; nL += c == 3;
    }
    return KeyLut{selS | (selM << 16), (4 * nS) | (nL << 5) | ((8 * nM) << 8) | (selL << 16), off[0] | (off[1] << 16), off[2] | (off[3] << 16)};
}

template <int NTH>
struct Unpack2Ws {
    uint32_t sa[2][NTH / 32], sb[2][NTH / 32];   // warp totals of the two scans of a tile (A: unit counts, B: delta sums | large count), double-buffered
    int err;
};

// Shared-memory byte addresses as plain integers: the streams live in the caller's shared-memory region, and a load
// through a generic pointer would take the slower generic path.
#ifdef RNSC_EMUL
typedef uintptr_t saddr_t;
__device__ __forceinline__ saddr_t to_saddr(const void* p) { return (uintptr_t)p; }
__device__ __forceinline__ uint32_t ldw(saddr_t a) { return *(const uint32_t*)a; }
#else
typedef uint32_t saddr_t;
__device__ __forceinline__ saddr_t to_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t ldw(saddr_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
#endif

// 32 bits starting at bit `bit` of the byte stream at shared address p (any alignment). Reads the two aligned words
// that cover them: up to 7 bytes past the last byte needed, never below the aligned word holding the first.
__device__ __forceinline__ uint32_t load_bits32(saddr_t p, uint32_t bit)
{
    const saddr_t a = p + (bit >> 3), al = a & ~(saddr_t)3;
    const uint32_t sh = ((uint32_t)(a & 3) << 3) + (bit & 7);
    return __funnelshift_r(ldw(al), ldw(al + 4), sh);
}
__device__ __forceinline__ void load_bits64(saddr_t p, uint32_t bit, uint32_t& lo, uint32_t& hi)
{
    const saddr_t a = p + (bit >> 3), al = a & ~(saddr_t)3;
    const uint32_t sh = ((uint32_t)(a & 3) << 3) + (bit & 7);
    const uint32_t w0 = ldw(al), w1 = ldw(al + 4), w2 = ldw(al + 8);
    lo = __funnelshift_r(w0, w1, sh); hi = __funnelshift_r(w1, w2, sh);
}

// two 16-bit lanes added without a carry between them (VIADD.16x2)
__device__ __forceinline__ uint32_t add16x2(uint32_t a, uint32_t b)
{
#ifdef RNSC_EMUL
    return ((a + b) & 0xFFFFu) | (((a >> 16) + (b >> 16)) << 16);
#else
    uint32_t r; asm("add.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#endif
}

// sum += zigzag^-1(z) in 32-bit arithmetic (the low 16 bits are what counts): z even -> z/2, z odd -> -(z/2) - 1 = z/2 - z
__device__ __forceinline__ uint32_t zz_dec(uint32_t z) { return (z >> 1) - (z & 1u) * z; }

// Stream k = base[v[k].off .. + v[k].len): v[0] = keys, v[1..4] = payload streams in wire order. `base` is the
// caller's shared-memory region (passed as base + offsets and inlined so that the loads compile to shared-space loads,
// not generic ones). Returns an rnsc_status (same in every thread).
template <int CODEC, int NTH>
__device__ __forceinline__ int svb_unpack2(const uint8_t* base, const KeyLut* lut, uint32_t lut_copies, const ViewPlan* v, uint32_t n, int16_t* out, Unpack2Ws<NTH>* ws)
{
    constexpr bool TWO = (CODEC == RNSC_VBZ0 || CODEC == RNSC_C4 || CODEC == RNSC_C5);
    constexpr bool WIN = (CODEC == RNSC_C4 || CODEC == RNSC_C5);       // register-window fast path
    constexpr bool ONE_STREAM = (CODEC == RNSC_VBZ || CODEC == RNSC_C1 || CODEC == RNSC_VBZ0);   // scan word A holds c1 alone
    constexpr int KB = TWO ? 2 : 1;
    constexpr int S = U2_SPT;
    const int tid = threadIdx.x;
    if (tid == 0) ws->err = 0;
    const uint8_t *s1 = base + v[1].off, *s2 = base + v[2].off, *s3 = base + v[3].off, *s4 = base + v[4].off;
    const saddr_t a0 = to_saddr(base + v[0].off), a1 = to_saddr(s1), a2 = to_saddr(s2), a3 = to_saddr(s3), a4 = to_saddr(s4);
    const saddr_t vs = to_saddr(v);                                    // the views: {off, len} x 5
    const uint32_t LR = lut_copies;                                    // copies of the key-byte table (a power of two)
    const KeyLut* const lutr = lut + ((uint32_t)tid & (LR - 1u));      // this lane's copy
    const uint32_t ph = (uint32_t)(((uintptr_t)out >> 1) & 7);        // samples past a 16-byte boundary
    int16_t* outa = out - ph;                                          // 16-byte aligned; virtual sample vi <-> outa[vi]
    const uint32_t nv = n + ph;
    uint32_t run1 = 0, run2 = 0, run3 = 0, carry = 0;
    bool oob = false;
    // key bits of the valid samples of this thread in the tile starting at virtual sample tb, at slot positions jlo..
    auto load_keys = [&](uint32_t tb) -> uint32_t {
        const uint32_t v0 = tb + (uint32_t)tid * S;
        const uint32_t vlo = max(v0, ph), vhi = min(v0 + S, nv);
        if (vhi <= vlo) return 0u;
        const uint32_t cnt = vhi - vlo;
        uint32_t kb = load_bits32(a0, KB * (vlo - ph));
        if (KB * cnt < 32) kb &= (1u << (KB * cnt)) - 1u;
        return kb << (KB * (vlo - v0));
    };
    // valid samples of this thread in the tile starting at tb
    auto tile_cnt = [&](uint32_t tb) -> uint32_t {
        const uint32_t v0 = tb + (uint32_t)tid * S;
        const uint32_t vlo = max(v0, ph), vhi = min(v0 + S, nv);
        return vhi > vlo ? vhi - vlo : 0u;
    };
    // class masks and unit counts of a tile from its key bits; wordA = what the positions scan adds up (see below)
    auto classify = [&](uint32_t kb, uint32_t cnt, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& m1, uint32_t& m2, uint32_t& m3) -> uint32_t {
        c1 = c2 = c3 = m1 = m2 = m3 = 0;
        if (TWO) {
            const uint32_t lo = kb & 0x55555555u, hi = (kb >> 1) & 0x55555555u;
            m1 = lo & ~hi; m2 = hi & ~lo; m3 = lo & hi;
            const uint32_t n1 = __popc(m1), n2 = __popc(m2), n3 = __popc(m3);
            if (CODEC == RNSC_VBZ0) c1 = n1 + 2 * n2 + 4 * n3; else { c1 = n1; c2 = n2; c3 = n3; }
        } else {
            const uint32_t nl = __popc(kb);
            if (CODEC == RNSC_VBZ || CODEC == RNSC_C1) c1 = cnt + nl;
            else if (CODEC == RNSC_C2) { c1 = cnt; c2 = nl; }
            else { c2 = nl; c1 = cnt - nl; }
        }
        // one word: a thread adds at most 64 units (VBZ0, alone in the word) or 16 + 16 (two halves), a tile at most 2^16 / 2^14
        return ONE_STREAM ? c1 : (c1 | (c2 << 16));
    };
    // ONE barrier per tile: after a tile is decoded, word B = (the thread's delta sum << 16) | large values of the NEXT
    // tile and word A = the unit counts of the NEXT tile are scanned together. The warp totals alternate between two
    // buffers (a warp may be a tile ahead of the slowest reader of the previous totals, never two).
    auto scan2 = [&](uint32_t a, uint32_t b, int flip, uint32_t& exA, uint32_t& totA, uint32_t& exB, uint32_t& totB) {
        constexpr int NW = NTH / 32;
        const int lane = tid & 31, warp = tid >> 5;
        const uint32_t ia = warp_inclusive_scan(a), ib = warp_inclusive_scan(b);
        if (lane == 31) { ws->sa[flip][warp] = ia; ws->sb[flip][warp] = ib; }
        __syncthreads();
        const uint32_t wa = lane < NW ? ws->sa[flip][lane] : 0u, wb = lane < NW ? ws->sb[flip][lane] : 0u;
        totA = __reduce_add_sync(0xffffffffu, wa); totB = __reduce_add_sync(0xffffffffu, wb);
        exA = __reduce_add_sync(0xffffffffu, lane < warp ? wa : 0u) + ia - a;
        exB = __reduce_add_sync(0xffffffffu, lane < warp ? wb : 0u) + ib - b;
    };
    __syncthreads();
    uint32_t kb_next = nv ? load_keys(0) : 0u, p3 = 0;
    uint32_t c1, c2, c3, m1, m2, m3, exA, totA;
    int flip = 0;
    {   // positions of the first tile
        uint32_t exB, totB;
        const uint32_t wordA = classify(kb_next, tile_cnt(0), c1, c2, c3, m1, m2, m3);
        scan2(wordA, c3, flip, exA, totA, exB, totB);
        p3 = exB; run3 = totB;
    }
    for (uint32_t tb = 0; tb < nv; tb += NTH * S) {
        flip ^= 1;
        const uint32_t v0 = tb + (uint32_t)tid * S;
        const uint32_t vlo = max(v0, ph), vhi = min(v0 + S, nv);
        const uint32_t cnt = vhi > vlo ? vhi - vlo : 0, jlo = cnt ? vlo - v0 : 0;
        uint32_t kb = kb_next;
        const bool more = tb + NTH * S < nv;
        kb_next = more ? load_keys(tb + NTH * S) : 0u;
        classify(kb, cnt, c1, c2, c3, m1, m2, m3);
        // ---- stream positions from the scan made at the end of the previous tile ----
        uint32_t p1, p2;
        if (ONE_STREAM) { p1 = run1 + exA; p2 = 0; run1 += totA; }
        else { p1 = run1 + (exA & 0xFFFFu); p2 = run2 + (exA >> 16); run1 += totA & 0xFFFFu; run2 += totA >> 16; }
        // ---- one bounds check per TILE, the same in every thread: run1..3 now count the units up to the end of this tile,
        //      and a thread's reads end at or before them (over-reads stay inside the padded buffers). The stream lengths
        //      are re-read from shared memory here instead of living in registers across the tile ----
        bool ok;
        {
            const uint32_t l1 = ldw(vs + 12), l2 = ldw(vs + 20), l3 = ldw(vs + 28), l4 = ldw(vs + 36);
            if (CODEC == RNSC_VBZ || CODEC == RNSC_C1) ok = run1 <= l1;
            else if (CODEC == RNSC_C2) ok = run1 <= l1 && run2 <= l2;
            else if (CODEC == RNSC_C3) ok = run1 <= l1 && run2 <= l2 && run2 <= l3;
            else if (CODEC == RNSC_VBZ0) ok = run1 <= 2 * l1;
            else ok = run1 <= 2 * l1 && run2 <= l2 && run3 <= l3 && run3 <= l4;
        }
        if (!ok) { oob = true; kb = 0; m1 = m2 = m3 = 0; c3 = 0; }
        // ---- gather, zigzag^-1, thread-local running sum ----
        uint32_t val[S]; uint32_t acc = 0;
        if (WIN) {
            constexpr uint32_t OL = CODEC == RNSC_C5 ? 273u : 0u;
            uint32_t sw[2] = {0, 0}, mlo[2] = {0, 0}, mhi[2] = {0, 0};
            if (cnt && ok) {
                const uint32_t n1a = __popc(m1 & 0xFFFFu), n2a = __popc(m2 & 0xFFFFu);
                sw[0] = load_bits32(a1, 4 * p1); sw[1] = load_bits32(a1, 4 * (p1 + n1a));
                load_bits64(a2, 8 * p2, mlo[0], mhi[0]); load_bits64(a2, 8 * (p2 + n2a), mlo[1], mhi[1]);
            }
            // the thread's (up to three) large values: bytes p3.. of the low and high byte streams, pulled into their samples by
            // the key byte's selectors like the other classes -- no branch, no patching. A high byte of 0xFE or more could
            // wrap the uint16 sum with the offset 273 (C5.hpp:224-228) into the neighbouring lane: such a thread (corrupt
            // input only) takes the byte-wise path below.
            uint32_t lo4 = 0, hi4 = 0;
            if (m3) { lo4 = load_bits32(a3, 8 * p3); hi4 = load_bits32(a4, 8 * p3); }
            const uint32_t nhi = ~(hi4 | 0x01010101u);
            const bool wrap = CODEC == RNSC_C5 && ((nhi - 0x01010101u) & ~nhi & 0x80808080u) != 0;
            if (c3 <= 3 && !wrap) {
                // four samples at a time: the key byte's table entry pulls their units out of the windows with byte
                // permutes; zigzag^-1 runs on two 16-bit lanes per register
                uint32_t zp[S / 2], r3x = 0;                                       // r3x: large values consumed so far x 0x1111
#pragma unroll
                for (int g = 0; g < S / 4; ++g) {
                    const int h = g >> 1;
                    const KeyLut e = lutr[__byte_perm(kb, 0u, 0x4440u | (uint32_t)g) * LR];
                    uint32_t x = sw[h] & 0xFFFFu;                                  // next four nibbles -> four bytes
                    x = (x | (x << 8)) & 0x00FF00FFu; x = (x | (x << 4)) & 0x0F0F0F0Fu;
                    const uint32_t sel3 = (e.y >> 16) + r3x;                       // selectors stay below 8 (no large value: 4 + at most 3 consumed)
                    r3x += ((e.y >> 5) & 7u) * 0x1111u;
                    const uint32_t zb = __byte_perm(x, 0u, e.x) | __byte_perm(mlo[h], 0u, e.x >> 16) | __byte_perm(lo4, 0u, sel3);
                    const uint32_t zh = __byte_perm(hi4, 0u, sel3);
                    sw[h] = __funnelshift_r(sw[h], 0u, e.y);                       // low 5 bits of e.y: 4 x nibbles used (<= 16)
                    const uint32_t shm = (e.y >> 8) & 63u;                         // 8 x bytes used (<= 32)
                    mlo[h] = __funnelshift_rc(mlo[h], mhi[h], shm); mhi[h] = __funnelshift_rc(mhi[h], 0u, shm);
                    zp[2 * g] = __byte_perm(zb, zh, 0x5140u) + e.z;
                    zp[2 * g + 1] = __byte_perm(zb, zh, 0x7362u) + e.w;
                }
#pragma unroll
                for (int q = 0; q < S / 2; ++q) {
                    const uint32_t z = zp[q];
                    const uint32_t d = ((z >> 1) & 0x7FFF7FFFu) ^ ((z & 0x00010001u) * 0xFFFFu);   // zigzag^-1 of both lanes
                    acc += d; val[2 * q] = acc;                                    // bits above 16 are don't-care from here on
                    acc += d >> 16; val[2 * q + 1] = acc;
                }
            } else {   // a burst of large values: same arithmetic, byte loads
                uint32_t k3 = p3;
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    const int h = j >> 3;
                    uint32_t z = 0;
                    if ((m1 >> (2 * j)) & 1u) { z = (sw[h] & 15u) + (CODEC == RNSC_C5 ? 1u : 0u); sw[h] >>= 4; }
                    if ((m2 >> (2 * j)) & 1u) { z = (mlo[h] & 255u) + (CODEC == RNSC_C5 ? 17u : 0u); mlo[h] = __funnelshift_r(mlo[h], mhi[h], 8); mhi[h] >>= 8; }
                    if ((m3 >> (2 * j)) & 1u) { z = ((ld8(s4 + k3) << 8) + ld8(s3 + k3) + OL) & 0xFFFFu; ++k3; }
                    acc += zz_dec(z);
                    val[j] = acc;
                }
            }
        } else if ((CODEC == RNSC_VBZ || CODEC == RNSC_C1) && cnt == S) {
            // full thread: the (16 .. 32) payload bytes go through four 64-bit register windows, one per 4 samples
            uint32_t used = 0;
#pragma unroll
            for (int g = 0; g < S / 4; ++g) {
                uint32_t wlo, whi;
                load_bits64(a1, 8 * (p1 + used), wlo, whi);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = 4 * g + jj;
                    const bool k = (kb >> j) & 1u;
                    const uint32_t z = k ? (wlo & 0xFFFFu) : (wlo & 0xFFu);
                    const uint32_t sh = k ? 16u : 8u;
                    wlo = __funnelshift_r(wlo, whi, sh); whi >>= sh;
                    acc += zz_dec(z);
                    val[j] = acc;
                }
                used += 4 + __popc((kb >> (4 * g)) & 15u);
            }
        } else {
#pragma unroll
            for (int j = 0; j < S; ++j) {
                uint32_t z = 0;
                if (CODEC == RNSC_VBZ || CODEC == RNSC_C1) {
                    const uint32_t k = (kb >> j) & 1u;
                    if ((uint32_t)j >= jlo && (uint32_t)j < jlo + cnt) { z = ld8(s1 + p1); if (k) z |= ld8(s1 + p1 + 1) << 8; p1 += 1 + k; }
                } else if (CODEC == RNSC_C2) {
                    const uint32_t k = (kb >> j) & 1u;
                    if ((uint32_t)j >= jlo && (uint32_t)j < jlo + cnt) { z = ld8(s1 + p1); ++p1; if (k) { z |= ld8(s2 + p2) << 8; ++p2; } }
                } else if (CODEC == RNSC_C3) {
                    const uint32_t k = (kb >> j) & 1u;
                    if ((uint32_t)j >= jlo && (uint32_t)j < jlo + cnt) {
                        if (k) { z = ld8(s2 + p2) | (ld8(s3 + p2) << 8); ++p2; } else { z = ld8(s1 + p1); ++p1; }
                    }
                } else {   // VBZ0: 0/1/2/4 nibbles of one stream
                    const uint32_t code = (kb >> (2 * j)) & 3u, nn = code == 3 ? 4u : code;
                    uint32_t x = 0;
                    for (uint32_t q = 0; q < nn; ++q) x |= ((ld8(s1 + ((p1 + q) >> 1)) >> (4 * ((p1 + q) & 1))) & 15u) << (4 * q);
                    p1 += nn;
                    z = code == 0 ? 0u : (x + (code == 1 ? 1u : (code == 2 ? 17u : 273u))) & 0xFFFFu;
                }
                acc += zz_dec(z);
                val[j] = acc;
            }
        }
        // ---- delta^-1: CTA scan of the per-thread sums (mod 2^16, upper half of word B); its lower half and word A scan
        //      the large-value and unit counts of the next tile ----
        // the running sums leave as 16-bit lanes, packed in pairs BEFORE the barrier (half the registers live across it)
        uint32_t w[S / 2];
#pragma unroll
        for (int q = 0; q < S / 2; ++q) w[q] = __byte_perm(val[2 * q], val[2 * q + 1], 0x5410);
        uint32_t exB, totB;
        {
            uint32_t n1, n2, n3, q1, q2, q3;
            const uint32_t wordA = classify(kb_next, more ? tile_cnt(tb + NTH * S) : 0u, n1, n2, n3, q1, q2, q3);
            scan2(wordA, (acc << 16) | n3, flip, exA, totA, exB, totB);
        }
        const uint32_t base = carry + (exB >> 16);
        carry += totB >> 16;
        if (WIN) { p3 = run3 + (exB & 0xFFFFu); run3 += totB & 0xFFFFu; }
        const uint32_t base2 = __byte_perm(base, 0u, 0x1010);            // the carry-in in both lanes
#pragma unroll
        for (int q = 0; q < S / 2; ++q) w[q] = add16x2(w[q], base2);
        if (cnt == S) {
            V16* dst = (V16*)(outa + v0);
            dst[0] = V16{w[0], w[1], w[2], w[3]};
            dst[1] = V16{w[4], w[5], w[6], w[7]};
        } else if (cnt) {
#pragma unroll
            for (int j = 0; j < S; ++j) if ((uint32_t)j >= jlo && (uint32_t)j < jlo + cnt) outa[v0 + j] = (int16_t)(uint16_t)(w[j >> 1] >> (16 * (j & 1)));
        }
    }
    // ---- the reference's "Remaining data at end of signal buffer" test (A.8) + out-of-bounds guard ----
    bool rem;
    const uint32_t l1 = v[1].len, l2 = v[2].len, l3 = v[3].len, l4 = v[4].len;
    if (CODEC == RNSC_VBZ || CODEC == RNSC_C1) rem = run1 != l1;
    else if (CODEC == RNSC_VBZ0) rem = ((run1 + 1) >> 1) != l1;
    else if (CODEC == RNSC_C2) rem = run2 != l2;
    else if (CODEC == RNSC_C3) rem = run2 != l3;
    else rem = run3 != l4;
    if (oob || (tid == 0 && rem)) ws->err = RNSC_ERR_REMAINING;
    __syncthreads();
    return ws->err;
}

// Shared memory the key-byte table of a codec needs in front of the region (only C4 and C5 have one).
__host__ __device__ inline uint32_t unpack2_lut_bytes(int codec) { return (codec == RNSC_C4 || codec == RNSC_C5) ? 256u * (uint32_t)sizeof(KeyLut) : 0u; }

// Once per CTA, before the first svb_unpack2; followed by a barrier at the caller. `lut` = unpack2_lut_bytes(codec) bytes.
template <int NTH>
__device__ inline void unpack2_init(int codec, KeyLut* lut)
{
    if (codec != RNSC_C4 && codec != RNSC_C5) return;
    for (uint32_t i = threadIdx.x; i < 256; i += NTH) lut[i] = codec == RNSC_C5 ? make_key_lut<RNSC_C5>(i) : make_key_lut<RNSC_C4>(i);
}

template <int NTH>
__device__ __forceinline__ int unpack2_any(int codec, const uint8_t* base, const KeyLut* lut, uint32_t lut_copies, const ViewPlan* v, uint32_t n, int16_t* out, Unpack2Ws<NTH>* ws)
{
    switch (codec) {
    case RNSC_VBZ: return svb_unpack2<RNSC_VBZ, NTH>(base, lut, lut_copies, v, n, out, ws);
    case RNSC_VBZ0: return svb_unpack2<RNSC_VBZ0, NTH>(base, lut, lut_copies, v, n, out, ws);
    case RNSC_C1: return svb_unpack2<RNSC_C1, NTH>(base, lut, lut_copies, v, n, out, ws);
    case RNSC_C2: return svb_unpack2<RNSC_C2, NTH>(base, lut, lut_copies, v, n, out, ws);
    case RNSC_C3: return svb_unpack2<RNSC_C3, NTH>(base, lut, lut_copies, v, n, out, ws);
    case RNSC_C4: return svb_unpack2<RNSC_C4, NTH>(base, lut, lut_copies, v, n, out, ws);
    default: return svb_unpack2<RNSC_C5, NTH>(base, lut, lut_copies, v, n, out, ws);
    }
}

}  // namespace rnsc
