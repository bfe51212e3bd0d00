// svb_codec.cuh -- the general splitter inverse of every codec (gather + zigzag^-1 + delta^-1), one CTA per chunk,
// NT threads x SPT consecutive samples per tile; works on generic (shared or global) stream pointers. The fast decoder
// uses svb_unpack2.cuh, the encoder svb_pack2.cuh.
//
// Reference (scalar, one sample at a time):
//   VBZ  svb16/encode_scalar.hpp:17-66, decode_scalar.hpp:36-75       VBZ0 pgnano/svb16/VBZ_0.hpp:61-172,186-290
//   C1   pgnano/svb16/C1.hpp:58-109,141-179     C2 C2.hpp:53-109,122-178     C3 C3.hpp:54-110,131-189
//   C4   C4.hpp:54-145,167-249                  C5 C5.hpp:58-152,174-258
// Here: each thread owns SPT consecutive samples, so its reads/writes in every stream are contiguous; a
// packed CTA exclusive scan of the per-thread unit counts gives the stream positions, a second scan
// undoes the delta (mod 2^16). Tiles move through shared memory so that global int16 traffic is 16-byte
// vectors whatever the chunk's alignment.
#pragma once
#include "block_prims.cuh"

namespace rnsc {

struct alignas(16) V16 { uint32_t x, y, z, w; };

// Shared-memory scratch of the splitter kernels.
struct SvbWs {
    unsigned long long scan64_a[NWARP + 1], scan64_b[NWARP + 1];
    uint32_t scan32_a[NWARP + 1], scan32_b[NWARP + 1];
    alignas(16) int16_t tile[TILE + 16];   // sample tile, phase-shifted to the global 16-byte alignment
    int err;
};

// ---- tile -> global, 16-byte vectors in the aligned body, scalars at the ragged ends; tile[phase + i] <-> g[i] ----
__device__ inline void tile_store(const int16_t* tile, int16_t* g, uint32_t m, int phase)
{
    int16_t* g0 = g - phase;
    const int ngroups = (phase + (int)m + 7) >> 3;
    for (int grp = threadIdx.x + 1; grp < ngroups; grp += NT) {
        const int lo = grp * 8;
        if (lo >= phase && lo + 8 <= phase + (int)m) *(V16*)(g0 + lo) = *(const V16*)(tile + lo);
        else for (int i = lo; i < lo + 8; ++i) if (i >= phase && i < phase + (int)m) g0[i] = tile[i];
    }
}

__device__ __forceinline__ uint32_t nib_units(uint32_t code) { return code == 3 ? 4u : code; }   // VBZ0: 0/1/2/4 nibbles

// =====================================================================================================
// decode: streams -> samples
// =====================================================================================================
// v[0] = keys, v[1..4] = payload streams in wire order. Returns an rnsc_status (same in every thread).
template <int CODEC>
__device__ int svb_unpack(const View* v, uint32_t n, int16_t* out, SvbWs* ws)
{
    constexpr bool TWO = (CODEC == RNSC_VBZ0 || CODEC == RNSC_C4 || CODEC == RNSC_C5);
    const int tid = threadIdx.x;
    if (tid == 0) ws->err = 0;
    const uint8_t* keys = v[0].p;
    const uint8_t *s1 = v[1].p, *s2 = v[2].p, *s3 = v[3].p, *s4 = v[4].p;
    const uint32_t l1 = v[1].len, l2 = v[2].len, l3 = v[3].len, l4 = v[4].len;
    uint32_t run1 = 0, run2 = 0, run3 = 0, carry = 0;
    bool oob = false;
    __syncthreads();
    for (uint32_t tb = 0; tb < n; tb += TILE) {
        const uint32_t m = min((uint32_t)TILE, n - tb);
        const uint32_t i0 = tb + (uint32_t)tid * SPT;
        const int cnt = i0 < n ? (int)min((uint32_t)SPT, n - i0) : 0;
        // ---- keys and unit counts ----
        uint32_t kb = 0;
        if (cnt > 0) {
            if (TWO) { kb = ld8(keys + (i0 >> 2)); if (cnt > 4) kb |= ld8(keys + (i0 >> 2) + 1) << 8; if (cnt < 8) kb &= (1u << (2 * cnt)) - 1; }
            else { kb = ld8(keys + (i0 >> 3)); if (cnt < 8) kb &= (1u << cnt) - 1; }
        }
        uint32_t c1 = 0, c2 = 0, c3 = 0;
        if (CODEC == RNSC_VBZ || CODEC == RNSC_C1) c1 = (uint32_t)cnt + __popc(kb);
        else if (CODEC == RNSC_C2) { c1 = (uint32_t)cnt; c2 = __popc(kb); }
        else if (CODEC == RNSC_C3) { c2 = __popc(kb); c1 = (uint32_t)cnt - c2; }
        else {
            const uint32_t lo = kb & 0x5555u, hi = (kb >> 1) & 0x5555u;
            const uint32_t n1 = __popc(lo & ~hi), n2 = __popc(hi & ~lo), n3 = __popc(lo & hi);
            if (CODEC == RNSC_VBZ0) c1 = n1 + 2 * n2 + 4 * n3; else { c1 = n1; c2 = n2; c3 = n3; }
        }
        unsigned long long tot64;
        const unsigned long long ex64 = block_exclusive_scan<unsigned long long>(
            (unsigned long long)c1 | ((unsigned long long)c2 << 20) | ((unsigned long long)c3 << 40), ws->scan64_a, &tot64);
        uint32_t p1 = run1 + (uint32_t)(ex64 & 0xFFFFF), p2 = run2 + (uint32_t)((ex64 >> 20) & 0xFFFFF), p3 = run3 + (uint32_t)(ex64 >> 40);
        run1 += (uint32_t)(tot64 & 0xFFFFF); run2 += (uint32_t)((tot64 >> 20) & 0xFFFFF); run3 += (uint32_t)(tot64 >> 40);
        // ---- gather, zigzag^-1, thread-local running sum ----
        uint32_t val[SPT]; uint32_t acc = 0;
#pragma unroll
        for (int j = 0; j < SPT; ++j) {
            uint32_t z = 0;
            if (j < cnt) {
                if (CODEC == RNSC_VBZ || CODEC == RNSC_C1) {
                    const uint32_t k = (kb >> j) & 1;
                    if (p1 + k < l1) { z = ld8(s1 + p1); if (k) z |= ld8(s1 + p1 + 1) << 8; } else oob = true;
                    p1 += 1 + k;
                } else if (CODEC == RNSC_C2) {
                    const uint32_t k = (kb >> j) & 1;
                    if (p1 < l1) z = ld8(s1 + p1); else oob = true;
                    p1++;
                    if (k) { if (p2 < l2) z |= ld8(s2 + p2) << 8; else oob = true; p2++; }
                } else if (CODEC == RNSC_C3) {
                    const uint32_t k = (kb >> j) & 1;
                    if (k) { if (p2 < l2 && p2 < l3) z = ld8(s2 + p2) | (ld8(s3 + p2) << 8); else oob = true; p2++; }
                    else { if (p1 < l1) z = ld8(s1 + p1); else oob = true; p1++; }
                } else if (CODEC == RNSC_VBZ0) {
                    const uint32_t code = (kb >> (2 * j)) & 3, nn = nib_units(code);
                    uint32_t x = 0;
                    for (uint32_t q = 0; q < nn; ++q) {
                        const uint32_t b = (p1 + q) >> 1;
                        if (b < l1) x |= ((ld8(s1 + b) >> (4 * ((p1 + q) & 1))) & 15u) << (4 * q); else oob = true;
                    }
                    p1 += nn;
                    z = code == 0 ? 0u : (x + (code == 1 ? 1u : (code == 2 ? 17u : 273u))) & 0xFFFFu;
                } else {   // C4 / C5
                    const uint32_t code = (kb >> (2 * j)) & 3;
                    uint32_t x = 0;
                    if (code == 1) { const uint32_t b = p1 >> 1; if (b < l1) x = (ld8(s1 + b) >> (4 * (p1 & 1))) & 15u; else oob = true; p1++; }
                    else if (code == 2) { if (p2 < l2) x = ld8(s2 + p2); else oob = true; p2++; }
                    else if (code == 3) { if (p3 < l3 && p3 < l4) x = (ld8(s4 + p3) << 8) + ld8(s3 + p3); else oob = true; p3++; }
                    if (CODEC == RNSC_C5 && code) x += (code == 1 ? 1u : (code == 2 ? 17u : 273u));
                    z = x & 0xFFFFu;
                }
            }
            acc += zigzag_dec16(z);
            val[j] = acc;
        }
        // ---- delta^-1: CTA scan of the per-thread sums (mod 2^16 arithmetic survives in 32 bits) ----
        uint32_t tot32;
        const uint32_t ex32 = block_exclusive_scan<uint32_t>(acc, ws->scan32_a, &tot32);
        const uint32_t base = carry + ex32;
        carry += tot32;
        const int phase = (int)(((uintptr_t)(out + tb) >> 1) & 7) + 8;
#pragma unroll
        for (int j = 0; j < SPT; ++j) if (j < cnt) ws->tile[phase + tid * SPT + j] = (int16_t)(uint16_t)(base + val[j]);
        __syncthreads();
        tile_store(ws->tile, out + tb, m, phase);
        __syncthreads();
    }
    // ---- the reference's "Remaining data at end of signal buffer" test (A.8) + out-of-bounds guard ----
    bool rem;
    if (CODEC == RNSC_VBZ || CODEC == RNSC_C1) rem = run1 != l1;
    else if (CODEC == RNSC_VBZ0) rem = ((run1 + 1) >> 1) != l1;
    else if (CODEC == RNSC_C2) rem = run2 != l2;
    else if (CODEC == RNSC_C3) rem = run2 != l3;
    else rem = run3 != l4;
    if (oob || (tid == 0 && rem)) ws->err = RNSC_ERR_REMAINING;
    __syncthreads();
    return ws->err;
}

}  // namespace rnsc
