// rc_codec.cuh -- the legacy pgnano range coder (codec RC), one WARP per chunk.
//
// Reference: pgnano/compressor.cpp:39-88, decompressor.cpp:45-73 (driver), commonLib/codecs.cpp:6-20 (sign
// fold == zigzag), fqzcomp/simple_model.h:169-176,206-221,325-343 (adaptive order-0 model, STEP 8,
// MAX_FREQ 65504, fixed symbol order) and fqzcomp/clr.h:54-130 (carry-less range coder, 64-bit low,
// 32-bit range/code, TOP = 2^24). The coder is strictly serial per chunk (adaptive state), so the only
// parallelism inside a chunk is the 256-entry cumulative-frequency search: each lane owns 8 consecutive
// symbols of every model (shared memory, 16-byte row per lane), a warp scan of the lane sums replaces the
// reference's linear scan. low/range/code are replicated in every lane; lane 0 emits/reads bytes.
#pragma once
#include "block_prims.cuh"

namespace rnsc {

constexpr uint32_t RC_TOP = 1u << 24;
constexpr uint32_t RC_MAX_FREQ = (1u << 16) - 32;
constexpr uint32_t RC_STEP = 8;
constexpr int RC_HEADER = 10;                    // [is_raw][n: 8 bytes LE][pore type]  (header.h:15, compressor.cpp:21-36)
constexpr int RC_WARPS = 8;                      // warps (= chunks) per CTA

struct RcModelSet { uint16_t f[3][256]; };       // H, L0 (hi == 0), L1 (hi != 0)

__device__ __forceinline__ void rc_models_init(RcModelSet* m, int lane)
{
    for (int k = 0; k < 3; ++k) for (int i = 0; i < 8; ++i) m->f[k][lane * 8 + i] = 1;
    __syncwarp();
}

// cumulative frequency below symbol s and its frequency (all lanes get both)
__device__ __forceinline__ void rc_lookup(const uint16_t* f, int lane, uint32_t s, uint32_t* cum, uint32_t* fr)
{
    uint32_t mine[8], sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { mine[i] = f[lane * 8 + i]; sum += mine[i]; }
    const uint32_t ex = warp_inclusive_scan(sum) - sum;
    uint32_t part = ex, fq = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { if ((uint32_t)i < (s & 7)) part += mine[i]; if ((uint32_t)i == (s & 7)) fq = mine[i]; }
    *cum = __shfl_sync(0xffffffffu, part, (int)(s >> 3));
    *fr = __shfl_sync(0xffffffffu, fq, (int)(s >> 3));
}

// +8 on the coded symbol; halve everything when the total exceeds MAX_FREQ (simple_model.h:169-176,216-220)
__device__ __forceinline__ void rc_update(uint16_t* f, int lane, uint32_t s, uint32_t* tot)
{
    if ((uint32_t)lane == (s >> 3)) f[s] = (uint16_t)(f[s] + RC_STEP);
    *tot += RC_STEP;
    __syncwarp();
    if (*tot > RC_MAX_FREQ) {
        uint32_t sum = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { uint32_t v = f[lane * 8 + i]; v -= v >> 1; f[lane * 8 + i] = (uint16_t)v; sum += v; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        *tot = sum;
        __syncwarp();
    }
}

// One warp encodes one chunk. Returns bytes written (header included) or 0 when the reference would take
// its raw fallback (bytes_written > 2n, compressor.cpp:60-65).
__device__ inline uint32_t rc_encode_chunk(const int16_t* in, uint32_t n, uint8_t* out, uint8_t pore, RcModelSet* m)
{
    const int lane = threadIdx.x & 31;
    rc_models_init(m, lane);
    if (lane == 0) { out[0] = 0; for (int i = 0; i < 8; ++i) out[1 + i] = (uint8_t)((unsigned long long)n >> (8 * i)); out[9] = pore; }
    uint8_t* o = out + RC_HEADER;
    unsigned long long low = 0; uint32_t range = 0xFFFFFFFFu, written = 0, prev = 0;
    uint32_t tot[3] = {256, 256, 256};
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t mine = base + lane < n ? (uint16_t)in[base + lane] : 0;
        const uint32_t cntw = min(32u, n - base);
        for (uint32_t t = 0; t < cntw; ++t) {
            const uint32_t x = __shfl_sync(0xffffffffu, mine, (int)t);
            const uint32_t mv = zigzag_enc16((x - prev) & 0xFFFFu); prev = x;
            const uint32_t hi = mv >> 8, lo = mv & 0xFF;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int k = pass == 0 ? 0 : (hi ? 2 : 1);
                const uint32_t s = pass == 0 ? hi : lo;
                uint32_t cum, fr;
                rc_lookup(m->f[k], lane, s, &cum, &fr);
                range /= tot[k]; low += (uint32_t)(cum * range); range *= fr;           // clr.h:80-83
                while (range < RC_TOP) {                                              // clr.h:86-101
                    if ((uint8_t)((low ^ (low + range)) >> 56)) range = (((uint32_t)low | (RC_TOP - 1)) - (uint32_t)low);
                    if (lane == 0) o[written] = (uint8_t)(low >> 56);
                    written++; range <<= 8; low <<= 8;
                }
                rc_update(m->f[k], lane, s, &tot[k]);
            }
            if (written > 2 * n) return 0;
        }
    }
    if (lane == 0) for (int i = 0; i < 8; ++i) o[written + i] = (uint8_t)(low >> (56 - 8 * i));   // clr.h:65-69
    return RC_HEADER + written + 8;
}

// One warp decodes one chunk. Returns an rnsc_status.
__device__ inline int rc_decode_chunk(const uint8_t* in, uint32_t clen, uint32_t n, int16_t* out, RcModelSet* m)
{
    const int lane = threadIdx.x & 31;
    if (clen < RC_HEADER + 8) return RNSC_ERR_REMAINING;
    if (ldg8(in) & 1) return RNSC_ERR_RC_RAW;
    unsigned long long hn = 0; for (int i = 0; i < 8; ++i) hn |= (unsigned long long)ldg8(in + 1 + i) << (8 * i);
    if (hn != n) return RNSC_ERR_REMAINING;
    rc_models_init(m, lane);
    const uint8_t* ip = in + RC_HEADER; const uint8_t* iend = in + clen;
    unsigned long long low = 0; uint32_t range = 0xFFFFFFFFu, code = 0, prev = 0;
    for (int i = 0; i < 8; ++i) code = (code << 8) | ldg8(ip++);                     // clr.h:54-59 (u32: first 4 shift out)
    uint32_t tot[3] = {256, 256, 256};
    uint32_t hold = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t hi = 0, sym = 0;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int k = pass == 0 ? 0 : (hi ? 2 : 1);
            const uint16_t* f = m->f[k];
            range /= tot[k]; const uint32_t fq = code / range;                        // clr.h:115-117
            uint32_t mine[8], sum = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { mine[q] = f[lane * 8 + q]; sum += mine[q]; }
            const uint32_t ex = warp_inclusive_scan(sum) - sum;
            const uint32_t b = __ballot_sync(0xffffffffu, ex <= fq);
            const int owner = 31 - __clz((int)b);                                    // lane 0 always has ex = 0 <= fq
            uint32_t acc = ex, s_local = 7;
#pragma unroll
            for (int q = 0; q < 8; ++q) { if (s_local == 7 && q < 7 && acc + mine[q] > fq) s_local = (uint32_t)q; if (s_local == 7 && q < 7) acc += mine[q]; }
            // acc = cumulative below s_local (if nothing matched, s_local = 7 and acc = sum of the first 7)
            const uint32_t cum = __shfl_sync(0xffffffffu, acc, owner);
            const uint32_t sl = __shfl_sync(0xffffffffu, s_local, owner);
            const uint32_t fr = __shfl_sync(0xffffffffu, mine[0] * (sl == 0) + mine[1] * (sl == 1) + mine[2] * (sl == 2) + mine[3] * (sl == 3)
                                                           + mine[4] * (sl == 4) + mine[5] * (sl == 5) + mine[6] * (sl == 6) + mine[7] * (sl == 7), owner);
            sym = (uint32_t)owner * 8 + sl;
            const uint32_t temp = cum * range; low += temp; code -= temp; range *= fr;   // clr.h:119-123
            while (range < RC_TOP) {
                if ((uint8_t)((low ^ (low + range)) >> 56)) range = (((uint32_t)low | (RC_TOP - 1)) - (uint32_t)low);
                const uint32_t byte = ip < iend ? ldg8(ip) : 0; ip++;
                code = (code << 8) | byte; range <<= 8; low <<= 8;
            }
            rc_update(m->f[k], lane, sym, &tot[k]);
            if (pass == 0) hi = sym;
        }
        prev = (prev + zigzag_dec16((hi << 8) | sym)) & 0xFFFFu;
        // stage 32 samples across the warp, one coalesced store per 32
        if ((i & 31) == (uint32_t)lane) hold = prev;
        if ((i & 31) == 31 || i + 1 == n) { const uint32_t b0 = i & ~31u; if (b0 + lane <= i) out[b0 + lane] = (int16_t)(uint16_t)hold; }
    }
    return RNSC_OK;
}

}  // namespace rnsc
