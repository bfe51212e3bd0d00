// rnsc_common.cuh -- shared definitions of the sm_100a kernels (compiled by nvcc; tests/emul can also
// compile the same sources with g++ through tests/emul/cuda_emul.h to step through them without a GPU).
#pragma once
#ifndef RNSC_EMUL
#include <cuda_runtime.h>
#include <stdint.h>
#define RNSC_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define RNSC_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif
#include "../../include/rnsc.h"

namespace rnsc {

constexpr int NT = 512;            // threads per CTA of the chunk kernels (16 warps; 2 CTAs per SM)
constexpr int NWARP = NT / 32;
constexpr int SPT = 8;             // samples per thread per tile
constexpr int TILE = NT * SPT;     // samples per tile

// svb16/svb16.h:11-21 and pgnano/svb16/svb16.h:11-22
__host__ __device__ inline uint32_t key_len1(uint32_t n) { return (n >> 3) + (((n & 7) + 7) >> 3); }
__host__ __device__ inline uint32_t key_len2(uint32_t n) { return (n >> 2) + (((n & 3) + 3) >> 2); }
__host__ __device__ inline bool codec_two_bit(int c) { return c == RNSC_VBZ0 || c == RNSC_C4 || c == RNSC_C5; }
__host__ __device__ inline uint32_t key_len(int c, uint32_t n) { return n == 0 ? 0 : (codec_two_bit(c) ? key_len2(n) : key_len1(n)); }
// zstd frames per chunk (container A.1-A.7) and logical streams per chunk
__host__ __device__ inline int codec_frames(int c) { return (c == RNSC_VBZ || c == RNSC_VBZ0) ? 1 : (c == RNSC_C1 ? 2 : (c == RNSC_C2 ? 3 : (c == RNSC_C3 ? 4 : 5))); }
__host__ __device__ inline int codec_streams(int c) { return (c == RNSC_VBZ || c == RNSC_VBZ0 || c == RNSC_C1) ? 2 : (c == RNSC_C2 ? 3 : (c == RNSC_C3 ? 4 : 5)); }
// Upper bound of all pre-zstd streams of a chunk (keys + payload): every codec stores at most 2 payload
// bytes per sample (C2: n + n).
__host__ __device__ inline uint64_t stream_bound(int c, uint64_t n) { return (uint64_t)key_len(c, (uint32_t)n) + 2 * n + 8; }

// svb16/encode_scalar.hpp:11-14, decode_scalar.hpp:14-17 (all arithmetic mod 2^16)
__host__ __device__ inline uint32_t zigzag_enc16(uint32_t v) { return ((v << 1) ^ (0u - ((v >> 15) & 1u))) & 0xFFFFu; }
__host__ __device__ inline uint32_t zigzag_dec16(uint32_t v) { return ((v >> 1) ^ (0u - (v & 1u))) & 0xFFFFu; }

__device__ __forceinline__ uint32_t ld8(const uint8_t* p) { return (uint32_t)*p; }
__device__ __forceinline__ uint32_t ldg8(const uint8_t* p) { return (uint32_t)__ldg(p); }
__device__ __forceinline__ uint32_t ldg32(const uint32_t* p) { return __ldg(p); }
// L1 prefetch of the 32-byte sector that holds p: a hint, issued a few hundred instructions ahead of a dependent load in the
// serial readers (the planner's threads otherwise pay an L2 round trip per sector they enter)
__device__ __forceinline__ void prefetch_sector(const void* p)
{
#ifndef RNSC_EMUL
    asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
#else
    (void)p;
#endif
}

// A contiguous byte range in shared OR global memory (generic address).
struct View { const uint8_t* p; uint32_t len; };

}  // namespace rnsc
