// block_prims.cuh -- CTA-wide prefix scans built from warp shuffles (one smem word per warp).
#pragma once
#include "rnsc_common.cuh"

namespace rnsc {

template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// 32-bit flavour: shfl.up returns "source lane exists" as a predicate, so a step is a shuffle and a predicated add.
// All 32 lanes must call it.
template <>
__device__ __forceinline__ uint32_t warp_inclusive_scan<uint32_t>(uint32_t v)
{
#ifdef RNSC_EMUL
    const int lane = threadIdx.x & 31;
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += o; }
#else
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
        asm volatile("{\n.reg .pred p;\n.reg .b32 r;\nshfl.sync.up.b32 r|p, %0, %1, 0, 0xffffffff;\n@p add.u32 %0, %0, r;\n}" : "+r"(v) : "r"(d));
#endif
    return v;
}

// Exclusive scan over the NT threads of the CTA. `scratch` holds NWARP+1 elements of T in shared memory.
// Every thread of the CTA must call it. Returns the exclusive prefix; *total receives the CTA sum.
// Two __syncthreads(): scratch may be reused right after the call returns for a *different* array only.
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* scratch, T* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = warp_inclusive_scan(v);
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = lane < NWARP ? scratch[lane] : T(0);
        T winc = warp_inclusive_scan(w);
        if (lane < NWARP) scratch[lane] = winc - w;
        if (lane == NWARP - 1) scratch[NWARP] = winc;
    }
    __syncthreads();
    T base = scratch[warp];
    *total = scratch[NWARP];
    return base + inc - v;
}

}  // namespace rnsc
