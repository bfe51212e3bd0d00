// tests/emul/cuda_emul.h -- TEST INFRASTRUCTURE ONLY.
//
// A tiny CUDA-on-CPU execution model so that the *same* kernel sources under
// rawnanoporesignalcompression_b200/csrc/ can be compiled with g++ and stepped through in the build
// container, which has no GPU. One OS thread; every CUDA thread of a block is a ucontext fiber; fibers
// switch only at __syncthreads()/__syncwarp()/shuffle/ballot points, blocks run one after another.
// It exists to debug indexing and bit-twiddling logic before spending GPU minutes; it is never part of
// the product library (librnsc.so is built by nvcc only and fails loudly without a CUDA device).
#pragma once
#include <ucontext.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define RNSC_EMUL 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static const

struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() {} dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint3e { unsigned x, y, z; };
struct uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
extern uint3e threadIdx, blockIdx;
extern dim3 blockDim, gridDim;
extern unsigned char* emul_dyn_smem;
static const int warpSize = 32;

namespace emul {
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void block_barrier();
void named_barrier(int id, int count);
unsigned warp_exchange(unsigned mask, unsigned long long val, unsigned long long* all32);  // returns arrived mask
}

inline void __syncthreads() { emul::block_barrier(); }
inline int __syncthreads_or(int p) { static int acc; static int res; if (threadIdx.x == 0) acc = 0; emul::block_barrier(); acc |= (p != 0); emul::block_barrier(); res = acc; emul::block_barrier(); return res; }
inline int __syncthreads_count(int p) { static int acc; static int res; if (threadIdx.x == 0) acc = 0; emul::block_barrier(); acc += (p != 0); emul::block_barrier(); res = acc; emul::block_barrier(); return res; }
inline void __syncwarp(unsigned mask = 0xffffffffu) { unsigned long long all[32]; emul::warp_exchange(mask, 0, all); }
inline long long clock64() { static long long t = 0; return ++t; }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() {}

template <typename T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    unsigned long long all[32], raw = 0; memcpy(&raw, &v, sizeof(T)); emul::warp_exchange(mask, raw, all);
    int lane = threadIdx.x & 31; int base = lane & ~(width - 1); T r; memcpy(&r, &all[base + (src & (width - 1))], sizeof(T)); return r; }
template <typename T> inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    unsigned long long all[32], raw = 0; memcpy(&raw, &v, sizeof(T)); emul::warp_exchange(mask, raw, all);
    int lane = threadIdx.x & 31; int base = lane & ~(width - 1); int s = lane - (int)d; if (s < base) s = lane; T r; memcpy(&r, &all[s], sizeof(T)); return r; }
template <typename T> inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    unsigned long long all[32], raw = 0; memcpy(&raw, &v, sizeof(T)); emul::warp_exchange(mask, raw, all);
    int lane = threadIdx.x & 31; int base = lane & ~(width - 1); int s = lane + (int)d; if (s >= base + width) s = lane; T r; memcpy(&r, &all[s], sizeof(T)); return r; }
template <typename T> inline T __shfl_xor_sync(unsigned mask, T v, int m, int width = 32) {
    unsigned long long all[32], raw = 0; memcpy(&raw, &v, sizeof(T)); emul::warp_exchange(mask, raw, all);
    int lane = threadIdx.x & 31; T r; memcpy(&r, &all[lane ^ m], sizeof(T)); return r; }
inline unsigned __ballot_sync(unsigned mask, int p) {
    unsigned long long all[32]; unsigned arrived = emul::warp_exchange(mask, p ? 1 : 0, all); unsigned r = 0;
    for (int i = 0; i < 32; ++i) if (((arrived >> i) & 1) && all[i]) r |= 1u << i; return r; }
inline int __any_sync(unsigned mask, int p) { return __ballot_sync(mask, p) != 0; }
inline int __all_sync(unsigned mask, int p) { unsigned long long all[32]; unsigned arrived = emul::warp_exchange(mask, p ? 1 : 0, all);
    for (int i = 0; i < 32; ++i) if (((arrived >> i) & 1) && !all[i]) return 0; return 1; }
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    unsigned long long all[32]; unsigned arrived = emul::warp_exchange(mask, v, all); unsigned r = 0;
    for (int i = 0; i < 32; ++i) if ((arrived >> i) & 1) r += (unsigned)all[i]; return r; }
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    unsigned long long all[32]; unsigned arrived = emul::warp_exchange(mask, v, all); unsigned r = 0;
    for (int i = 0; i < 32; ++i) if (((arrived >> i) & 1) && (unsigned)all[i] > r) r = (unsigned)all[i]; return r; }
inline unsigned __activemask() { return 0xffffffffu; }

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) if (v >> i & 1) r |= 1u << (31 - i); return r; }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { unsigned long long v = ((unsigned long long)hi << 32) | lo; return (unsigned)(v >> (s & 31)); }
inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned s) { unsigned long long v = ((unsigned long long)hi << 32) | lo; return (unsigned)(v >> (s > 32 ? 32 : s)); }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { unsigned long long v = ((unsigned long long)hi << 32) | lo; return (unsigned)((v << (s & 31)) >> 32); }
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) { unsigned long long v = ((unsigned long long)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; ++i) { unsigned sel = (s >> (4 * i)) & 7; r |= (unsigned)((v >> (8 * sel)) & 0xff) << (8 * i); } return r; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcg(const T* p) { return *p; }
template <typename T> inline void __stcg(T* p, T v) { *p = v; }
template <typename T> inline T min(T a, T b) { return a < b ? a : b; }
template <typename T> inline T max(T a, T b) { return a > b ? a : b; }

template <typename T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T> inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

// ---- host runtime shims (enough for csrc/api.cu) ----
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaPeekAtLastError() { return 0; }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)1; return 0; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (cudaStream_t)1; return 0; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* least, int* greatest) { *least = 0; *greatest = -1; return 0; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)1; return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline const char* cudaGetErrorString(cudaError_t) { return "emul"; }
struct cudaDeviceProp { int multiProcessorCount = 4; size_t sharedMemPerBlockOptin = 227 * 1024; char name[64] = "emul"; };
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp(); return 0; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize, cudaFuncAttributePreferredSharedMemoryCarveout };
enum { cudaSharedmemCarveoutMaxShared = 100 };
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }

#define RNSC_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emul::launch(dim3(grid), dim3(block), (smem), [&]() { kernel(__VA_ARGS__); })
#define RNSC_DYN_SMEM(name) unsigned char* name = emul_dyn_smem
