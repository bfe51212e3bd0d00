// kernels.cuh -- __global__ entry points of librnsc (sm_100a). Launch geometry lives in api.cu.
#pragma once
#include "zstd_decode.cuh"
#include "svb_codec.cuh"
#include "svb_pack2.cuh"
#include "rc_codec.cuh"
#include "plan.cuh"
#include "fast_decode.cuh"
#include "zstd_encode.cuh"

namespace rnsc {

// ---------------------------------------------------------------------------------------------------
// Fused whole-chunk decode: container -> zstd frames -> streams (shared-memory arena) -> int16.
// Persistent CTAs pull chunk indices from a global counter. One CTA owns one chunk at a time.
// ---------------------------------------------------------------------------------------------------
struct DecodeParams {
    const uint8_t* comp; const uint64_t* comp_off; const uint32_t* n; const uint64_t* sample_off;
    int16_t* samples; int32_t* status; uint32_t nchunks; int codec;
    uint8_t* scratch; unsigned long long scratch_stride;   // per-CTA global scratch: [literals 128K+64][stream spill]
    uint32_t* counter; uint32_t arena_bytes;               // shared-memory arena after the workspace
    uint32_t spill_bytes;                                  // bytes of stream spill space per CTA
    const uint32_t* list; const uint32_t* count;           // optional: chunk indices to decode (planner's slow list)
};

constexpr uint32_t LIT_SCRATCH = ZSTD_BLOCK_MAX + 64;

union DecodeWs { ZstdWs z; SvbWs s; };

struct ChunkCtl {
    View view[5];
    uint32_t f_off[5], f_len[5], f_fcs[5], f_raw[5];
    uint8_t* f_dst[5];
    int status; uint32_t chunk;
};

template <int CODEC>
__device__ __forceinline__ int unpack_dispatch(const View* v, uint32_t n, int16_t* out, SvbWs* ws) { return svb_unpack<CODEC>(v, n, out, ws); }

__device__ inline int unpack_any(int codec, const View* v, uint32_t n, int16_t* out, SvbWs* ws)
{
    switch (codec) {
    case RNSC_VBZ: return svb_unpack<RNSC_VBZ>(v, n, out, ws);
    case RNSC_VBZ0: return svb_unpack<RNSC_VBZ0>(v, n, out, ws);
    case RNSC_C1: return svb_unpack<RNSC_C1>(v, n, out, ws);
    case RNSC_C2: return svb_unpack<RNSC_C2>(v, n, out, ws);
    case RNSC_C3: return svb_unpack<RNSC_C3>(v, n, out, ws);
    case RNSC_C4: return svb_unpack<RNSC_C4>(v, n, out, ws);
    default: return svb_unpack<RNSC_C5>(v, n, out, ws);
    }
}

__global__ void __launch_bounds__(NT, 2) k_decode_chunks(DecodeParams p)
{
    RNSC_DYN_SMEM(smem);
    DecodeWs* ws = (DecodeWs*)smem;
    ChunkCtl* ctl = (ChunkCtl*)(smem + ((sizeof(DecodeWs) + 15) & ~(size_t)15));
    uint8_t* arena = (uint8_t*)ctl + ((sizeof(ChunkCtl) + 15) & ~(size_t)15);
    const int tid = threadIdx.x;
    uint8_t* lit = p.scratch + (size_t)blockIdx.x * p.scratch_stride;
    uint8_t* spill = lit + LIT_SCRATCH;
    const int nf = codec_frames(p.codec);
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const uint32_t i = atomicAdd(p.counter, 1u);
            ctl->chunk = p.list ? (i < *p.count ? p.list[i] : 0xFFFFFFFFu) : i;
        }
        __syncthreads();
        const uint32_t c = ctl->chunk;
        if (c >= p.nchunks) break;
        const uint8_t* src = p.comp + p.comp_off[c];
        const uint32_t clen = (uint32_t)(p.comp_off[c + 1] - p.comp_off[c]);
        const uint32_t n = p.n[c];
        // ---- container + frame headers + placement (thread 0) ----
        if (tid == 0) {
            int st = RNSC_OK; uint32_t pos = 0, arena_used = 0, spill_used = 0;
            const uint64_t bound = stream_bound(p.codec, n);
            for (int f = 0; f < nf && st == RNSC_OK; ++f) {
                uint32_t fl;
                if (f + 1 < nf) {
                    if (pos + 8 > clen) { st = RNSC_ERR_CONTAINER; break; }
                    unsigned long long v = 0; for (int i = 0; i < 8; ++i) v |= (unsigned long long)ldg8(src + pos + i) << (8 * i);
                    pos += 8;
                    if (v > clen - pos) { st = RNSC_ERR_CONTAINER; break; }
                    fl = (uint32_t)v;
                } else fl = clen - pos;
                uint32_t hdr, known, ck; uint64_t fcs;
                int r = zstd_parse_frame_header(src + pos, fl, &hdr, &fcs, &known, &ck);
                if (r || !known) { st = RNSC_ERR_NOT_ZSTD; break; }       // ZSTD_getFrameContentSize must succeed
                if (fcs > bound) { st = RNSC_ERR_REMAINING; break; }      // more stream bytes than any valid chunk of n samples
                ctl->f_off[f] = pos; ctl->f_len[f] = fl; ctl->f_fcs[f] = (uint32_t)fcs;
                // a frame that is exactly one raw block is used in place (zero copy)
                uint32_t raw = 0;
                if (hdr + 3 <= fl) {
                    uint32_t bh = ldg8(src + pos + hdr) | (ldg8(src + pos + hdr + 1) << 8) | (ldg8(src + pos + hdr + 2) << 16);
                    if ((bh & 1) && ((bh >> 1) & 3) == 0 && (bh >> 3) == (uint32_t)fcs && hdr + 3 + (bh >> 3) + (ck ? 4 : 0) == fl) raw = hdr + 3;
                }
                ctl->f_raw[f] = raw;
                if (raw) ctl->f_dst[f] = nullptr;
                else if (arena_used + (uint32_t)fcs <= p.arena_bytes) { ctl->f_dst[f] = arena + arena_used; arena_used += ((uint32_t)fcs + 15) & ~15u; }
                else if (spill_used + (uint32_t)fcs <= p.spill_bytes) { ctl->f_dst[f] = spill + spill_used; spill_used += ((uint32_t)fcs + 15) & ~15u; }
                else { st = RNSC_ERR_REMAINING; break; }
                pos += fl;
            }
            ctl->status = st;
        }
        __syncthreads();
        // ---- frames ----
        for (int f = 0; f < nf && ctl->status == RNSC_OK; ++f) {
            if (ctl->f_raw[f]) continue;
            if (tid == 0) ws->z.err = 0;
            __syncthreads();
            int r = zstd_decode_frame(&ws->z, src + ctl->f_off[f], ctl->f_len[f], ctl->f_dst[f], ctl->f_fcs[f], lit);
            if (r && tid == 0) ctl->status = r;
            __syncthreads();
        }
        // ---- stream views (A.8: slices of the decompressed buffer) ----
        if (tid == 0 && ctl->status == RNSC_OK) {
            const uint32_t kl = key_len(p.codec, n);
            for (int k = 0; k < 5; ++k) { ctl->view[k].p = src; ctl->view[k].len = 0; }
            auto fptr = [&](int f) -> const uint8_t* { return ctl->f_raw[f] ? src + ctl->f_off[f] + ctl->f_raw[f] : ctl->f_dst[f]; };
            if (nf == 1) {
                if (ctl->f_fcs[0] < kl) ctl->status = RNSC_ERR_REMAINING;
                else { ctl->view[0].p = fptr(0); ctl->view[0].len = kl; ctl->view[1].p = fptr(0) + kl; ctl->view[1].len = ctl->f_fcs[0] - kl; }
            } else {
                if (ctl->f_fcs[0] != kl) ctl->status = RNSC_ERR_REMAINING;
                for (int f = 0; f < nf; ++f) { ctl->view[f].p = fptr(f); ctl->view[f].len = ctl->f_fcs[f]; }
            }
            if (n == 0) { uint32_t tot = 0; for (int f = 0; f < nf; ++f) tot += ctl->f_fcs[f]; if (tot) ctl->status = RNSC_ERR_REMAINING; }
        }
        __syncthreads();
        if (ctl->status == RNSC_OK && n > 0) {
            int r = unpack_any(p.codec, ctl->view, n, p.samples + p.sample_off[c], &ws->s);
            if (r && tid == 0) ctl->status = r;
        }
        __syncthreads();
        if (tid == 0) p.status[c] = ctl->status;
    }
}

// ---------------------------------------------------------------------------------------------------
// Fast whole-chunk decode of planned chunks (plan.cuh, fast_decode.cuh). Persistent CTAs pull chunk indices of
// one size class from the planner's list. The dynamic shared memory is [FastWs | key-byte table (C4/C5) | region]; the region holds the
// staged compressed chunk and every decoded stream.
// ---------------------------------------------------------------------------------------------------
template <int NTH>
__global__ void __launch_bounds__(NTH, 1024 / NTH) k_decode_fast(FastParams P)
{
    RNSC_DYN_SMEM(smem);
    FastWs<NTH>* ws = (FastWs<NTH>*)smem;
    KeyLut* lut = (KeyLut*)(smem + ((sizeof(FastWs<NTH>) + 15) & ~(size_t)15));
    uint8_t* region = (uint8_t*)lut + unpack2_lut_bytes(P.codec);
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(&ws->mbar, 1); ws->stash_tag = FAST_NONE; ws->stash_bad = 0u; const uint32_t i = atomicAdd(P.counter, 1u); ws->next = i < *P.count ? P.list[i] : FAST_NONE; }
    unpack2_init<NTH>(P.codec, lut);
    __syncthreads();
    uint32_t parity = 0;
    PhaseClock pc; pc.start(P.prof, P.prof_warp ? (uint32_t)NTH - 32u * min(P.prof_warp, (uint32_t)NTH / 32u) : 0u);
    const unsigned long long gt0 = (P.span && tid == 0) ? global_ns() : 0ull;      // profiling: this CTA's life span and chunk count
    uint32_t nch = 0;
    for (;;) {
        const uint32_t c = ws->next;            // thread 0 replaces it (prefetch) only after the first barrier of the chunk
        if (c == FAST_NONE) break;
        fast_decode_chunk<NTH>(ws, region, P, c, parity, pc); ++nch;
        parity ^= 1u;
        const uint32_t n = ws->n;
        if (ws->status == RNSC_OK && n > 0) {
            const uint32_t lo = ws->head.lut_off;      // the chunk's replicated key-byte table (fast_decode_chunk built it), or none
            const int r = unpack2_any<NTH>(P.codec, region, lo ? (const KeyLut*)(region + lo) : lut, lo ? key_lut_copies(NTH) : 1u, ws->head.view, n, P.samples + ws->sample_off, &ws->u2);
            if (r && tid == 0) ws->status = r;
        }
        __syncthreads();
        pc.mark(PH_UNPACK);
        if (tid == 0) { P.status[c] = ws->status; if (P.prof) atomicAdd(P.prof + PH_CHUNKS, 1ull); }
        __syncthreads();                        // the region and ws->status are reused by the next chunk
        pc.mark(PH_OTHER);
    }
    if (P.span && tid == 0 && blockIdx.x < 1024) { unsigned long long* sp = P.span + 3 * (size_t)blockIdx.x; sp[0] = gt0; sp[1] = global_ns(); sp[2] = nch; }
}

// ---------------------------------------------------------------------------------------------------
// Splitter stage only: streams <-> samples (one CTA per chunk, grid-stride over chunks)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 2) k_svb_unpack(int codec, const uint8_t* streams, const uint64_t* stream_off, const uint32_t* n,
                                                      const uint64_t* sample_off, uint32_t nchunks, int16_t* samples, int32_t* status)
{
    __shared__ SvbWs ws;
    __shared__ View view[5];
    for (uint32_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < 5) { const uint64_t a = stream_off[5 * (size_t)c + threadIdx.x], b = stream_off[5 * (size_t)c + threadIdx.x + 1];
                               view[threadIdx.x].p = streams + a; view[threadIdx.x].len = (uint32_t)(b - a); }
        __syncthreads();
        const uint32_t nn = n[c];
        int r = RNSC_OK;
        if (nn == 0) { uint32_t tot = 0; for (int k = 0; k < 5; ++k) tot += view[k].len; if (tot) r = RNSC_ERR_REMAINING; }
        else if (view[0].len != key_len(codec, nn)) r = RNSC_ERR_REMAINING;
        else r = unpack_any(codec, view, nn, samples + sample_off[c], &ws);
        if (threadIdx.x == 0) status[c] = r;
    }
}

// ---------------------------------------------------------------------------------------------------
// Second-generation splitter kernels (svb_pack2.cuh): one CTA of P2_NT threads per chunk
// ---------------------------------------------------------------------------------------------------
template <int CODEC>
__global__ void __launch_bounds__(P2_NT, 1024 / P2_NT) k_svb_sizes2(const int16_t* samples, const uint64_t* sample_off, const uint32_t* n, uint32_t nchunks, uint32_t* sizes)
{
    __shared__ Pack2Ws<P2_NT> ws;
    for (uint32_t c = blockIdx.x; c < nchunks; c += gridDim.x)
        svb_sizes2<CODEC, P2_NT>(samples + sample_off[c], n[c], sizes + 5 * (size_t)c, &ws);
}

template <int CODEC>
__global__ void __launch_bounds__(P2_NT, 1024 / P2_NT) k_svb_pack2(const int16_t* samples, const uint64_t* sample_off, const uint32_t* n, uint32_t nchunks,
                                                        uint8_t* streams, const uint64_t* stream_off)
{
    RNSC_DYN_SMEM(stage);
    __shared__ Pack2Ws<P2_NT> ws;
    for (uint32_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        if (threadIdx.x < 5) ws.dst[threadIdx.x] = streams + stream_off[5 * (size_t)c + threadIdx.x];
        __syncthreads();
        svb_pack2<CODEC, P2_NT>(samples + sample_off[c], n[c], ws.dst, &ws, stage);
    }
}

// ---------------------------------------------------------------------------------------------------
// Splitter encode in ONE pass over the samples, for consumers that do not need the streams dense (the GPU entropy stage):
// every chunk owns a SLAB sized by the streams' upper bounds (keys: exact; nibbles <= n/2; bytes <= n each; 2n for the
// one-data-stream codecs), stream k at a fixed offset inside it, so the scatter needs neither the size pass nor the prefix
// scan of the dense layout -- the samples are read from HBM once (the dense entry point rnsc_dev_svb_encode reads them
// twice: k_svb_sizes2, scan, k_svb_pack2). Start and size of every stream are left in stream_start / stream_size.
// (Tried instead: sizes + scan + scatter fused into one kernel with a chunk-level decoupled look-back. The chunk's second
// read did NOT come from L2 -- 1184 resident CTAs x 200 KB exceed its 126 MB -- and the look-back chain cost more than the
// two small launches it saved: 0.78 ms against 0.69 ms on 6000 reads, DRAM reads unchanged at 2 x the samples.)
// ---------------------------------------------------------------------------------------------------
__host__ __device__ inline void slab_layout(int codec, uint32_t n, uint32_t rel[6])
{
    const uint32_t kl = key_len(codec, n), pad = 16;
    rel[0] = 0; rel[1] = kl;                                          // keys are exact: a one-frame codec's payload follows them directly (A.1, A.2)
    uint32_t cap[4] = {0, 0, 0, 0};
    if (codec == RNSC_VBZ || codec == RNSC_C1 || codec == RNSC_VBZ0) cap[0] = 2 * n + pad;
    else if (codec == RNSC_C2) { cap[0] = n + pad; cap[1] = n + pad; }
    else if (codec == RNSC_C3) { cap[0] = n + pad; cap[1] = n + pad; cap[2] = n + pad; }
    else { cap[0] = (n + 1) / 2 + pad; cap[1] = n + pad; cap[2] = n + pad; cap[3] = n + pad; }
    for (int k = 0; k < 4; ++k) rel[k + 2] = rel[k + 1] + cap[k];
    rel[5] = (rel[5] + 15u) & ~15u;                                   // slab size
}

// slab size (and output slot capacity = the reference's output capacity, see k_zenc_slot_sizes) of every chunk
__global__ void k_slab_sizes(const uint32_t* n, uint32_t nchunks, int codec, uint32_t* slab)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nchunks) return;
    uint32_t rel[6]; slab_layout(codec, n[i], rel);
    slab[i] = rel[5];
}

template <int CODEC>
__global__ void __launch_bounds__(P2_NT, 1024 / P2_NT) k_svb_pack_slab(const int16_t* samples, const uint64_t* sample_off, const uint32_t* n, uint32_t nchunks,
                                                                       uint8_t* streams, const uint64_t* slab_off, unsigned long long* stream_start, uint32_t* stream_size)
{
    RNSC_DYN_SMEM(stage);
    __shared__ Pack2Ws<P2_NT> ws;
    __shared__ uint32_t runs[3];
    for (uint32_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const uint32_t nn = n[c];
        uint32_t rel[6]; slab_layout(CODEC, nn, rel);
        const unsigned long long base = slab_off[c];
        if (threadIdx.x < 5) ws.dst[threadIdx.x] = streams + base + rel[threadIdx.x];
        __syncthreads();
        svb_pack2<CODEC, P2_NT>(samples + sample_off[c], nn, ws.dst, &ws, stage, runs);
        if (threadIdx.x == 0) {
            uint32_t t1 = runs[0], t2 = runs[1], t3 = runs[2], sizes[5];
            stream_sizes(CODEC, nn, t1, t2, t3, sizes);
            for (int k = 0; k < 5; ++k) { stream_start[5 * (size_t)c + k] = base + rel[k]; stream_size[5 * (size_t)c + k] = sizes[k]; }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// Device prefix scan that assigns output offsets: out[0] = 0, out[i+1] = out[i] + in[i].
// One CTA walks the array, SCAN_ITEMS consecutive entries per thread per step (count is ~5 x chunks: a few 100k).
// ---------------------------------------------------------------------------------------------------
constexpr int SCAN_ITEMS = 16;
__global__ void __launch_bounds__(NT, 1) k_exclusive_scan_u32(const uint32_t* in, unsigned long long count, unsigned long long* out)
{
    __shared__ unsigned long long scratch[2][NWARP + 1];
    unsigned long long run = 0; int flip = 0;
    for (unsigned long long base = 0; base < count; base += (unsigned long long)NT * SCAN_ITEMS, flip ^= 1) {
        const unsigned long long i0 = base + (unsigned long long)threadIdx.x * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS];
        unsigned long long sum = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = i0 + k < count ? in[i0 + k] : 0u; sum += v[k]; }
        unsigned long long tot;
        unsigned long long ex = run + block_exclusive_scan<unsigned long long>(sum, scratch[flip], &tot);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) { if (i0 + k < count) out[i0 + k] = ex; ex += v[k]; }
        run += tot;
    }
    if (threadIdx.x == 0) out[count] = run;
}

// {sequence pool entries the planner asked for, chunks} of this decode -> the host's pinned mirror (api.cu), written directly
__global__ void k_pool_demand(const unsigned long long* cursor, unsigned long long nchunks, unsigned long long* out)
{
    out[0] = *cursor; out[1] = nchunks;
    __threadfence_system();
}

// ---------------------------------------------------------------------------------------------------
// Range coder: one warp per chunk
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * RC_WARPS) k_rc_encode(const int16_t* samples, const uint64_t* sample_off, const uint32_t* n, uint32_t nchunks,
                                                             uint8_t pore, uint8_t* comp, const uint64_t* slot_off, uint32_t* comp_size, int32_t* status)
{
    __shared__ RcModelSet models[RC_WARPS];
    const uint32_t w = threadIdx.x >> 5;
    for (uint32_t c = blockIdx.x * RC_WARPS + w; c < nchunks; c += gridDim.x * RC_WARPS) {
        const uint32_t r = rc_encode_chunk(samples + sample_off[c], n[c], comp + slot_off[c], pore, &models[w]);
        if ((threadIdx.x & 31) == 0) { comp_size[c] = r; status[c] = r ? RNSC_OK : RNSC_ERR_RC_RAW; }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(32 * RC_WARPS) k_rc_decode(const uint8_t* comp, const uint64_t* comp_off, const uint32_t* n, const uint64_t* sample_off,
                                                             uint32_t nchunks, int16_t* samples, int32_t* status)
{
    __shared__ RcModelSet models[RC_WARPS];
    const uint32_t w = threadIdx.x >> 5;
    for (uint32_t c = blockIdx.x * RC_WARPS + w; c < nchunks; c += gridDim.x * RC_WARPS) {
        const int r = rc_decode_chunk(comp + comp_off[c], (uint32_t)(comp_off[c + 1] - comp_off[c]), n[c], samples + sample_off[c], &models[w]);
        if ((threadIdx.x & 31) == 0) status[c] = r;
        __syncwarp();
    }
}

}  // namespace rnsc
