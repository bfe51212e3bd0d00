// zstd_encode.cuh -- GPU entropy stage of the encoders: every pre-zstd stream of a chunk becomes one VALID Zstandard frame
// (RFC 8878) on the device, so that whole-chunk encode never leaves HBM.
//
// The reference ends every encoder in ZSTD_compress(dst, cap, stream, len, 1) (pod5_format/signal_compression.cpp:65;
// pgnano/svb16/C5.hpp:337,356,373,390,407 and the same calls in VBZ_0/C1..C4.hpp). libzstd's source is not part of the
// reference and its level-1 bytes change between versions, so this stage does NOT try to be
// byte-identical to it: that is what the host-libzstd mode of rnsc_compress_batch is for. It writes what level 1 writes
// for these streams in substance (order-0 Huffman literals with no sequences for keys and
// M, raw blocks for the incompressible nibble / byte streams): per frame one block that is
//   * a RAW block (frame content <= 128 KiB: one block; larger: several), or
//   * an RLE block (one repeated byte), or
//   * a COMPRESSED block = Huffman-coded literals (1 or 4 streams, canonical zstd weights, length-limited to 11 bits,
//     tree description direct or FSE-compressed) + an empty sequences section,
// whichever is smallest. Any conforming decoder (libzstd, the reference's decompress_signal_*, this library's own fast
// path) returns the same stream bytes, hence the same samples.
//
// Three kernels after the splitter:
//   k_zenc_histogram  one CTA per chunk: byte histogram of every frame (shared-memory atomics, one copy per warp);
//   k_zenc_build      one WARP per frame, all frames of the batch at once: rank sort of the symbols by count and the canonical
//                     codes (zstd's table order) across the lanes, the inherently serial steps (two-queue Huffman tree,
//                     length limit, FSE-compressed weights of the tree description) on lane 0 out of shared memory -- 10^5
//                     short serial chains overlap each other instead of stalling a CTA each;
//   k_zenc_emit       one CTA per chunk: per frame, bit lengths per thread, a suffix scan (zstd's Huffman streams are
//                     written last symbol first), bits packed in registers and written straight into the chunk's slot
//                     (plain stores for a thread's whole words, atomic OR for the ragged ones), headers, container (A.1-A.8:
//                     u64 size prefixes); k_zenc_gather then compacts the slots.
#pragma once
#include "block_prims.cuh"
#include "svb_codec.cuh"

namespace rnsc {

constexpr int ZE_NT = 256;                       // threads per CTA of the histogram / emit kernels
#ifndef RNSC_ZE_EMIT_CTAS
#define RNSC_ZE_EMIT_CTAS 6
#endif
constexpr int ZE_EMIT_CTAS = RNSC_ZE_EMIT_CTAS;   // resident emit CTAs per SM the register allocation aims at
constexpr uint32_t ZE_BLOCK_MAX = 128u << 10;    // Block_Maximum_Size
constexpr uint32_t ZE_HUF_MIN = 64;              // shorter frames are not worth a tree description: raw
constexpr uint32_t ZE_LOG_MAX = 11;              // HUF_TABLELOG_DEFAULT of the literals' Huffman codes
enum : uint32_t { ZM_RAW = 0, ZM_RLE = 1, ZM_HUF = 2 };

// What k_zenc_build leaves per frame for k_zenc_emit.
struct ZencFrame {
    uint32_t mode;               // ZM_*
    uint32_t rle_byte;
    uint32_t log;                // table log L of the code (max length)
    uint32_t desc_len;           // bytes of the tree description
    uint32_t est_bytes;          // estimated literal payload (sum count x length / 8), for the raw-vs-compressed decision
    uint32_t pad[3];
    uint16_t code[256];          // length << 11 | code bits (length 0: symbol absent)
    uint8_t desc[136];           // tree description: header byte + direct nibbles or FSE-compressed weights
};

__host__ __device__ inline uint32_t ze_highbit(uint32_t v) { uint32_t r = 0; while (v >>= 1) ++r; return r; }

// frame f of a chunk = stream f, except for the codecs that put keys and payload in ONE frame (VBZ, VBZ0: A.1, A.2)
// (sz == nullptr: dense streams, stream k ends where k+1 starts; else slab layout: start and size per stream, kernels.cuh)
__host__ __device__ inline void ze_frame_range(int nf, const uint64_t* so, const uint32_t* sz, int f, uint64_t& a, uint64_t& b)
{
    if (sz) { a = so[f]; b = a + sz[f] + (nf == 1 ? sz[1] : 0u); }
    else if (nf == 1) { a = so[0]; b = so[2]; } else { a = so[f]; b = so[f + 1]; }
}

// -------------------------------------------------------------------------------------------------------------
// 1. histograms
// -------------------------------------------------------------------------------------------------------------
struct ZencParams {
    const uint8_t* streams; const uint64_t* stream_off;      // splitter output: stream k of chunk i = [stream_off[5i+k], stream_off[5i+k+1])
    const uint32_t* stream_size = nullptr;                   // slab layout: stream k of chunk i = stream_size[5i+k] bytes at stream_off[5i+k]
    const uint32_t* n; uint32_t nchunks; int codec;
    uint32_t* hist;                                          // [nchunks][nf][256]
    ZencFrame* frames;                                       // [nchunks][nf]
    uint8_t* slots; const uint64_t* slot_off;                // per chunk output slot (capacity slot_off[i+1] - slot_off[i])
    uint32_t* out_size; int32_t* status;                     // bytes written per chunk; rnsc_status per chunk
};

__global__ void __launch_bounds__(ZE_NT) k_zenc_histogram(ZencParams P)
{
    __shared__ uint32_t h[ZE_NT / 32][256];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int nf = codec_frames(P.codec);
    for (uint32_t c = blockIdx.x; c < P.nchunks; c += gridDim.x) {
        const uint64_t* so = P.stream_off + 5 * (size_t)c;
        const uint32_t* ssz = P.stream_size ? P.stream_size + 5 * (size_t)c : nullptr;
        for (int f = 0; f < nf; ++f) {
            uint64_t a, b; ze_frame_range(nf, so, ssz, f, a, b);
            const uint32_t size = (uint32_t)(b - a);
            uint32_t* out = P.hist + ((size_t)c * nf + f) * 256;
            if (size < ZE_HUF_MIN || size > ZE_BLOCK_MAX) continue;        // raw without looking (k_zenc_build applies the same test)
            for (int i = tid; i < (ZE_NT / 32) * 256; i += ZE_NT) (&h[0][0])[i] = 0;
            __syncthreads();
            const uint8_t* p = P.streams + a;
            // aligned 16-byte middle, byte-wise ragged ends
            const uint32_t head = min(size, (uint32_t)((16 - ((uintptr_t)p & 15)) & 15));
            if ((uint32_t)tid < head) atomicAdd(&h[warp][p[tid]], 1u);
            const uint32_t nv = (size - head) >> 4;
            const V16* v = (const V16*)(p + head);
            for (uint32_t i = tid; i < nv; i += ZE_NT) {
                const V16 x = v[i];
                const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    atomicAdd(&h[warp][w[k] & 0xFF], 1u); atomicAdd(&h[warp][(w[k] >> 8) & 0xFF], 1u);
                    atomicAdd(&h[warp][(w[k] >> 16) & 0xFF], 1u); atomicAdd(&h[warp][w[k] >> 24], 1u);
                }
            }
            const uint32_t done = head + (nv << 4);
            if (done + (uint32_t)tid < size) atomicAdd(&h[warp][p[done + tid]], 1u);
            __syncthreads();
            uint32_t s = 0;
#pragma unroll
            for (int k = 0; k < ZE_NT / 32; ++k) s += h[k][tid];
            out[tid] = s;                                                   // ZE_NT == 256
            __syncthreads();
        }
    }
}

// -------------------------------------------------------------------------------------------------------------
// 2. serial part, one thread per frame
// -------------------------------------------------------------------------------------------------------------
// LSB-first bit writer into a small word buffer in shared memory (FSE table description + FSE bitstream of the weights)
struct ZeBits {
    uint32_t* p; uint32_t cap; uint32_t pos; unsigned long long acc; uint32_t fill;
    __device__ __forceinline__ void init(uint32_t* buf, uint32_t cap_words) { p = buf; cap = cap_words; pos = 0; acc = 0; fill = 0; }
    __device__ __forceinline__ void add(uint32_t v, uint32_t nb)
    {
        acc |= (unsigned long long)(v & ((1u << nb) - 1u)) << fill; fill += nb;
        if (fill >= 32) { if (pos < cap) p[pos] = (uint32_t)acc; ++pos; acc >>= 32; fill -= 32; }
    }
    __device__ __forceinline__ void pad_to_byte() { fill = (fill + 7u) & ~7u; if (fill >= 32) { if (pos < cap) p[pos] = (uint32_t)acc; ++pos; acc >>= 32; fill -= 32; } }
    __device__ __forceinline__ uint32_t bytes() const { return 4 * pos + ((fill + 7) >> 3); }
    __device__ __forceinline__ void flush() { if (fill && pos < cap) p[pos] = (uint32_t)acc; }
};

// FSE-compressed Huffman weights (RFC 8878 4.2.1.1 / HUF_compressWeights): returns the bytes written to out (<= cap) or 0
// when this representation is not usable (fewer than two weights, one distinct value, no gain possible).
struct ZeFseWs { uint8_t tsym[64]; uint16_t state_tab[64]; uint32_t cumul[14]; int dnb[13], dfs[13], norm[13]; uint32_t cnt[13]; uint32_t out[36]; };   // scratch of ze_fse_weights (shared memory: dynamically indexed local arrays cost an L2 round trip per access)

// One thread. F.cnt[0..maxw] = how often each weight value occurs among w[0..n) (counted by the warp). The description is
// left in F.out (bytes, little endian words); returns its length (<= 127) or 0.
__device__ inline uint32_t ze_fse_weights(const uint8_t* w, uint32_t n, uint32_t maxw, ZeFseWs& F)
{
    if (n < 2) return 0;
    const uint32_t* cnt = F.cnt;
    uint32_t present = 0, big = 0;
    for (uint32_t s = 0; s <= maxw; ++s) { if (cnt[s]) ++present; if (cnt[s] > cnt[big]) big = s; }
    if (present < 2) return 0;
    const uint32_t tl = n >= 32 ? 6u : 5u, size = 1u << tl;           // any log in 5..6 is valid; the alphabet has <= 13 symbols
    // normalised counts: proportional, every present symbol at least 1, the remainder to the most frequent one
    int* norm = F.norm; int sum = 0;
    for (uint32_t s = 0; s <= maxw; ++s) { norm[s] = cnt[s] ? (int)max(1u, (cnt[s] * size) / n) : 0; sum += norm[s]; }
    norm[big] += (int)size - sum;
    if (norm[big] < 1) return 0;
    // ---- table description (FSE_writeNCount; mirror of fse_read_ncount in zstd_decode.cuh) ----
    ZeBits b; b.init(F.out, 36);
    b.add(tl - 5, 4);
    {
        int remaining = (int)size + 1, threshold = (int)size, nbBits = (int)tl + 1; bool prev0 = false;
        uint32_t s = 0;
        while (remaining > 1) {
            if (prev0) {
                uint32_t start = s;
                while (norm[s] == 0) ++s;                             // remaining > 1: a non-zero count follows
                while (s >= start + 3) { b.add(3, 2); start += 3; }
                b.add(s - start, 2);
            }
            int count = norm[s++];
            const int mx = (2 * threshold - 1) - remaining;
            remaining -= count;
            ++count;                                                  // +1: room for the "less than one" value -1 (unused here)
            if (count >= threshold) count += mx;
            b.add((uint32_t)count, (uint32_t)nbBits - (count < mx ? 1u : 0u));
            prev0 = count == 1;
            while (remaining < threshold) { --nbBits; threshold >>= 1; }
        }
    }
    b.pad_to_byte();                                                  // the bitstream of the weights starts at the next byte
    // ---- compression table (FSE_buildCTable) ----
    uint8_t* tsym = F.tsym; uint16_t* state_tab = F.state_tab; uint32_t* cumul = F.cumul;
    int* dnb = F.dnb; int* dfs = F.dfs;
    {
        const uint32_t step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
        uint32_t pos = 0;
        cumul[0] = 0;
        for (uint32_t s = 0; s <= maxw; ++s) { cumul[s + 1] = cumul[s] + (uint32_t)norm[s]; for (int i = 0; i < norm[s]; ++i) { tsym[pos] = (uint8_t)s; pos = (pos + step) & mask; } }
        for (uint32_t u = 0; u < size; ++u) { const uint32_t s = tsym[u]; state_tab[cumul[s]++] = (uint16_t)(size + u); }
        int total = 0;
        for (uint32_t s = 0; s <= maxw; ++s) {
            if (norm[s] == 0) { dnb[s] = (int)(((tl + 1) << 16) - size); dfs[s] = 0; }
            else if (norm[s] == 1) { dnb[s] = (int)((tl << 16) - size); dfs[s] = total - 1; total += 1; }
            else {
                const uint32_t mbo = tl - ze_highbit((uint32_t)norm[s] - 1);
                dnb[s] = (int)((mbo << 16) - ((uint32_t)norm[s] << mbo)); dfs[s] = total - norm[s]; total += norm[s];
            }
        }
    }
    // ---- weights, last to first, two interleaved states (FSE_compress_usingCTable) ----
    auto init_state = [&](uint32_t sym) -> uint32_t {                // FSE_initCState2: the symbol's state that costs the most bits next time
        const uint32_t nb = (uint32_t)(dnb[sym] + (1 << 15)) >> 16;
        const uint32_t value = (nb << 16) - (uint32_t)dnb[sym];
        return state_tab[(int)(value >> nb) + dfs[sym]];
    };
    auto encode = [&](uint32_t& st, uint32_t sym) { const uint32_t nb = (uint32_t)((int)st + dnb[sym]) >> 16; b.add(st, nb); st = state_tab[(int)(st >> nb) + dfs[sym]]; };
    uint32_t ip = n, s1, s2;
    if (n & 1) { s1 = init_state(w[--ip]); s2 = init_state(w[--ip]); encode(s1, w[--ip]); }
    else { s2 = init_state(w[--ip]); s1 = init_state(w[--ip]); }
    while (ip > 0) { encode(s2, w[--ip]); encode(s1, w[--ip]); }
    b.add(s2, tl); b.add(s1, tl);
    b.add(1, 1);                                                      // end mark
    b.flush();
    const uint32_t total_bytes = b.bytes();
    if (b.pos >= b.cap || total_bytes > 127) return 0;
    return total_bytes;
}

// Scratch of one frame's build, in shared memory: thread-private arrays of this size in LOCAL memory (the first version,
// one thread per frame) made every access an L2 round trip -- 2.7 ms for the 110k frames of a 20k-read batch.
// 2.6 KB per frame (4.5 KB with every array on its own: three instead of five CTAs per SM, and the kernel is bound by the
// frames in flight: 2.35 -> 1.9 ms at four CTAs on the 20k-read batch). The internal nodes' weights are dead once the tree
// stands, before the code lengths, the weight bytes and the FSE scratch of the tree description come to life: they share
// their bytes. A symbol's zstd weight replaces its code length in place (a lane converts the eight symbols it has just
// read). A parent index is below 256, and an internal node's depth replaces its parent index in place (root first).
struct ZencBuildWs {
    uint32_t key[256];           // present symbols: count << 8 | symbol, sorted ascending
    union {
        uint32_t wgt[256];       // stage B: weights of the internal nodes of the Huffman tree
        struct {
            union { uint8_t len[256]; uint8_t wts[256]; };   // stages C-E: code length per symbol; E-G: its zstd weight
            ZeFseWs fse;         // stages E-G: scratch of the FSE-compressed tree description
        };
    };
    uint8_t par[512];            // parent (index among the internal nodes) of leaf i; of internal node j at [256 + j], then (end of stage B) the depth of node j
    uint32_t m, maxs, L, kraft, nfull, maxw, dl, have, size, live;   // state of the frame between the stages
    unsigned long long bits;
};
static_assert(sizeof(ZeFseWs) + 256 <= 1024, "the tree description's scratch must fit behind the internal nodes' weights");
// (Tried on the serial stages: the first two elements of either Huffman queue in registers so that the refill load is issued a
// pick ahead, node depths by pointer jumping with all lanes instead of the serial walk, the next pair's weights and table
// rows fetched while the FSE states make their round trip. Fewer dependent shared-memory round trips, but 966 M -> 1.05 G
// warp instructions, and with five CTAs per SM the kernel is bound by issue slots -- a serial stage costs a full warp
// instruction for four active lanes -- not by the chains' latency: 1.84 -> 1.92 ms on the C5 set; VBZ, with a fifth of
// the frames, gained 3 %.)
constexpr int ZE_BUILD_WARPS = 4;        // warps per CTA
constexpr int ZE_BUILD_SLOTS = 4;        // frames a warp builds at once: the serial stages run on ZE_BUILD_SLOTS lanes together
constexpr size_t ZE_BUILD_SMEM = sizeof(ZencBuildWs) * ZE_BUILD_WARPS * ZE_BUILD_SLOTS;

// A warp takes ZE_BUILD_SLOTS frames per round. The stages that parallelise (histogram scan, entropy bound, bitonic sort,
// length assignment, canonical codes, weight counts) run with all 32 lanes, one frame after the other; the inherently
// serial stages (two-queue Huffman tree, Kraft repair of the length limit, FSE coding of the weights) run ONCE, on lanes
// 0, 8, 16, 24 at the same time, one frame each -- a serial stage costs the warp's issue slots once for four frames.
__global__ void __launch_bounds__(32 * ZE_BUILD_WARPS) k_zenc_build(ZencParams P)
{
    RNSC_DYN_SMEM(smem_raw);
    ZencBuildWs* wsa = (ZencBuildWs*)smem_raw + (threadIdx.x >> 5) * ZE_BUILD_SLOTS;
    const int nf = codec_frames(P.codec);
    const uint32_t lane = threadIdx.x & 31u;
    const size_t nframes = (size_t)P.nchunks * nf;
    const size_t warp_id = (size_t)blockIdx.x * ZE_BUILD_WARPS + (threadIdx.x >> 5), nwarps = (size_t)gridDim.x * ZE_BUILD_WARPS;
    const uint32_t my_slot = lane >> 3;                               // the slot a leader lane (lane % 8 == 0) serves
    const bool leader = (lane & 7u) == 0 && my_slot < (uint32_t)ZE_BUILD_SLOTS;
    for (size_t idx0 = warp_id * ZE_BUILD_SLOTS; idx0 < nframes; idx0 += nwarps * ZE_BUILD_SLOTS) {
        __syncwarp();
        // ---- stage A (all lanes, slot by slot): histogram -> keys, entropy bound, sort ----
        for (int q = 0; q < ZE_BUILD_SLOTS; ++q) {
            ZencBuildWs& w = wsa[q];
            const size_t idx = idx0 + q;
            if (lane == 0) w.live = 0;
            if (idx >= nframes) { __syncwarp(); continue; }
            const uint32_t c = (uint32_t)(idx / nf); const int f = (int)(idx % nf);
            uint64_t a0, b0; ze_frame_range(nf, P.stream_off + 5 * (size_t)c, P.stream_size ? P.stream_size + 5 * (size_t)c : nullptr, f, a0, b0);
            const uint32_t size = (uint32_t)(b0 - a0);
            ZencFrame* fr = P.frames + idx;
            if (lane == 0) { fr->mode = ZM_RAW; fr->rle_byte = 0; fr->log = 0; fr->desc_len = 0; fr->est_bytes = size; }
            __syncwarp();
            if (size < ZE_HUF_MIN || size > ZE_BLOCK_MAX) continue;
            const uint32_t* hist = P.hist + idx * 256;
            uint32_t mine[8], npres = 0, maxs = 0;
            float entropy = 0.f;                                      // order-0 entropy of the frame in bits: no prefix code does better
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t sym = lane * 8 + k, cnt = hist[sym];
                mine[k] = cnt ? ((cnt << 8) | sym) : 0u;
                if (cnt) { ++npres; maxs = sym; entropy += (float)cnt * log2f((float)size / (float)cnt); }
            }
            const uint32_t inc = warp_inclusive_scan(npres);
            const uint32_t m = __shfl_sync(0xffffffffu, inc, 31);
            maxs = __reduce_max_sync(0xffffffffu, maxs);
#pragma unroll
            for (int d = 16; d; d >>= 1) entropy += __shfl_xor_sync(0xffffffffu, entropy, d);
            if (m == 1) {
#pragma unroll
                for (int k = 0; k < 8; ++k) if (mine[k]) { fr->mode = ZM_RLE; fr->rle_byte = mine[k] & 0xFF; }
            }
            if (m <= 1) continue;
            // nothing to gain (nibble pairs, noise): even the entropy bound saves less than ~2 % -- leave the frame raw
            if (entropy * 0.125f + 24.f >= (float)size * 0.98f) continue;
            {
                uint32_t pos = inc - npres;
#pragma unroll
                for (int k = 0; k < 8; ++k) if (mine[k]) w.key[pos++] = mine[k];       // unsorted, compact
            }
            // bitonic sort of the 256-entry key array (absent symbols: 0xFFFFFFFF, they end up behind the m present ones): 36
            // stages of 128 compare-exchanges, four per lane
            for (uint32_t i = m + lane; i < 256; i += 32) w.key[i] = 0xFFFFFFFFu;
            __syncwarp();
            for (uint32_t k = 2; k <= 256; k <<= 1)
                for (uint32_t j = k >> 1; j; j >>= 1) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const uint32_t x = lane + 32 * t;             // pair index 0..127 -> the lower element of the pair
                        const uint32_t i = ((x & ~(j - 1)) << 1) | (x & (j - 1)), l = i | j;
                        const uint32_t ka = w.key[i], kb = w.key[l];
                        const bool up = (i & k) == 0;
                        if ((ka > kb) == up) { w.key[i] = kb; w.key[l] = ka; }
                    }
                    __syncwarp();
                }
            if (lane == 0) { w.m = m; w.maxs = maxs; w.size = size; w.live = 1; }
            __syncwarp();
        }
        __syncwarp();
        // ---- stage B (leaders, all slots at once): Huffman tree by the two-queue method -- leaves in key order, internal
        //      nodes are created in non-decreasing weight ----
        if (leader && wsa[my_slot].live) {
            ZencBuildWs& w = wsa[my_slot];
            const uint32_t m = w.m;
            constexpr uint32_t INF = 0xFFFFFFFFu;
            uint32_t li = 0, qi = 0, qn = 0, lw = w.key[0] >> 8, iw = INF;        // heads of the two queues stay in registers
            while (qn < m - 1) {
                uint32_t sum = 0;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (lw <= iw) { sum += lw; w.par[li++] = (uint8_t)qn; lw = li < m ? w.key[li] >> 8 : INF; }
                    else { sum += iw; w.par[256 + qi++] = (uint8_t)qn; iw = qi < qn ? w.wgt[qi] : INF; }
                }
                w.wgt[qn] = sum;
                if (iw == INF && qi == qn) iw = sum;                  // the internal queue was empty: the new node is its head
                ++qn;
            }
            w.par[256 + m - 2] = 0;                                   // depths in place of the parent indices: the root is the last internal node,
            for (int j = (int)m - 3; j >= 0; --j) w.par[256 + j] = (uint8_t)(w.par[256 + w.par[256 + j]] + 1);   // a parent has a higher index than its child and holds its depth already
        }
        __syncwarp();
        // ---- stage C (all lanes, slot by slot): code lengths per symbol, cut at ZE_LOG_MAX, Kraft sum ----
        for (int q = 0; q < ZE_BUILD_SLOTS; ++q) {
            ZencBuildWs& w = wsa[q];
            if (!w.live) continue;
            const uint32_t m = w.m;
#pragma unroll
            for (int k = 0; k < 8; ++k) w.len[lane * 8 + k] = 0;
            __syncwarp();
            uint32_t kraft = 0, nfull = 0, maxlen = 0;
            for (uint32_t i = lane; i < m; i += 32) {
                uint32_t l = w.par[256 + w.par[i]] + 1u;
                maxlen = max(maxlen, l);
                if (l > ZE_LOG_MAX) l = ZE_LOG_MAX;
                w.len[w.key[i] & 0xFF] = (uint8_t)l; kraft += 1u << (ZE_LOG_MAX - l); nfull += l == ZE_LOG_MAX;
            }
            kraft = __reduce_add_sync(0xffffffffu, kraft); nfull = __reduce_add_sync(0xffffffffu, nfull); maxlen = __reduce_max_sync(0xffffffffu, maxlen);
            if (lane == 0) { w.kraft = maxlen > ZE_LOG_MAX ? kraft : 0u; w.nfull = nfull; }      // kraft 0: no repair needed
            __syncwarp();
        }
        __syncwarp();
        // ---- stage D (leaders): length limit (very skewed histograms) -- repair the Kraft sum: lengthen the rarest symbols that
        //      still can be lengthened until the code fits (lengths do not grow along the sorted order: the first nfull symbols
        //      are at the limit), then give spare code space back to the most frequent ones ----
        if (leader && wsa[my_slot].live && wsa[my_slot].kraft) {
            ZencBuildWs& w = wsa[my_slot];
            const uint32_t m = w.m, full = 1u << ZE_LOG_MAX;
            uint32_t kraft = w.kraft;
            while (kraft > full)
                for (uint32_t i = w.nfull; i < m && kraft > full; ++i) { uint8_t& l = w.len[w.key[i] & 0xFF]; if (l < ZE_LOG_MAX) { kraft -= 1u << (ZE_LOG_MAX - l - 1); ++l; } }
            for (int i = (int)m - 1; i >= 0 && kraft < full; --i) { uint8_t& l = w.len[w.key[i] & 0xFF]; while (l > 1 && kraft + (1u << (ZE_LOG_MAX - l)) <= full) { kraft += 1u << (ZE_LOG_MAX - l); --l; } }
            if (kraft != full) w.live = 0;                            // cannot happen; raw is always valid
        }
        __syncwarp();
        // ---- stage E (all lanes, slot by slot): codes in the order of zstd's decoding table (HUF_readDTableX1): ascending weight
        //      = descending length, symbols in natural order inside a weight; a symbol of weight wt owns 2^(wt-1) consecutive
        //      entries. Per lane 8 consecutive symbols; first entry of each = entries of lighter symbols + entries of equal
        //      weight in front. Then the weights of the tree description and how often each value occurs. ----
        for (int q = 0; q < ZE_BUILD_SLOTS; ++q) {
            ZencBuildWs& w = wsa[q];
            if (!w.live) continue;
            const size_t idx = idx0 + q;
            ZencFrame* fr = P.frames + idx;
            const uint32_t* hist = P.hist + idx * 256;
            const uint32_t maxs = w.maxs;
            uint32_t lens[8], L = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { lens[k] = w.len[lane * 8 + k]; L = max(L, lens[k]); }
            L = __reduce_max_sync(0xffffffffu, L);
            uint32_t cntw[13];                                        // entries of weight wt among my 8 symbols
#pragma unroll
            for (int k = 0; k < 13; ++k) cntw[k] = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) if (lens[k]) { const uint32_t wt = L + 1 - lens[k]; for (int t = 1; t < 13; ++t) if ((uint32_t)t == wt) cntw[t] += 1u << (wt - 1); }
            uint32_t before[13], startw[13];                          // exclusive scan over the lanes, per weight; total per weight
            uint32_t run = 0;
#pragma unroll
            for (int t = 1; t < 13; ++t) { const uint32_t i2 = warp_inclusive_scan(cntw[t]); before[t] = i2 - cntw[t]; startw[t] = run; run += __shfl_sync(0xffffffffu, i2, 31); }
            unsigned long long bits = 0;
            if (lane < 13) w.fse.cnt[lane] = 0;
            __syncwarp();
            uint32_t maxw = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint16_t cw = 0;
                const uint32_t wt = lens[k] ? L + 1 - lens[k] : 0u;
                if (lens[k]) {
                    uint32_t e = 0;
#pragma unroll
                    for (int t = 1; t < 13; ++t) if ((uint32_t)t == wt) { e = startw[t] + before[t]; before[t] += 1u << (wt - 1); }
                    cw = (uint16_t)((lens[k] << 11) | (e >> (wt - 1)));
                    bits += (unsigned long long)hist[lane * 8 + k] * lens[k];
                }
                fr->code[lane * 8 + k] = cw;
                w.wts[lane * 8 + k] = (uint8_t)wt;
                if (lane * 8 + k < maxs) { atomicAdd(&w.fse.cnt[wt], 1u); maxw = max(maxw, wt); }
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, d);
            maxw = __reduce_max_sync(0xffffffffu, maxw);
            if (lane == 0) { w.L = L; w.bits = bits; w.maxw = maxw; }
            __syncwarp();
        }
        __syncwarp();
        // ---- stage F (leaders): tree description = weights of symbols 0 .. maxsym-1 (the last one is implied), FSE-compressed
        //      (left in the scratch, copied out by the warp) or direct ----
        if (leader && wsa[my_slot].live) {
            ZencBuildWs& w = wsa[my_slot];
            ZencFrame* fr = P.frames + idx0 + my_slot;
            const uint32_t maxsym = w.maxs;
            uint32_t dl = ze_fse_weights(w.wts, maxsym, w.maxw, w.fse), have = 1;
            if (dl && dl < 128 && (maxsym > 128 || dl < (maxsym + 1) / 2)) fr->desc[0] = (uint8_t)dl;
            else if (maxsym <= 128) {                                 // direct: 4 bits per weight, first weight in the high nibble
                fr->desc[0] = (uint8_t)(127 + maxsym);
                dl = (maxsym + 1) / 2;
                for (uint32_t i = 0; i < dl; ++i) fr->desc[1 + i] = (uint8_t)((w.wts[2 * i] << 4) | (2 * i + 1 < maxsym ? w.wts[2 * i + 1] : 0));
                dl |= 0x80000000u;                                    // already written
            } else have = 0;                                          // no representation: raw
            w.dl = dl; w.have = have;
        }
        __syncwarp();
        // ---- stage G (all lanes): the description's bytes, the verdict ----
        for (int q = 0; q < ZE_BUILD_SLOTS; ++q) {
            ZencBuildWs& w = wsa[q];
            if (!w.live || !w.have) continue;
            ZencFrame* fr = P.frames + idx0 + q;
            uint32_t dl = w.dl;
            if (!(dl & 0x80000000u)) for (uint32_t i = lane; i < dl; i += 32) fr->desc[1 + i] = (uint8_t)(w.fse.out[i >> 2] >> (8 * (i & 3)));
            dl &= 0x7FFFFFFFu;
            if (lane == 0) {
                const uint32_t est = (uint32_t)((w.bits + 7) >> 3);
                fr->desc_len = 1 + dl; fr->log = w.L; fr->est_bytes = est;
                // worth it only if the literals section beats a raw block (ZSTD_minGain-like margin)
                if (1 + dl + est + 16 + (w.size >> 6) < w.size) fr->mode = ZM_HUF;
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------------------
// 3. emit
// -------------------------------------------------------------------------------------------------------------
// Bytes [0, n) from src (any alignment) to dst (any alignment), all threads of the CTA: 16-byte stores on the destination's
// alignment, the source read as aligned words and funnel-shifted into place (one 128-bit load where it is co-aligned); up to
// 15 bytes at either end one by one. Two blocks per thread and round so that two round trips are in flight. (One 32-bit word
// per thread and round left the gather kernel at 11 % issue-active with 115 long-scoreboard stalls per issue.)
// Reads the aligned source word that holds the last byte (at most 3 bytes past it).
__device__ __forceinline__ V16 ze_load16(const uint8_t* s)
{
    const uint32_t sa = (uint32_t)((uintptr_t)s & 3);
    if (((uintptr_t)s & 15) == 0) return *(const V16*)s;
    const uint32_t* q = (const uint32_t*)(s - sa);
    const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3];
    if (sa == 0) return V16{w0, w1, w2, w3};
    const uint32_t w4 = q[4], sh = 8 * sa;
    return V16{__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh)};
}
__device__ inline void ze_copy(uint8_t* dst, const uint8_t* src, uint32_t n)
{
    const int tid = threadIdx.x;
    const uint32_t head = min(n, (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15));
    if ((uint32_t)tid < head) dst[tid] = src[tid];
    const uint32_t nv = (n - head) >> 4;
    const uint8_t* s = src + head; V16* d = (V16*)(dst + head);
    uint32_t i = tid;
    for (; i + ZE_NT < nv; i += 2 * ZE_NT) {
        const V16 a = ze_load16(s + 16 * (size_t)i), b = ze_load16(s + 16 * (size_t)(i + ZE_NT));
        d[i] = a; d[i + ZE_NT] = b;
    }
    if (i < nv) d[i] = ze_load16(s + 16 * (size_t)i);
    const uint32_t done = head + (nv << 4);
    if (done + (uint32_t)tid < n) dst[done + tid] = src[done + tid];
}

// frame header for a content of `size` bytes (single segment, no checksum, no dictionary: what ZSTD_compress writes)
__device__ inline uint32_t ze_frame_header(uint8_t* o, uint32_t size)
{
    o[0] = 0x28; o[1] = 0xB5; o[2] = 0x2F; o[3] = 0xFD;
    if (size < 256) { o[4] = 0x20; o[5] = (uint8_t)size; return 6; }
    if (size < 65536 + 256) { o[4] = 0x60; const uint32_t v = size - 256; o[5] = (uint8_t)v; o[6] = (uint8_t)(v >> 8); return 7; }
    o[4] = 0xA0; o[5] = (uint8_t)size; o[6] = (uint8_t)(size >> 8); o[7] = (uint8_t)(size >> 16); o[8] = (uint8_t)(size >> 24); return 9;
}
__device__ inline void ze_block_header(uint8_t* o, uint32_t last, uint32_t type, uint32_t size) { const uint32_t v = last | (type << 1) | (size << 3); o[0] = (uint8_t)v; o[1] = (uint8_t)(v >> 8); o[2] = (uint8_t)(v >> 16); }

// Calls f(byte) for the bytes p[b-1], p[b-2], .. p[a] (last first: the order zstd's Huffman streams are written in). The
// middle of the range is read as aligned 32-bit words, at most three bytes at either end one by one; every lane walks the
// same three loops (ranges of a warp have the same length up to one symbol), so the lanes do not serialise on their
// alignment. (16-byte blocks with a predicated body for the ragged ones were tried: the ranges are 25..250 symbols, the
// two bodies ran for almost every block and the kernel was 2.4 x slower.)
template <typename F>
__device__ __forceinline__ void ze_for_bytes_rev(const uint8_t* p, uint32_t a, uint32_t b, F f)
{
    const uint32_t nt = min(b - a, (uint32_t)((uintptr_t)(p + b) & 3u));
#pragma unroll
    for (uint32_t k = 0; k < 3; ++k) if (k < nt) f((uint32_t)p[b - 1 - k]);
    uint32_t i = b - nt;
    uint32_t nw = (i - a) >> 2;
    // four words per round trip: the streams were written by an earlier kernel, so the loads go through the read-only path and
    // are all issued before the first symbol is handled (a load per word in front of its four symbols left the global latency
    // exposed once per word: 30 % of the emit kernel's stall samples)
    for (; nw >= 4; nw -= 4) {
        i -= 16;
        const uint32_t v3 = __ldg((const uint32_t*)(p + i + 12)), v2 = __ldg((const uint32_t*)(p + i + 8)), v1 = __ldg((const uint32_t*)(p + i + 4)), v0 = __ldg((const uint32_t*)(p + i));
        f(v3 >> 24); f((v3 >> 16) & 0xFFu); f((v3 >> 8) & 0xFFu); f(v3 & 0xFFu);
        f(v2 >> 24); f((v2 >> 16) & 0xFFu); f((v2 >> 8) & 0xFFu); f(v2 & 0xFFu);
        f(v1 >> 24); f((v1 >> 16) & 0xFFu); f((v1 >> 8) & 0xFFu); f(v1 & 0xFFu);
        f(v0 >> 24); f((v0 >> 16) & 0xFFu); f((v0 >> 8) & 0xFFu); f(v0 & 0xFFu);
    }
    for (; nw; --nw) { i -= 4; const uint32_t v = __ldg((const uint32_t*)(p + i)); f(v >> 24); f((v >> 16) & 0xFFu); f((v >> 8) & 0xFFu); f(v & 0xFFu); }
    const uint32_t rest = i - a;
#pragma unroll
    for (uint32_t k = 0; k < 3; ++k) if (k < rest) f((uint32_t)p[i - 1 - k]);
}

// Zeroes bytes [0, n) at dst (any alignment), all threads of the CTA.
__device__ inline void ze_zero(uint8_t* dst, uint32_t n)
{
    const int tid = threadIdx.x;
    const uint32_t head = min(n, (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3));
    if ((uint32_t)tid < head) dst[tid] = 0;
    const uint32_t nw = (n - head) >> 2;
    uint32_t* d = (uint32_t*)(dst + head);
    for (uint32_t i = tid; i < nw; i += ZE_NT) d[i] = 0;
    const uint32_t done = head + (nw << 2);
    if (done + (uint32_t)tid < n) dst[done + tid] = 0;
}

struct ZencWs {
    uint32_t warp_tot[ZE_NT / 32];
    uint32_t pos;                // bytes of the chunk's slot written so far
    uint32_t fail;
    uint16_t code[256];
    uint8_t len[256];            // code lengths alone: the first pass adds up lengths only (a byte load, no shift)
};

__global__ void __launch_bounds__(ZE_NT, ZE_EMIT_CTAS) k_zenc_emit(ZencParams P)
{
    __shared__ ZencWs ws;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nf = codec_frames(P.codec);
    for (uint32_t c = blockIdx.x; c < P.nchunks; c += gridDim.x) {
        const uint64_t* so = P.stream_off + 5 * (size_t)c;
        const uint32_t* ssz = P.stream_size ? P.stream_size + 5 * (size_t)c : nullptr;
        uint8_t* slot = P.slots + P.slot_off[c];
        const uint32_t cap = (uint32_t)(P.slot_off[c + 1] - P.slot_off[c]);
        __syncthreads();
        if (tid == 0) { ws.pos = 0; ws.fail = 0; }
        __syncthreads();
        for (int f = 0; f < nf; ++f) {
            uint64_t a, b; ze_frame_range(nf, so, ssz, f, a, b);
            const uint32_t size = (uint32_t)(b - a);
            const uint8_t* src = P.streams + a;
            const ZencFrame* fr = P.frames + (size_t)c * nf + f;
            const uint32_t pre = f + 1 < nf ? 8u : 0u;                 // u64 size prefix of every frame but the last (A.3-A.7)
            const uint32_t base = ws.pos;
            uint32_t mode = fr->mode;
            // ---- Huffman: the frame's 1 or 4 streams at once, a quarter of the CTA each; bit lengths per thread, suffix scan ----
            uint32_t nstr = 1, q = size, sbytes[4] = {0, 0, 0, 0}, tot_bits[4] = {0, 0, 0, 0};
            uint32_t my_bits = 0, my_start = 0, r0 = 0, r1 = 0, my_j = 0;
            if (mode == ZM_HUF) {
                for (int i = tid; i < 256; i += ZE_NT) { const uint16_t cw = fr->code[i]; ws.code[i] = cw; ws.len[i] = (uint8_t)(cw >> 11); }
                __syncthreads();
                nstr = size >= 256 ? 4u : 1u; q = nstr == 4 ? (size + 3) / 4 : size;
                const uint32_t G = ZE_NT / nstr;                        // threads per stream (whole warps)
                my_j = (uint32_t)tid / G;
                const uint32_t t = (uint32_t)tid % G, s0 = my_j * q, s1 = min(size, s0 + q), per = (s1 - s0 + G - 1) / G;
                r0 = min(s1, s0 + t * per); r1 = min(s1, r0 + per);
                uint32_t bits = 0;
                ze_for_bytes_rev(src, r0, r1, [&](uint32_t sym) { bits += ws.len[sym]; });
                my_bits = bits;
                // exclusive SUFFIX sum over the stream's threads: its last symbol is written first (lowest bits)
                const uint32_t inc = warp_inclusive_scan(bits);
                if (lane == 31) ws.warp_tot[warp] = inc;
                __syncthreads();
                uint32_t before = 0;
                const uint32_t wps = G / 32, gw0 = my_j * wps;
#pragma unroll
                for (int k = 0; k < ZE_NT / 32; ++k) {
                    const uint32_t tw = ws.warp_tot[k];
                    tot_bits[(uint32_t)k / wps < 4 ? (uint32_t)k / wps : 3] += tw;
                    if ((uint32_t)k >= gw0 && k < warp) before += tw;
                }
                my_start = tot_bits[my_j] - (before + inc);             // bits of all symbols behind my range
                for (uint32_t j = 0; j < nstr; ++j) {
                    sbytes[j] = (tot_bits[j] + 1 + 7) >> 3;
                    if (nstr == 4 && j < 3 && sbytes[j] > 0xFFFF) mode = ZM_RAW;      // does not fit the jump table
                }
                __syncthreads();
            }
            uint32_t lit_comp = 0, lhs = 0;
            if (mode == ZM_HUF) {
                lit_comp = fr->desc_len + (nstr == 4 ? 6u : 0u) + sbytes[0] + sbytes[1] + sbytes[2] + sbytes[3];
                lhs = (nstr == 1 || (size < 1024 && lit_comp < 1024)) ? 3u : ((size < 16384 && lit_comp < 16384) ? 4u : 5u);
                if (nstr == 1 && lit_comp >= 1024) mode = ZM_RAW;
                if (lhs + lit_comp + 1 >= size) mode = ZM_RAW;          // no gain after all
            }
            // ---- frame size, capacity check (C5.hpp:420-427: "Not enough space in destination buffer") ----
            const uint32_t fhs = size < 256 ? 6u : (size < 65536 + 256 ? 7u : 9u);
            uint32_t body;
            if (mode == ZM_HUF) body = 3 + lhs + lit_comp + 1;
            else if (mode == ZM_RLE) body = 3 + 1;
            else body = size + 3 * max(1u, (size + ZE_BLOCK_MAX - 1) / ZE_BLOCK_MAX);
            const uint32_t flen = fhs + body;
            if (base + pre + flen > cap) { if (tid == 0) ws.fail = 1; __syncthreads(); break; }
            uint8_t* o = slot + base + pre;
            if (tid == 0) {
                if (pre) { const unsigned long long v = flen; for (int i = 0; i < 8; ++i) slot[base + i] = (uint8_t)(v >> (8 * i)); }
                ze_frame_header(o, size);
            }
            uint8_t* blk = o + fhs;
            if (mode == ZM_RLE) { if (tid == 0) { ze_block_header(blk, 1, 1, size); blk[3] = (uint8_t)fr->rle_byte; } }
            else if (mode == ZM_RAW) {
                uint32_t done = 0, off = 0;
                do {
                    const uint32_t bs = min(size - done, ZE_BLOCK_MAX);
                    if (tid == 0) ze_block_header(blk + off, done + bs == size, 0, bs);
                    ze_copy(blk + off + 3, src + done, bs);
                    done += bs; off += 3 + bs;
                } while (done < size);
            } else {
                if (tid == 0) {
                    ze_block_header(blk, 1, 2, lhs + lit_comp + 1);
                    uint8_t* h = blk + 3;
                    if (lhs == 3) { const uint32_t v = 2u | ((nstr == 4 ? 1u : 0u) << 2) | (size << 4) | (lit_comp << 14); h[0] = (uint8_t)v; h[1] = (uint8_t)(v >> 8); h[2] = (uint8_t)(v >> 16); }
                    else if (lhs == 4) { const uint32_t v = 2u | (2u << 2) | (size << 4) | (lit_comp << 18); h[0] = (uint8_t)v; h[1] = (uint8_t)(v >> 8); h[2] = (uint8_t)(v >> 16); h[3] = (uint8_t)(v >> 24); }
                    else { const unsigned long long v = 2ull | (3ull << 2) | ((unsigned long long)size << 4) | ((unsigned long long)lit_comp << 22); for (int i = 0; i < 5; ++i) h[i] = (uint8_t)(v >> (8 * i)); }
                    uint8_t* d = h + lhs;
                    for (uint32_t i = 0; i < fr->desc_len; ++i) d[i] = fr->desc[i];
                    d += fr->desc_len;
                    if (nstr == 4) for (int j = 0; j < 3; ++j) { d[2 * j] = (uint8_t)sbytes[j]; d[2 * j + 1] = (uint8_t)(sbytes[j] >> 8); }
                    blk[3 + lhs + lit_comp] = 0;                          // sequences section: Number_of_Sequences = 0
                }
                // The streams lie back to back behind the jump table. Their bytes are zeroed, then every thread ORs / stores its
                // bits straight into the slot: whole 32-bit words of its own as plain stores, the two ragged words at the ends
                // of its range -- shared with the neighbouring threads or the neighbouring bytes -- with atomic OR.
                uint8_t* sdst = blk + 3 + lhs + fr->desc_len + (nstr == 4 ? 6u : 0u);
                ze_zero(sdst, sbytes[0] + sbytes[1] + sbytes[2] + sbytes[3]);
                __syncthreads();
                uint8_t* mine = sdst;
                for (uint32_t j = 0; j < my_j; ++j) mine += sbytes[j];
                uint32_t* wbase = (uint32_t*)((uintptr_t)mine & ~(uintptr_t)3);
                const uint32_t bit0 = 8u * (uint32_t)((uintptr_t)mine & 3u);
                if ((uint32_t)tid % (ZE_NT / nstr) == 0) { const uint32_t eb = bit0 + tot_bits[my_j]; atomicOr(wbase + (eb >> 5), 1u << (eb & 31)); }      // end mark above the first symbol's code
                if (my_bits) {
                    const uint32_t bit = bit0 + my_start;
                    unsigned long long acc = 0; uint32_t fill = bit & 31, wi = bit >> 5; bool first = true;
                    ze_for_bytes_rev(src, r0, r1, [&](uint32_t sym) {
                        const uint32_t cw = ws.code[sym];
                        acc |= (unsigned long long)(cw & 0x7FFu) << fill; fill += cw >> 11;
                        if (fill >= 32) {
                            if (first) { atomicOr(wbase + wi, (uint32_t)acc); first = false; } else wbase[wi] = (uint32_t)acc;
                            ++wi; acc >>= 32; fill -= 32;
                        }
                    });
                    if (fill) atomicOr(wbase + wi, (uint32_t)acc);
                }
            }
            __syncthreads();
            if (tid == 0) ws.pos = base + pre + flen;
            __syncthreads();
        }
        if (tid == 0) { P.out_size[c] = ws.fail ? 0u : ws.pos; P.status[c] = ws.fail ? RNSC_ERR_DST_TOO_SMALL : RNSC_OK; }
    }
}

// Capacity of every chunk's slot: the reference's output capacity (pgnano/compressor.h:39-45: max(2n+26, 1024); VBZ:
// ZSTD_compressBound(ceil(n/8)+2n), signal_compression.cpp:30-35), rounded up to 16 bytes.
__global__ void k_zenc_slot_sizes(const uint32_t* n, uint32_t nchunks, int codec, uint32_t* sizes)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nchunks) return;
    const unsigned long long k = n[i];
    unsigned long long s;
    if (codec == RNSC_VBZ) { const unsigned long long src = key_len1((uint32_t)k) + 2 * k; s = src + (src >> 8) + (src < (128u << 10) ? (((128u << 10) - src) >> 11) : 0); }
    else { s = 2 * k + 26; if (s < 1024) s = 1024; }
    sizes[i] = (uint32_t)((s + 15) & ~15ull);
}

// Compacts the slots into the caller's buffer: chunk i -> comp[comp_off[i] .. comp_off[i+1]).
__global__ void __launch_bounds__(ZE_NT) k_zenc_gather(const uint8_t* slots, const uint64_t* slot_off, const uint64_t* comp_off, uint32_t nchunks, uint8_t* comp, unsigned long long comp_cap, int32_t* status)
{
    for (uint32_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const unsigned long long a = comp_off[c], b = comp_off[c + 1];
        if (b > comp_cap) { if (threadIdx.x == 0 && status[c] == RNSC_OK) status[c] = RNSC_ERR_DST_TOO_SMALL; continue; }
        ze_copy(comp + a, slots + slot_off[c], (uint32_t)(b - a));
    }
}

}  // namespace rnsc
