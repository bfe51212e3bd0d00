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
//   k_zenc_build      ONE THREAD per frame (all frames of the batch at once, like the planner): everything serial --
//                     sort by count, Huffman tree, length limit, canonical codes in zstd's table order, tree description
//                     with FSE-compressed weights -- so that 10^5 serial chains overlap instead of stalling a CTA each;
//   k_zenc_emit       one CTA per chunk: per frame, bit lengths per thread, a suffix scan (zstd's Huffman streams are
//                     written last symbol first), bit packing through shared memory, headers, container (A.1-A.8: u64
//                     size prefixes), into the chunk's slot; k_zenc_gather then compacts the slots.
#pragma once
#include "block_prims.cuh"
#include "svb_codec.cuh"

namespace rnsc {

constexpr int ZE_NT = 256;                       // threads per CTA of the histogram / emit kernels
constexpr uint32_t ZE_BLOCK_MAX = 128u << 10;    // Block_Maximum_Size
constexpr uint32_t ZE_HUF_MIN = 64;              // shorter frames are not worth a tree description: raw
constexpr uint32_t ZE_LOG_MAX = 11;              // HUF_TABLELOG_DEFAULT of the literals' Huffman codes
constexpr uint32_t ZE_SEG_SMEM = 32u << 10;      // shared-memory bit buffer of one Huffman stream: a quarter of the largest block at < 8 bits per symbol
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
__host__ __device__ inline void ze_frame_range(int nf, const uint64_t* so, int f, uint64_t& a, uint64_t& b)
{
    if (nf == 1) { a = so[0]; b = so[2]; } else { a = so[f]; b = so[f + 1]; }
}

// -------------------------------------------------------------------------------------------------------------
// 1. histograms
// -------------------------------------------------------------------------------------------------------------
struct ZencParams {
    const uint8_t* streams; const uint64_t* stream_off;      // splitter output: stream k of chunk i = [stream_off[5i+k], stream_off[5i+k+1])
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
        for (int f = 0; f < nf; ++f) {
            uint64_t a, b; ze_frame_range(nf, so, f, a, b);
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
// LSB-first bit writer into a small byte buffer (FSE table description and FSE bitstream of the weights)
struct ZeBits {
    uint8_t* p; uint32_t cap; uint32_t pos; unsigned long long acc; uint32_t fill; bool ovf;
    __device__ void init(uint8_t* buf, uint32_t capacity) { p = buf; cap = capacity; pos = 0; acc = 0; fill = 0; ovf = false; }
    __device__ void add(uint32_t v, uint32_t nb) { acc |= (unsigned long long)(v & ((1u << nb) - 1u)) << fill; fill += nb; while (fill >= 8) { if (pos < cap) p[pos] = (uint8_t)acc; else ovf = true; ++pos; acc >>= 8; fill -= 8; } }
    __device__ uint32_t close_bytes() { if (fill) { if (pos < cap) p[pos] = (uint8_t)acc; else ovf = true; ++pos; acc = 0; fill = 0; } return pos; }
};

// FSE-compressed Huffman weights (RFC 8878 4.2.1.1 / HUF_compressWeights): returns the bytes written to out (<= cap) or 0
// when this representation is not usable (fewer than two weights, one distinct value, no gain possible).
__device__ inline uint32_t ze_fse_weights(const uint8_t* w, uint32_t n, uint8_t* out, uint32_t cap)
{
    if (n < 2) return 0;
    uint32_t cnt[13];
    for (int i = 0; i < 13; ++i) cnt[i] = 0;
    uint32_t maxw = 0;
    for (uint32_t i = 0; i < n; ++i) { cnt[w[i]]++; if (w[i] > maxw) maxw = w[i]; }
    uint32_t present = 0, big = 0;
    for (uint32_t s = 0; s <= maxw; ++s) { if (cnt[s]) ++present; if (cnt[s] > cnt[big]) big = s; }
    if (present < 2) return 0;
    const uint32_t tl = n >= 32 ? 6u : 5u, size = 1u << tl;           // any log in 5..6 is valid; the alphabet has <= 13 symbols
    // normalised counts: proportional, every present symbol at least 1, the remainder to the most frequent one
    int norm[13]; int sum = 0;
    for (uint32_t s = 0; s <= maxw; ++s) { norm[s] = cnt[s] ? (int)max(1u, (cnt[s] * size) / n) : 0; sum += norm[s]; }
    norm[big] += (int)size - sum;
    if (norm[big] < 1) return 0;
    // ---- table description (FSE_writeNCount; mirror of fse_read_ncount in zstd_decode.cuh) ----
    ZeBits b; b.init(out, cap);
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
    const uint32_t hdr = b.close_bytes();
    // ---- compression table (FSE_buildCTable) ----
    uint8_t tsym[64]; uint16_t state_tab[64]; uint32_t cumul[14];
    int dnb[13], dfs[13];
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
    ZeBits e; e.init(out + hdr, cap > hdr ? cap - hdr : 0);
    auto init_state = [&](uint32_t sym) -> uint32_t {                // FSE_initCState2: the symbol's state that costs the most bits next time
        const uint32_t nb = (uint32_t)(dnb[sym] + (1 << 15)) >> 16;
        const uint32_t value = (nb << 16) - (uint32_t)dnb[sym];
        return state_tab[(int)(value >> nb) + dfs[sym]];
    };
    auto encode = [&](uint32_t& st, uint32_t sym) { const uint32_t nb = (uint32_t)((int)st + dnb[sym]) >> 16; e.add(st, nb); st = state_tab[(int)(st >> nb) + dfs[sym]]; };
    uint32_t ip = n, s1, s2;
    if (n & 1) { s1 = init_state(w[--ip]); s2 = init_state(w[--ip]); encode(s1, w[--ip]); }
    else { s2 = init_state(w[--ip]); s1 = init_state(w[--ip]); }
    while (ip > 0) { encode(s2, w[--ip]); encode(s1, w[--ip]); }
    e.add(s2, tl); e.add(s1, tl);
    e.add(1, 1);                                                      // end mark
    const uint32_t body = e.close_bytes();
    if (b.ovf || e.ovf) return 0;
    return hdr + body;
}

__global__ void __launch_bounds__(128) k_zenc_build(ZencParams P)
{
    const int nf = codec_frames(P.codec);
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)P.nchunks * nf) return;
    const uint32_t c = (uint32_t)(idx / nf); const int f = (int)(idx % nf);
    uint64_t a, b; ze_frame_range(nf, P.stream_off + 5 * (size_t)c, f, a, b);
    const uint32_t size = (uint32_t)(b - a);
    ZencFrame* fr = P.frames + idx;
    fr->mode = ZM_RAW; fr->rle_byte = 0; fr->log = 0; fr->desc_len = 0; fr->est_bytes = size;
    if (size < ZE_HUF_MIN || size > ZE_BLOCK_MAX) return;
    const uint32_t* hist = P.hist + idx * 256;
    // ---- present symbols, sorted by (count, symbol) ascending: shell sort of packed keys ----
    uint32_t key[256]; uint32_t m = 0, maxsym = 0;
    for (uint32_t s = 0; s < 256; ++s) { const uint32_t k = hist[s]; if (k) { key[m++] = (k << 8) | s; maxsym = s; } }
    if (m == 0) return;
    if (m == 1) { fr->mode = ZM_RLE; fr->rle_byte = key[0] & 0xFF; return; }
    for (uint32_t gap = m >> 1; gap; gap = gap == 2 ? 1 : (gap * 5) / 11) {
        for (uint32_t i = gap; i < m; ++i) { const uint32_t v = key[i]; uint32_t j = i; while (j >= gap && key[j - gap] > v) { key[j] = key[j - gap]; j -= gap; } key[j] = v; }
        if (gap == 1) break;
    }
    // ---- Huffman tree by the two-queue method: leaves in key order, internal nodes are created in non-decreasing weight ----
    uint32_t wgt[256];           // weights of the internal nodes (m - 1 of them)
    uint16_t par[512];           // parent (index among internal nodes) of leaf i (0..m-1) / internal node j (256 + j)
    {
        uint32_t li = 0, qi = 0, qn = 0;
        while (qn < m - 1) {
            uint32_t sum = 0;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const bool leaf = li < m && (qi >= qn || (key[li] >> 8) <= wgt[qi]);
                if (leaf) { sum += key[li] >> 8; par[li++] = (uint16_t)qn; } else { sum += wgt[qi]; par[256 + qi++] = (uint16_t)qn; }
            }
            wgt[qn++] = sum;
        }
    }
    uint8_t depth[256];          // of the internal nodes; the root is the last one
    depth[m - 2] = 0;
    for (int j = (int)m - 3; j >= 0; --j) depth[j] = (uint8_t)(depth[par[256 + j]] + 1);
    uint8_t len[256];            // code length per SYMBOL
    for (uint32_t s = 0; s < 256; ++s) len[s] = 0;
    uint32_t maxlen = 0;
    for (uint32_t i = 0; i < m; ++i) { const uint32_t l = depth[par[i]] + 1u; len[key[i] & 0xFF] = (uint8_t)l; if (l > maxlen) maxlen = l; }
    if (maxlen > ZE_LOG_MAX) {
        // length limit: cut at ZE_LOG_MAX, then repair the Kraft sum -- lengthen the rarest symbols that still can be
        // lengthened until the code fits, then give spare code space back to the most frequent symbols that can use it
        uint32_t kraft = 0;
        for (uint32_t i = 0; i < m; ++i) { uint8_t& l = len[key[i] & 0xFF]; if (l > ZE_LOG_MAX) l = ZE_LOG_MAX; kraft += 1u << (ZE_LOG_MAX - l); }
        const uint32_t full = 1u << ZE_LOG_MAX;
        while (kraft > full) {
            for (uint32_t i = 0; i < m && kraft > full; ++i) { uint8_t& l = len[key[i] & 0xFF]; if (l < ZE_LOG_MAX) { kraft -= 1u << (ZE_LOG_MAX - l - 1); ++l; } }
        }
        for (int i = (int)m - 1; i >= 0 && kraft < full; --i) { uint8_t& l = len[key[i] & 0xFF]; while (l > 1 && kraft + (1u << (ZE_LOG_MAX - l)) <= full) { kraft += 1u << (ZE_LOG_MAX - l); --l; } }
        if (kraft != full) return;                                    // cannot happen; raw is always valid
        maxlen = 0;
        for (uint32_t i = 0; i < m; ++i) maxlen = max(maxlen, (uint32_t)len[key[i] & 0xFF]);
    }
    const uint32_t L = maxlen;
    // ---- codes in the order of zstd's decoding table (HUF_readDTableX1): ascending weight = descending length, symbols
    //      in natural order inside a weight; a symbol of weight w owns 2^(w-1) consecutive entries ----
    uint32_t start[14]; for (int k = 0; k < 14; ++k) start[k] = 0;
    for (uint32_t s = 0; s <= maxsym; ++s) if (len[s]) start[L + 1 - len[s] + 1] += 1u << (L - len[s]);      // entries of weight w, accumulated at w + 1
    for (uint32_t k = 1; k < 14; ++k) start[k] += start[k - 1];      // start[w] = first entry of weight w
    unsigned long long bits = 0;
    for (uint32_t s = 0; s < 256; ++s) {
        uint16_t cw = 0;
        if (s <= maxsym && len[s]) {
            const uint32_t w = L + 1 - len[s];
            const uint32_t e = start[w]; start[w] += 1u << (w - 1);
            cw = (uint16_t)((len[s] << 11) | (e >> (w - 1)));
            bits += (unsigned long long)hist[s] * len[s];
        }
        fr->code[s] = cw;
    }
    // ---- tree description: weights of symbols 0 .. maxsym-1 (the last one is implied) ----
    uint8_t wts[256];
    for (uint32_t s = 0; s < maxsym; ++s) wts[s] = len[s] ? (uint8_t)(L + 1 - len[s]) : 0;
    uint32_t dl = ze_fse_weights(wts, maxsym, fr->desc + 1, 127);
    if (dl && dl < 128 && (maxsym > 128 || dl < (maxsym + 1) / 2)) fr->desc[0] = (uint8_t)dl;
    else if (maxsym <= 128) {                                         // direct: 4 bits per weight, first weight in the high nibble
        fr->desc[0] = (uint8_t)(127 + maxsym);
        dl = (maxsym + 1) / 2;
        for (uint32_t i = 0; i < dl; ++i) fr->desc[1 + i] = (uint8_t)((wts[2 * i] << 4) | (2 * i + 1 < maxsym ? wts[2 * i + 1] : 0));
    } else return;                                                    // no representation: raw
    fr->desc_len = 1 + dl;
    fr->log = L;
    fr->est_bytes = (uint32_t)((bits + 7) >> 3);
    // worth it only if the literals section beats a raw block (ZSTD_minGain-like margin)
    if (fr->desc_len + fr->est_bytes + 16 + (size >> 6) < size) fr->mode = ZM_HUF;
}

// -------------------------------------------------------------------------------------------------------------
// 3. emit
// -------------------------------------------------------------------------------------------------------------
// Bytes [0, n) from src (any alignment) to dst (any alignment), all threads of the CTA. Word-wise on the destination's
// alignment where the run is long enough.
__device__ inline void ze_copy(uint8_t* dst, const uint8_t* src, uint32_t n)
{
    const int tid = threadIdx.x;
    const uint32_t head = min(n, (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3));
    if ((uint32_t)tid < head) dst[tid] = src[tid];
    const uint32_t nw = (n - head) >> 2;
    const uint8_t* s = src + head; uint32_t* d = (uint32_t*)(dst + head);
    const uint32_t sa = (uint32_t)((uintptr_t)s & 3);
    const uint32_t* sw = (const uint32_t*)(s - sa);
    if (sa == 0) for (uint32_t i = tid; i < nw; i += ZE_NT) d[i] = sw[i];
    else for (uint32_t i = tid; i < nw; i += ZE_NT) d[i] = __funnelshift_r(sw[i], sw[i + 1], 8 * sa);
    const uint32_t done = head + (nw << 2);
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

struct ZencWs {
    uint32_t warp_tot[ZE_NT / 32];
    uint32_t seg_bits[4];
    uint32_t pos;                // bytes of the chunk's slot written so far
    uint32_t fail;
    uint16_t code[256];
    uint32_t buf[ZE_SEG_SMEM / 4 + 4];
};

__global__ void __launch_bounds__(ZE_NT) k_zenc_emit(ZencParams P)
{
    __shared__ ZencWs ws;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nf = codec_frames(P.codec);
    for (uint32_t c = blockIdx.x; c < P.nchunks; c += gridDim.x) {
        const uint64_t* so = P.stream_off + 5 * (size_t)c;
        uint8_t* slot = P.slots + P.slot_off[c];
        const uint32_t cap = (uint32_t)(P.slot_off[c + 1] - P.slot_off[c]);
        __syncthreads();
        if (tid == 0) { ws.pos = 0; ws.fail = 0; }
        __syncthreads();
        for (int f = 0; f < nf; ++f) {
            uint64_t a, b; ze_frame_range(nf, so, f, a, b);
            const uint32_t size = (uint32_t)(b - a);
            const uint8_t* src = P.streams + a;
            const ZencFrame* fr = P.frames + (size_t)c * nf + f;
            const uint32_t pre = f + 1 < nf ? 8u : 0u;                 // u64 size prefix of every frame but the last (A.3-A.7)
            const uint32_t base = ws.pos;
            uint32_t mode = fr->mode;
            // ---- Huffman: bit lengths per thread and stream, suffix scans ----
            uint32_t nstr = 1, q = size, sbytes[4] = {0, 0, 0, 0}, my_bits[4] = {0, 0, 0, 0}, my_start[4] = {0, 0, 0, 0};
            uint32_t r0[4], r1[4];                                    // my symbol range in every stream
            if (mode == ZM_HUF) {
                for (int i = tid; i < 256; i += ZE_NT) ws.code[i] = fr->code[i];
                __syncthreads();
                nstr = size >= 256 ? 4u : 1u; q = nstr == 4 ? (size + 3) / 4 : size;
                for (uint32_t j = 0; j < nstr; ++j) {
                    const uint32_t s0 = j * q, s1 = min(size, s0 + q), m = s1 - s0, per = (m + ZE_NT - 1) / ZE_NT;
                    r0[j] = min(s1, s0 + (uint32_t)tid * per); r1[j] = min(s1, r0[j] + per);
                    uint32_t bits = 0;
                    for (uint32_t i = r0[j]; i < r1[j]; ++i) bits += ws.code[src[i]] >> 11;
                    my_bits[j] = bits;
                    // exclusive SUFFIX sum over the threads: the stream's last symbol is written first (lowest bits)
                    const uint32_t inc = warp_inclusive_scan(bits);
                    if (lane == 31) ws.warp_tot[warp] = inc;
                    __syncthreads();
                    uint32_t before = 0, total = 0;
#pragma unroll
                    for (int k = 0; k < ZE_NT / 32; ++k) { const uint32_t t = ws.warp_tot[k]; total += t; if (k < warp) before += t; }
                    my_start[j] = total - (before + inc);              // bits of all symbols behind my range
                    if (tid == 0) ws.seg_bits[j] = total;
                    __syncthreads();
                    sbytes[j] = (ws.seg_bits[j] + 1 + 7) >> 3;
                    if (sbytes[j] > ZE_SEG_SMEM || (nstr == 4 && j < 3 && sbytes[j] > 0xFFFF)) mode = ZM_RAW;      // does not fit the bit buffer / the jump table
                }
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
                uint8_t* sdst = blk + 3 + lhs + fr->desc_len + (nstr == 4 ? 6u : 0u);
                for (uint32_t j = 0; j < nstr; ++j) {
                    const uint32_t nwords = (sbytes[j] + 3) >> 2;
                    for (uint32_t i = tid; i < nwords; i += ZE_NT) ws.buf[i] = 0;
                    __syncthreads();
                    if (tid == 0) atomicOr(&ws.buf[ws.seg_bits[j] >> 5], 1u << (ws.seg_bits[j] & 31));      // end mark above the first symbol's code
                    // my symbols, last first, appended LSB-first from bit my_start: whole words are mine alone and go out as
                    // plain stores, the two ragged words are shared with the neighbours (atomic OR)
                    if (r1[j] > r0[j] && my_bits[j]) {
                        uint32_t bit = my_start[j];
                        unsigned long long acc = 0; uint32_t fill = bit & 31, wi = bit >> 5; bool first = true;
                        for (uint32_t i = r1[j]; i-- > r0[j];) {
                            const uint32_t cw = ws.code[src[i]];
                            acc |= (unsigned long long)(cw & 0x7FFu) << fill; fill += cw >> 11;
                            if (fill >= 32) {
                                if (first) { atomicOr(&ws.buf[wi], (uint32_t)acc); first = false; } else ws.buf[wi] = (uint32_t)acc;
                                ++wi; acc >>= 32; fill -= 32;
                            }
                        }
                        if (fill) atomicOr(&ws.buf[wi], (uint32_t)acc);
                    }
                    __syncthreads();
                    ze_copy(sdst, (const uint8_t*)ws.buf, sbytes[j]);
                    sdst += sbytes[j];
                    __syncthreads();
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
