// api.cu -- host side of librnsc.so: the C ABI of include/rnsc.h over the sm_100a kernels.
//
// Device-resident batcher: a context owns one CUDA stream, grow-only device staging (samples, compressed
// bytes, offset tables, per-CTA scratch) and pinned host staging. Host-buffer entry points stage the batch
// into HBM, run the kernels and copy the result back; rnsc_dev_* entry points take device pointers and
// only enqueue work. The zstd *frame encoder* stage of the compressors runs on host threads through the
// image's libzstd.so.1 (same shared object the reference links), fed by the device splitter kernels --
// the only way to be byte-identical with the reference's ZSTD_compress(.., 1) output.
#ifdef RNSC_EMUL_BUILD
#include "cuda_emul.h"
#endif
#include "kernels.cuh"

#include <dlfcn.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace rnsc;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_OK(expr)                                                                                   \
    do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return fail(RNSC_ERR_NO_DEVICE, "%s: %s", #expr, cudaGetErrorString(e_)); } while (0)

// ---- libzstd (host entropy stage of the encoders) ----
struct Zstd {
    size_t (*compress)(void*, size_t, const void*, size_t, int) = nullptr;
    size_t (*bound)(size_t) = nullptr;
    unsigned (*is_error)(size_t) = nullptr;
    unsigned (*version)(void) = nullptr;
    bool ok = false;
};
Zstd& zstd()
{
    static Zstd z;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!h) return;
        z.compress = (size_t(*)(void*, size_t, const void*, size_t, int))dlsym(h, "ZSTD_compress");
        z.bound = (size_t(*)(size_t))dlsym(h, "ZSTD_compressBound");
        z.is_error = (unsigned (*)(size_t))dlsym(h, "ZSTD_isError");
        z.version = (unsigned (*)(void))dlsym(h, "ZSTD_versionNumber");
        z.ok = z.compress && z.bound && z.is_error;
    });
    return z;
}
// ZSTD_compressBound's published formula, so that sizes do not need libzstd loaded
size_t zstd_bound(size_t n) { return n + (n >> 8) + (n < (128u << 10) ? (((128u << 10) - n) >> 11) : 0); }

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    int reserve(size_t n)
    {
        if (n <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; return -1; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
    void* p = nullptr; size_t cap = 0;
    int reserve(size_t n)
    {
        if (n <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 256;
        if (cudaMallocHost(&p, want) != cudaSuccess) { p = nullptr; return -1; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

constexpr size_t kDecodeSmem = 112 * 1024;   // two CTAs per SM (228 KB per SM, 1 KB reserved per CTA)
// fast kernel size classes: dynamic shared memory per CTA so that 4 / 2 / 1 CTAs fit the 228 KB of an SM (1 KB reserved per CTA)
constexpr size_t kFastSmem[3] = {56 * 1024, 113 * 1024, 227 * 1024};
constexpr size_t kFastWs[3] = {(sizeof(FastWs<256>) + 15) & ~(size_t)15, (sizeof(FastWs<512>) + 15) & ~(size_t)15, (sizeof(FastWs<1024>) + 15) & ~(size_t)15};

// Predefined sequence decoding tables (RFC 8878 3.1.1.3.2.2), entry = (newState << 16) | (nbBits << 8) | symbol,
// laid out LL[64] OF[32] ML[64]; built once on the host with the same spread as fse_build (zstd_decode.cuh).
void host_fse_build(uint32_t* t, const int16_t* norm, int nsym, int tl)
{
    const int size = 1 << tl; int high = size - 1; uint16_t next[64];
    for (int s = 0; s < nsym; ++s) { if (norm[s] == -1) { t[high--] = (uint32_t)s; next[s] = 1; } else next[s] = (uint16_t)norm[s]; }
    const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1; int pos = 0;
    for (int s = 0; s < nsym; ++s)
        for (int i = 0; i < norm[s]; ++i) { t[pos] = (uint32_t)s; pos = (pos + step) & mask; while (pos > high) pos = (pos + step) & mask; }
    for (int u = 0; u < size; ++u) {
        const uint32_t s = t[u] & 0xFF, ns = next[s]++;
        int hb = 0; while ((ns >> (hb + 1)) != 0) ++hb;
        const int nb = tl - hb;
        t[u] = (((ns << nb) - (uint32_t)size) << 16) | ((uint32_t)nb << 8) | s;
    }
}
void host_default_tables(uint32_t* tab /* 160 */)
{
    static const int16_t ll[36] = {4,3,2,2,2,2,2,2,2,2,2,2,2,1,1,1,2,2,2,2,2,2,2,2,2,3,2,1,1,1,1,1,-1,-1,-1,-1};
    static const int16_t ml[53] = {1,4,3,2,2,2,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1,-1,-1};
    static const int16_t of[29] = {1,1,1,1,1,1,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1};
    host_fse_build(tab, ll, 36, 6); host_fse_build(tab + 64, of, 29, 5); host_fse_build(tab + 96, ml, 53, 6);
}

}  // namespace

// Device/pinned staging of one slice of a host-buffer batch (two lanes: copies overlap the kernels)
struct Lane {
    DevBuf comp, samples, off_a, off_b, n, status;
    PinBuf h_off;
    PinBuf h_status;                          // the slice's statuses land here (pinned: a D2H copy into the caller's pageable array would block the issuing thread until the slice's kernels are done)
    size_t st_first = 0, st_count = 0;        // which chunks of the call they belong to (copied out by the host once the slice is through)
    cudaEvent_t ev_in = nullptr, ev_dec = nullptr, ev_out = nullptr, ev_tab = nullptr;
    bool used = false;
};
static uint64_t slice_samples()                   // samples per pipeline slice (default 128 MiB of int16); RNSC_SLICE_SAMPLES overrides (tuning)
{
    static const uint64_t v = [] { const char* e = getenv("RNSC_SLICE_SAMPLES"); const unsigned long long x = e ? strtoull(e, nullptr, 10) : 0; return x ? (uint64_t)x : (64ull << 20); }();
    return v;
}

struct rnsc_ctx {
    int device = 0;
    int sm_count = 1;
    cudaStream_t stream = nullptr;
    int host_threads = 0;
    uint64_t launches = 0;
    std::mutex mu;   // host-buffer entry points serialise on the context's staging
    DevBuf d_samples, d_comp, d_off_a, d_off_b, d_n, d_status, d_sizes, d_streams, d_stream_off, d_scratch, d_counter, d_aux;
    DevBuf d_plans, d_lists, d_deftab, d_prof, d_seqpool, d_origin;
    DevBuf d_slab;                            // rnsc_dev_encode: slab offsets (u64) and sizes (u32) of the one-pass splitter
    DevBuf d_zhist, d_zframes, d_zslots, d_zslot_sz, d_zslot_off, d_zout_sz;   // GPU entropy stage of the encoders (zstd_encode.cuh)
    int entropy_stage = 0;                    // rnsc_compress_batch: 0 = host libzstd level 1 (byte-identical to the reference), 1 = GPU entropy stage
    bool profiling = false;
    uint32_t prof_warp = 0;                   // RNSC_PROF_WARP (tuning): which warp, counted from the end, times the phases
    uint32_t cost_bit = 1, cost_sym = 12, cost_run = 1200;   // planner's cost model of the Huffman stage (plan.cuh); RNSC_COST_BIT / RNSC_COST_SYM / RNSC_COST_RUN override (tuning)
    Lane lane[2];
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};   // the size classes of the fast decoder and the general decoder run concurrently
    cudaStream_t hi = nullptr;                             // high priority: the general decoder must get its SMs before the persistent CTAs fill them (see launch_decode)
    cudaEvent_t ev_plan = nullptr, ev_hi = nullptr, ev_side[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_dec_done = nullptr; bool dec_pending = false;   // end of the previous whole-chunk decode: the next one (on any stream) waits for it, they share the scratch
    PinBuf h_a, h_b;
    PinBuf h_pool_seen;                      // {sequence pool entries the planner asked for, chunks} of the latest finished decode (pinned, written by a D2H copy)
    size_t pool_per_chunk[2] = {0, 0};       // learned pool demand per chunk: [0] separate-stream codecs, [1] LZ-heavy frames
};

namespace {

int use_device(rnsc_ctx* c) { CUDA_OK(cudaSetDevice(c->device)); return 0; }

size_t decode_arena_bytes()
{
    const size_t used = ((sizeof(DecodeWs) + 15) & ~(size_t)15) + ((sizeof(ChunkCtl) + 15) & ~(size_t)15);
    return kDecodeSmem - used;
}

int launch_decode(rnsc_ctx* c, int codec, const uint8_t* d_comp, const uint64_t* d_comp_off, const uint32_t* d_n, const uint64_t* d_sample_off,
                  size_t nchunks, uint32_t max_n, int16_t* d_samples, int32_t* d_status, cudaStream_t st, size_t comp_bytes_hint = 0)
{
    if (nchunks == 0) return RNSC_OK;
    const int grid = (int)std::min<size_t>((size_t)c->sm_count * 2, nchunks);
    const size_t spill = ((size_t)stream_bound(codec, max_n) + 5 * 16 + 255) & ~(size_t)255;
    const size_t stride = ((LIT_SCRATCH + 255) & ~(size_t)255) + spill;
    if (c->d_scratch.reserve(stride * (size_t)c->sm_count * 2)) return fail(RNSC_ERR_NO_DEVICE, "cudaMalloc(scratch %zu) failed", stride * (size_t)c->sm_count * 2);
    if (c->d_counter.reserve(256)) return fail(RNSC_ERR_NO_DEVICE, "cudaMalloc(counter) failed");
    if (c->d_plans.reserve(nchunks * sizeof(ChunkPlan)) || c->d_lists.reserve(4 * nchunks * sizeof(uint32_t) + 16)) return fail(RNSC_ERR_NO_DEVICE, "cudaMalloc(plans) failed");
    // Sequence pool: bump-allocated by the planner, SEQ_WORDS words per LZ sequence. Sized from the data, not from a fixed
    // worst case: frames that hold keys and payload together (VBZ, VBZ0) or the VBZ payload alone (C1) carry
    // ~1 sequence per 30-100 compressed bytes, the separate streams of C2-C5 a handful per chunk. The demand of the previous
    // calls (read back below, a call late) raises the estimate; a batch that still overflows is not an error -- the
    // overflowing chunks take the general decoder, rnsc_decode_class_counts shows them, and the next call has room.
    const bool lz_heavy = codec_frames(codec) == 1 || codec == RNSC_C1;
    size_t per_chunk = lz_heavy ? 1024 : 64;
    if (c->h_pool_seen.p) {
        const volatile unsigned long long* seen = (const volatile unsigned long long*)c->h_pool_seen.p;     // {entries asked for, chunks} of an earlier call
        if (seen[1]) { const size_t d = (size_t)((seen[0] + seen[1] - 1) / seen[1]); c->pool_per_chunk[lz_heavy] = std::max(c->pool_per_chunk[lz_heavy], d + d / 4 + 1); }
    }
    per_chunk = std::max(per_chunk, c->pool_per_chunk[lz_heavy]);
    // A floor of 2^20 entries (32 MiB) whatever the estimate says: the first call of a context on frames of another writer
    // (zstd level 9: thousands of sequences in a C5 frame that level 1 writes without any) has no learned demand yet, and a
    // small batch should not fall to the general decoder for a few KB of pool.
    size_t pool_cap = std::max<size_t>(nchunks * per_chunk, (size_t)1 << 20) + 4096;
    if (comp_bytes_hint) pool_cap = std::min(pool_cap, std::max<size_t>(comp_bytes_hint / 8, nchunks * 16) + 4096);   // a sequence costs more than a byte in any real frame
    if (c->d_seqpool.reserve(pool_cap * SEQ_WORDS * sizeof(uint32_t))) return fail(RNSC_ERR_NO_DEVICE, "cudaMalloc(sequence pool) failed");
    if (!c->h_pool_seen.p) { if (c->h_pool_seen.reserve(64)) return fail(RNSC_ERR_NO_DEVICE, "cudaMallocHost(pool demand) failed"); memset(c->h_pool_seen.p, 0, 64); }
    // counter words: [0..3] list lengths (class S, A, B, general), [4..5] sequence pool cursor (64-bit), [8..11] the work cursors of the four kernels
    uint32_t* cnt = (uint32_t*)c->d_counter.p;
    // Calls on one context share the plans, lists, counters, sequence pool and side streams: a call enqueued on another
    // stream than its predecessor must not start before that one has finished (double-buffered callers).
    if (c->dec_pending) CUDA_OK(cudaStreamWaitEvent(st, c->ev_dec_done, 0));
    CUDA_OK(cudaMemsetAsync(cnt, 0, 64, st));
    // 1. planner: one thread per chunk does every serial step (headers, Huffman tree descriptions, sequences)
    PlanParams pp;
    pp.comp = d_comp; pp.comp_off = d_comp_off; pp.n = d_n; pp.nchunks = (uint32_t)nchunks; pp.codec = codec;
    pp.plans = (ChunkPlan*)c->d_plans.p; pp.status = d_status; pp.lists = (uint32_t*)c->d_lists.p; pp.counts = cnt;
    for (int k = 0; k < 3; ++k) pp.region[k] = (uint32_t)(kFastSmem[k] - kFastWs[k] - unpack2_lut_bytes(codec));
    pp.deftab = (const uint32_t*)c->d_deftab.p;
    pp.cost_bit = c->cost_bit; pp.cost_sym = c->cost_sym; pp.cost_run = c->cost_run;
    pp.has_origin = (lz_heavy || c->pool_per_chunk[0] > 64 || c->d_origin.p) ? 1u : 0u;
    pp.seq_pool = (uint32_t*)c->d_seqpool.p; pp.pool_cap = (uint32_t)std::min<size_t>(pool_cap, 0xFFFFFFFFu); pp.pool_cursor = (unsigned long long*)(cnt + 4);   /* words 4-5, 8-byte aligned */
    RNSC_LAUNCH(k_plan_chunks, (int)((nchunks + PLAN_THREADS - 1) / PLAN_THREADS), PLAN_THREADS, 0, st, pp);
    {   // the pool cursor (entries asked for) and the chunk count travel to the pinned mirror; nobody waits for them. The kernel
        // stores straight into the pinned host words (mapped under unified addressing): a D2H copy on this stream would queue
        // behind whatever the copy engine is doing -- in rnsc_decompress_batch the 128 MiB D2H of the previous slice -- and hold
        // the decode kernels behind it back with it (measured: a bubble of one decode, ~0.4 ms, per slice in the D2H stream)
        unsigned long long* mirror = nullptr;
        CUDA_OK(cudaHostGetDevicePointer((void**)&mirror, c->h_pool_seen.p, 0));
        RNSC_LAUNCH(k_pool_demand, 1, 1, 0, st, (const unsigned long long*)(cnt + 4), (unsigned long long)nchunks, mirror);
    }
    // 2. fast kernel over the three size classes, 3. general decoder over what is left. Fork: class B (large chunks, one
    // CTA per SM) stays on the caller's stream and is enqueued first; the other classes and the general decoder go to side
    // streams and fill SMs as class B CTAs retire. Join before returning to the caller's stream.
    FastParams fp;
    fp.comp = d_comp; fp.comp_off = d_comp_off; fp.n = d_n; fp.sample_off = d_sample_off; fp.samples = d_samples; fp.status = d_status; fp.codec = codec;
    fp.prof_warp = c->prof_warp;
    fp.plans = (const ChunkPlan*)c->d_plans.p; fp.seq_pool = (const uint32_t*)c->d_seqpool.p; fp.nchunks = (uint32_t)nchunks;
    // match resolver scratch in global memory: one origin word per matched byte of a chunk whose origin words do not fit the
    // dead bitstream area of its staged bytes. ~400 MB per context, so it exists only where sequence-heavy frames are
    // expected (LZ-heavy codecs, or a context that has seen such frames); without it the planner routes the chunks that
    // would need it to the general decoder.
    size_t origin_words[3], origin_base[3], origin_total = 0;
    for (int k = 0; k < 3; ++k) { origin_words[k] = kFastSmem[k]; origin_base[k] = origin_total; origin_total += origin_words[k] * (size_t)c->sm_count * (4 >> k); }
    const bool want_origin = lz_heavy || c->pool_per_chunk[0] > 64;
    if (want_origin && c->d_origin.reserve(origin_total * sizeof(uint32_t))) return fail(RNSC_ERR_NO_DEVICE, "cudaMalloc(origin scratch) failed");
    auto set_class = [&](uint32_t cls) { fp.list = (const uint32_t*)c->d_lists.p + (size_t)cls * nchunks; fp.count = cnt + cls; fp.counter = cnt + 8 + cls;
                                         fp.origin = (uint32_t*)c->d_origin.p + origin_base[cls]; fp.origin_stride = origin_words[cls];
                                         fp.prof = c->profiling ? (unsigned long long*)c->d_prof.p + 8 * cls : nullptr;
                                         fp.span = c->profiling ? (unsigned long long*)c->d_prof.p + 24 + 3 * 1024 * cls : nullptr; };
    CUDA_OK(cudaEventRecord(c->ev_plan, st));
    for (int k = 0; k < 2; ++k) CUDA_OK(cudaStreamWaitEvent(c->side[k], c->ev_plan, 0));
    CUDA_OK(cudaStreamWaitEvent(c->hi, c->ev_plan, 0));
    // The general decoder goes FIRST and on the high-priority stream: it usually has a handful of chunks, but every SM is
    // about to be filled with persistent CTAs of the fast kernel that stay until their lists run dry -- launched behind
    // them it would start when everything else is over and add its whole run time to the call.
    DecodeParams p;
    p.comp = d_comp; p.comp_off = d_comp_off; p.n = d_n; p.sample_off = d_sample_off; p.samples = d_samples; p.status = d_status;
    p.nchunks = (uint32_t)nchunks; p.codec = codec; p.scratch = (uint8_t*)c->d_scratch.p; p.scratch_stride = stride;
    p.counter = cnt + 8 + CLS_SLOW; p.arena_bytes = (uint32_t)decode_arena_bytes(); p.spill_bytes = (uint32_t)spill;
    p.list = (const uint32_t*)c->d_lists.p + (size_t)CLS_SLOW * nchunks; p.count = cnt + CLS_SLOW;
    RNSC_LAUNCH(k_decode_chunks, grid, NT, kDecodeSmem, c->hi, p);
    CUDA_OK(cudaEventRecord(c->ev_hi, c->hi));
    set_class(CLS_B);
    RNSC_LAUNCH(k_decode_fast<1024>, (int)std::min<size_t>((size_t)c->sm_count, nchunks), 1024, kFastSmem[2], st, fp);
    set_class(CLS_A);
    RNSC_LAUNCH(k_decode_fast<512>, (int)std::min<size_t>((size_t)c->sm_count * 2, nchunks), 512, kFastSmem[1], c->side[0], fp);
    set_class(CLS_S);
    RNSC_LAUNCH(k_decode_fast<256>, (int)std::min<size_t>((size_t)c->sm_count * 4, nchunks), 256, kFastSmem[0], c->side[1], fp);
    for (int k = 0; k < 2; ++k) { CUDA_OK(cudaEventRecord(c->ev_side[k], c->side[k])); CUDA_OK(cudaStreamWaitEvent(st, c->ev_side[k], 0)); }
    CUDA_OK(cudaStreamWaitEvent(st, c->ev_hi, 0));
    CUDA_OK(cudaEventRecord(c->ev_dec_done, st)); c->dec_pending = true;
    c->launches += 6;
    CUDA_OK(cudaGetLastError());
    return RNSC_OK;
}

bool codec_ok(int codec) { return codec >= RNSC_VBZ && codec <= RNSC_RC; }
bool svb_codec_ok(int codec) { return codec >= RNSC_VBZ && codec <= RNSC_C5; }

template <typename F>
void parallel_for(size_t n, int threads, F&& f)
{
    if (threads <= 1 || n <= 1) { for (size_t i = 0; i < n; ++i) f(i); return; }
    std::atomic<size_t> next{0};
    auto body = [&] { for (;;) { size_t i = next.fetch_add(1); if (i >= n) break; f(i); } };
    std::vector<std::thread> pool;
    const int nt = (int)std::min<size_t>((size_t)threads, n);
    for (int t = 1; t < nt; ++t) pool.emplace_back(body);
    body();
    for (auto& t : pool) t.join();
}

}  // namespace

extern "C" {

const char* rnsc_version(void) { return "rnsc-b200 0.1 (sm_100a)"; }

int rnsc_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char* rnsc_status_string(int s)
{
    switch (s) {
    case RNSC_OK: return "OK";
    case RNSC_ERR_NOT_ZSTD: return "Input data not compressed by zstd";
    case RNSC_ERR_ZSTD_DECODE: return "Input data failed to decompress using zstd";
    case RNSC_ERR_DST_TOO_SMALL: return "Not enough space in destination buffer";
    case RNSC_ERR_REMAINING: return "Remaining data at end of signal buffer";
    case RNSC_ERR_CONTAINER: return "Malformed pgnano container (size prefix past end of chunk)";
    case RNSC_ERR_ZSTD_ENCODE: return "Failed to compress data";
    case RNSC_ERR_RC_RAW: return "Range coder output exceeds raw size (reference raw fallback is unsupported)";
    case RNSC_ERR_BAD_ARG: return "Invalid argument";
    case RNSC_ERR_NO_DEVICE: return "CUDA device unavailable or CUDA error";
    case RNSC_ERR_NO_ZSTD: return "libzstd.so.1 could not be loaded";
    default: return "Unknown status";
    }
}

const char* rnsc_last_error(void) { return g_err.c_str(); }

rnsc_ctx* rnsc_create(int device)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { fail(RNSC_ERR_NO_DEVICE, "no CUDA device: librnsc has no CPU fallback"); return nullptr; }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= count) { fail(RNSC_ERR_BAD_ARG, "device %d out of range (%d devices)", device, count); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { fail(RNSC_ERR_NO_DEVICE, "cudaSetDevice(%d) failed", device); return nullptr; }
    rnsc_ctx* c = new rnsc_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { rnsc_destroy(c); fail(RNSC_ERR_NO_DEVICE, "cudaGetDeviceProperties failed"); return nullptr; }
    c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreate(&c->stream) != cudaSuccess) { rnsc_destroy(c); fail(RNSC_ERR_NO_DEVICE, "cudaStreamCreate failed"); return nullptr; }
    bool ok = cudaEventCreateWithFlags(&c->ev_plan, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&c->ev_dec_done, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; k < 3 && ok; ++k) ok = cudaStreamCreateWithFlags(&c->side[k], cudaStreamNonBlocking) == cudaSuccess && cudaEventCreateWithFlags(&c->ev_side[k], cudaEventDisableTiming) == cudaSuccess;
    {
        int least = 0, greatest = 0;
        ok = ok && cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess
                && cudaStreamCreateWithPriority(&c->hi, cudaStreamNonBlocking, greatest) == cudaSuccess && cudaEventCreateWithFlags(&c->ev_hi, cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking) == cudaSuccess && cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking) == cudaSuccess;
    for (int k = 0; k < 2 && ok; ++k) ok = cudaEventCreateWithFlags(&c->lane[k].ev_in, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&c->lane[k].ev_dec, cudaEventDisableTiming) == cudaSuccess &&
                                           cudaEventCreateWithFlags(&c->lane[k].ev_out, cudaEventDisableTiming) == cudaSuccess &&
                                           cudaEventCreateWithFlags(&c->lane[k].ev_tab, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { rnsc_destroy(c); fail(RNSC_ERR_NO_DEVICE, "cudaStreamCreate/cudaEventCreate failed"); return nullptr; }
    cudaError_t e1 = cudaFuncSetAttribute(k_decode_chunks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDecodeSmem);
    { cudaError_t e0 = cudaFuncSetAttribute(k_zenc_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZE_BUILD_SMEM); if (e1 == cudaSuccess) e1 = e0; }
    cudaError_t e2 = cudaSuccess;
    cudaError_t e3 = cudaFuncSetAttribute(k_decode_fast<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFastSmem[2]);
    if (e1 == cudaSuccess) e1 = e3;
    e3 = cudaFuncSetAttribute(k_decode_fast<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFastSmem[1]);
    if (e1 == cudaSuccess) e1 = e3;
    e3 = cudaFuncSetAttribute(k_decode_fast<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFastSmem[0]);
    if (e1 == cudaSuccess) e1 = e3;
    // one shared-memory carveout for every decode kernel: kernels that ask for different L1/shared splits cannot share an SM,
    // and an SM has to drain before it switches
    for (const void* f : {(const void*)k_decode_fast<1024>, (const void*)k_decode_fast<512>, (const void*)k_decode_fast<256>, (const void*)k_decode_chunks, (const void*)k_plan_chunks}) {
        e3 = cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e1 == cudaSuccess) e1 = e3;
    }
    if (e1 != cudaSuccess || e2 != cudaSuccess) { rnsc_destroy(c); fail(RNSC_ERR_NO_DEVICE, "cudaFuncSetAttribute: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2)); return nullptr; }
    uint32_t deftab[160];
    host_default_tables(deftab);
    if (c->d_deftab.reserve(sizeof deftab) || cudaMemcpy(c->d_deftab.p, deftab, sizeof deftab, cudaMemcpyHostToDevice) != cudaSuccess) {
        rnsc_destroy(c); fail(RNSC_ERR_NO_DEVICE, "upload of the predefined FSE tables failed"); return nullptr;
    }
    unsigned hc = std::thread::hardware_concurrency();
    c->host_threads = hc ? (int)hc : 1;
    if (const char* e = getenv("RNSC_PROF_WARP")) c->prof_warp = (uint32_t)strtoul(e, nullptr, 10);
    if (const char* e = getenv("RNSC_COST_BIT")) c->cost_bit = (uint32_t)strtoul(e, nullptr, 10);
    if (const char* e = getenv("RNSC_COST_SYM")) c->cost_sym = (uint32_t)strtoul(e, nullptr, 10);
    if (const char* e = getenv("RNSC_COST_RUN")) c->cost_run = (uint32_t)strtoul(e, nullptr, 10);
    if (c->cost_bit + c->cost_sym == 0) c->cost_bit = 1;
    return c;
}

void rnsc_destroy(rnsc_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);      // a half-built context (rnsc_create failure) has null handles
    for (DevBuf* b : {&c->d_samples, &c->d_comp, &c->d_off_a, &c->d_off_b, &c->d_n, &c->d_status, &c->d_sizes, &c->d_streams, &c->d_stream_off, &c->d_scratch, &c->d_counter, &c->d_aux, &c->d_plans, &c->d_lists, &c->d_deftab, &c->d_prof, &c->d_seqpool, &c->d_origin, &c->d_zhist, &c->d_zframes, &c->d_zslots, &c->d_zslot_sz, &c->d_zslot_off, &c->d_zout_sz, &c->d_slab}) b->release();
    c->h_a.release(); c->h_b.release(); c->h_pool_seen.release();
    if (c->hi) { cudaStreamSynchronize(c->hi); cudaStreamDestroy(c->hi); } if (c->ev_hi) cudaEventDestroy(c->ev_hi);
    for (int i = 0; i < 3; ++i) { if (c->side[i]) { cudaStreamSynchronize(c->side[i]); cudaStreamDestroy(c->side[i]); } if (c->ev_side[i]) cudaEventDestroy(c->ev_side[i]); }
    if (c->ev_plan) cudaEventDestroy(c->ev_plan);
    if (c->ev_dec_done) cudaEventDestroy(c->ev_dec_done);
    for (int k = 0; k < 2; ++k) {
        Lane& L = c->lane[k];
        for (DevBuf* b : {&L.comp, &L.samples, &L.off_a, &L.off_b, &L.n, &L.status}) b->release();
        L.h_off.release(); L.h_status.release();
        for (cudaEvent_t e : {L.ev_in, L.ev_dec, L.ev_out, L.ev_tab}) if (e) cudaEventDestroy(e);
    }
    if (c->h2d) { cudaStreamSynchronize(c->h2d); cudaStreamDestroy(c->h2d); }
    if (c->d2h) { cudaStreamSynchronize(c->d2h); cudaStreamDestroy(c->d2h); }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

void rnsc_set_host_threads(rnsc_ctx* c, int n) { if (c && n > 0) c->host_threads = n; }
int rnsc_set_entropy_stage(rnsc_ctx* c, int mode) { if (!c || mode < 0 || mode > 1) return fail(RNSC_ERR_BAD_ARG, "rnsc_set_entropy_stage: mode must be RNSC_ENTROPY_LIBZSTD or RNSC_ENTROPY_GPU"); c->entropy_stage = mode; return RNSC_OK; }
int rnsc_get_entropy_stage(const rnsc_ctx* c) { return c ? c->entropy_stage : -1; }
uint64_t rnsc_launch_count(const rnsc_ctx* c) { return c ? c->launches : 0; }

size_t rnsc_max_compressed_size(int codec, size_t n)
{
    if (codec == RNSC_VBZ) return zstd_bound((size_t)key_len1((uint32_t)n) + 2 * n);
    const size_t s = 2 * n + 10 + 16;
    return s > 1024 ? s : 1024;
}
size_t rnsc_max_stream_bytes(int codec, size_t n) { return (size_t)stream_bound(codec, n); }
int rnsc_stream_count(int codec) { return svb_codec_ok(codec) ? codec_streams(codec) : 0; }

// ---------------------------------------------------------------------------------------------------
// device-resident entry points
// ---------------------------------------------------------------------------------------------------
int rnsc_dev_decode(rnsc_ctx* c, int codec, const uint8_t* d_comp, const uint64_t* d_comp_off, const uint32_t* d_n, const uint64_t* d_sample_off,
                    size_t nchunks, uint32_t max_n, int16_t* d_samples, int32_t* d_status, void* stream)
{
    if (!c || !svb_codec_ok(codec) || (nchunks && (!d_comp || !d_comp_off || !d_n || !d_sample_off || !d_samples || !d_status))) return fail(RNSC_ERR_BAD_ARG, "rnsc_dev_decode: bad argument");
    if (use_device(c)) return RNSC_ERR_NO_DEVICE;
    return launch_decode(c, codec, d_comp, d_comp_off, d_n, d_sample_off, nchunks, max_n, d_samples, d_status, stream ? (cudaStream_t)stream : c->stream);
}

int rnsc_decode_class_counts(rnsc_ctx* c, uint32_t out[4], void* stream)
{
    if (!c || !out) return fail(RNSC_ERR_BAD_ARG, "rnsc_decode_class_counts: bad argument");
    out[0] = out[1] = out[2] = out[3] = 0;
    if (!c->d_counter.p) return RNSC_OK;
    if (use_device(c)) return RNSC_ERR_NO_DEVICE;
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    CUDA_OK(cudaMemcpyAsync(out, c->d_counter.p, 16, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    return RNSC_OK;
}

int rnsc_set_profiling(rnsc_ctx* c, int on)
{
    if (!c) return fail(RNSC_ERR_BAD_ARG, "rnsc_set_profiling: bad argument");
    if (use_device(c)) return RNSC_ERR_NO_DEVICE;
    if (on) { if (c->d_prof.reserve((24 + RNSC_SPAN_WORDS) * 8)) return fail(RNSC_ERR_NO_DEVICE, "cudaMalloc(prof) failed"); CUDA_OK(cudaMemset(c->d_prof.p, 0, (24 + RNSC_SPAN_WORDS) * 8)); }
    c->profiling = on != 0;
    return RNSC_OK;
}

int rnsc_get_phase_cycles(rnsc_ctx* c, uint64_t out[24])
{
    if (!c || !out) return fail(RNSC_ERR_BAD_ARG, "rnsc_get_phase_cycles: bad argument");
    memset(out, 0, 24 * 8);
    if (!c->d_prof.p) return RNSC_OK;
    if (use_device(c)) return RNSC_ERR_NO_DEVICE;
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpy(out, c->d_prof.p, 24 * 8, cudaMemcpyDeviceToHost));
    return RNSC_OK;
}

int rnsc_get_cta_spans(rnsc_ctx* c, uint64_t out[RNSC_SPAN_WORDS])
{
    if (!c || !out) return fail(RNSC_ERR_BAD_ARG, "rnsc_get_cta_spans: bad argument");
    memset(out, 0, RNSC_SPAN_WORDS * 8);
    if (!c->d_prof.p) return RNSC_OK;
    if (use_device(c)) return RNSC_ERR_NO_DEVICE;
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpy(out, (const uint64_t*)c->d_prof.p + 24, RNSC_SPAN_WORDS * 8, cudaMemcpyDeviceToHost));
    return RNSC_OK;
}

int rnsc_dev_exclusive_scan_u32(rnsc_ctx* c, const uint32_t* d_in, size_t count, uint64_t* d_out, void* stream)
{
    if (!c || !d_out || (count && !d_in)) return fail(RNSC_ERR_BAD_ARG, "rnsc_dev_exclusive_scan_u32: bad argument");
    if (use_device(c)) return RNSC_ERR_NO_DEVICE;
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    RNSC_LAUNCH(k_exclusive_scan_u32, 1, NT, 0, st, d_in, (unsigned long long)count, (unsigned long long*)d_out);
    c->launches += 1;
    CUDA_OK(cudaGetLastError());
    return RNSC_OK;
}

// sizes -> device prefix scan -> scatter. check_cap: fetch the total (one stream synchronisation) and compare it with the
// caller's capacity; the internal callers size the buffer by the bound and skip it.
static int launch_svb_encode(rnsc_ctx* c, int codec, const int16_t* d_samples, const uint64_t* d_sample_off, const uint32_t* d_n, size_t nchunks,
                             uint8_t* d_streams, size_t streams_cap, uint64_t* d_stream_off, cudaStream_t st, bool check_cap)
{
    if (c->d_sizes.reserve(5 * nchunks * sizeof(uint32_t) + 16)) return fail(RNSC_ERR_NO_DEVICE, "cudaMalloc(sizes) failed");
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)1 << 20, nchunks));   // one CTA per chunk: the hardware scheduler balances the sizes
    if (nchunks) {
#define RNSC_SIZES2(C) case C: RNSC_LAUNCH(k_svb_sizes2<C>, grid, P2_NT, 0, st, d_samples, d_sample_off, d_n, (uint32_t)nchunks, (uint32_t*)c->d_sizes.p); break
        switch (codec) { RNSC_SIZES2(RNSC_VBZ); RNSC_SIZES2(RNSC_VBZ0); RNSC_SIZES2(RNSC_C1); RNSC_SIZES2(RNSC_C2); RNSC_SIZES2(RNSC_C3); RNSC_SIZES2(RNSC_C4); default: RNSC_SIZES2(RNSC_C5); }
#undef RNSC_SIZES2
        c->launches += 1;
    }
    RNSC_LAUNCH(k_exclusive_scan_u32, 1, NT, 0, st, (const uint32_t*)c->d_sizes.p, (unsigned long long)(5 * nchunks), (unsigned long long*)d_stream_off);
    c->launches += 1;
    if (check_cap) {
        uint64_t total = 0;
        CUDA_OK(cudaMemcpyAsync(&total, d_stream_off + 5 * nchunks, sizeof total, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        if (total > streams_cap) return fail(RNSC_ERR_DST_TOO_SMALL, "stream buffer too small: need %llu bytes, have %zu", (unsigned long long)total, streams_cap);
    }
    if (nchunks) {
#define RNSC_PACK2(C) case C: RNSC_LAUNCH(k_svb_pack2<C>, grid, P2_NT, (size_t)P2_STAGE_BYTES, st, d_samples, d_sample_off, d_n, (uint32_t)nchunks, d_streams, (const uint64_t*)d_stream_off); break
        switch (codec) { RNSC_PACK2(RNSC_VBZ); RNSC_PACK2(RNSC_VBZ0); RNSC_PACK2(RNSC_C1); RNSC_PACK2(RNSC_C2); RNSC_PACK2(RNSC_C3); RNSC_PACK2(RNSC_C4); default: RNSC_PACK2(RNSC_C5); }
#undef RNSC_PACK2
        c->launches += 1;
    }
    CUDA_OK(cudaGetLastError());
    return RNSC_OK;
}

int rnsc_dev_svb_encode(rnsc_ctx* c, int codec, const int16_t* d_samples, const uint64_t* d_sample_off, const uint32_t* d_n, size_t nchunks,
                        uint8_t* d_streams, size_t streams_cap, uint64_t* d_stream_off, void* stream)
{
    if (!c || !svb_codec_ok(codec) || !d_stream_off || (nchunks && (!d_samples || !d_sample_off || !d_n || !d_streams))) return fail(RNSC_ERR_BAD_ARG, "rnsc_dev_svb_encode: bad argument");
    if (use_device(c)) return RNSC_ERR_NO_DEVICE;
    return launch_svb_encode(c, codec, d_samples, d_sample_off, d_n, nchunks, d_streams, streams_cap, d_stream_off, stream ? (cudaStream_t)stream : c->stream, true);
}

// One-pass splitter into per-chunk slabs (kernels.cuh: k_svb_pack_slab): slab sizes by formula, one scan, one kernel over the samples.
static int launch_svb_encode_slabs(rnsc_ctx* c, int codec, const int16_t* d_samples, const uint64_t* d_sample_off, const uint32_t* d_n, size_t nchunks,
                                   uint8_t* d_streams, uint64_t* d_stream_start, uint32_t* d_stream_size, cudaStream_t st)
{
    if (c->d_slab.reserve(nchunks * 4 + 16 + (nchunks + 1) * 8)) return fail(RNSC_ERR_NO_DEVICE, "cudaMalloc(slab table) failed");
    uint64_t* slab_off = (uint64_t*)c->d_slab.p; uint32_t* slab_sz = (uint32_t*)((uint8_t*)c->d_slab.p + (nchunks + 1) * 8);
    RNSC_LAUNCH(k_slab_sizes, (int)((nchunks + 255) / 256), 256, 0, st, d_n, (uint32_t)nchunks, codec, slab_sz);
    RNSC_LAUNCH(k_exclusive_scan_u32, 1, NT, 0, st, (const uint32_t*)slab_sz, (unsigned long long)nchunks, (unsigned long long*)slab_off);
    const int grid = (int)std::min<size_t>((size_t)1 << 30, nchunks);
#define RNSC_SLAB(C) case C: RNSC_LAUNCH(k_svb_pack_slab<C>, grid, P2_NT, (size_t)P2_STAGE_BYTES, st, d_samples, d_sample_off, d_n, (uint32_t)nchunks, d_streams, (const uint64_t*)slab_off, (unsigned long long*)d_stream_start, d_stream_size); break
    switch (codec) { RNSC_SLAB(RNSC_VBZ); RNSC_SLAB(RNSC_VBZ0); RNSC_SLAB(RNSC_C1); RNSC_SLAB(RNSC_C2); RNSC_SLAB(RNSC_C3); RNSC_SLAB(RNSC_C4); default: RNSC_SLAB(RNSC_C5); }
#undef RNSC_SLAB
    c->launches += 3;
    CUDA_OK(cudaGetLastError());
    return RNSC_OK;
}

size_t rnsc_svb_slab_bytes(int codec, uint32_t n)
{
    if (!svb_codec_ok(codec)) return 0;
    uint32_t rel[6]; slab_layout(codec, n, rel);
    return rel[5];
}

int rnsc_dev_svb_encode_slabs(rnsc_ctx* c, int codec, const int16_t* d_samples, const uint64_t* d_sample_off, const uint32_t* d_n, size_t nchunks, size_t total_samples,
                              uint8_t* d_streams, size_t streams_cap, uint64_t* d_stream_start, uint32_t* d_stream_size, void* stream)
{
    if (!c || !svb_codec_ok(codec) || (nchunks && (!d_samples || !d_sample_off || !d_n || !d_streams || !d_stream_start || !d_stream_size))) return fail(RNSC_ERR_BAD_ARG, "rnsc_dev_svb_encode_slabs: bad argument");
    if (use_device(c)) return RNSC_ERR_NO_DEVICE;
    if (!nchunks) return RNSC_OK;
    // every slab is at most keys (<= n/4 + 1) + 3.5 n + 5 x 16 bytes of padding + 15 of rounding
    if (streams_cap < 4 * total_samples + 128 * nchunks + 64) return fail(RNSC_ERR_DST_TOO_SMALL, "stream buffer too small: %zu bytes for the slabs of %zu samples in %zu chunks (need 4 x samples + 128 x chunks + 64)", streams_cap, total_samples, nchunks);
    return launch_svb_encode_slabs(c, codec, d_samples, d_sample_off, d_n, nchunks, d_streams, d_stream_start, d_stream_size, stream ? (cudaStream_t)stream : c->stream);
}

// Whole-chunk encode on the device: splitter, then the GPU entropy stage (zstd_encode.cuh): valid zstd frames, decodable by
// libzstd / the reference / this library, NOT byte-identical to libzstd's level-1 output.
int rnsc_dev_encode(rnsc_ctx* c, int codec, const int16_t* d_samples, const uint64_t* d_sample_off, const uint32_t* d_n, size_t nchunks, size_t total_samples,
                    uint8_t* d_comp, size_t comp_cap, uint64_t* d_comp_off, int32_t* d_status, void* stream)
{
    if (!c || !svb_codec_ok(codec) || !d_comp_off || (nchunks && (!d_samples || !d_sample_off || !d_n || !d_comp || !d_status))) return fail(RNSC_ERR_BAD_ARG, "rnsc_dev_encode: bad argument");
    if (use_device(c)) return RNSC_ERR_NO_DEVICE;
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    if (!nchunks) { CUDA_OK(cudaMemsetAsync(d_comp_off, 0, 8, st)); return RNSC_OK; }
    const int nf = codec_frames(codec);
    const size_t streams_cap = 4 * total_samples + 128 * nchunks + 64;                                        // >= the sum of the chunks' slabs (slab_layout: keys + every stream at its upper bound)
    const size_t slots_cap = 2 * total_samples + total_samples / 4 + 1104 * nchunks + 64;                     // sum of the reference's output capacities, 16-byte rounded
    if (c->d_streams.reserve(streams_cap + 64) || c->d_stream_off.reserve((5 * nchunks + 1) * 8) || c->d_sizes.reserve(5 * nchunks * sizeof(uint32_t) + 16) ||
        c->d_zhist.reserve(nchunks * nf * 256 * sizeof(uint32_t)) ||
        c->d_zframes.reserve(nchunks * nf * sizeof(ZencFrame)) || c->d_zslots.reserve(slots_cap + 64) || c->d_zslot_sz.reserve(nchunks * 4 + 16) ||
        c->d_zslot_off.reserve((nchunks + 1) * 8) || c->d_zout_sz.reserve(nchunks * 4 + 16))
        return fail(RNSC_ERR_NO_DEVICE, "cudaMalloc(encoder scratch) failed");
    // splitter into per-chunk slabs: one pass over the samples (no size pass, no scan of the stream sizes)
    { const int r = launch_svb_encode_slabs(c, codec, d_samples, d_sample_off, d_n, nchunks, (uint8_t*)c->d_streams.p, (uint64_t*)c->d_stream_off.p, (uint32_t*)c->d_sizes.p, st); if (r) return r; }
    RNSC_LAUNCH(k_zenc_slot_sizes, (int)((nchunks + 255) / 256), 256, 0, st, d_n, (uint32_t)nchunks, codec, (uint32_t*)c->d_zslot_sz.p);
    RNSC_LAUNCH(k_exclusive_scan_u32, 1, NT, 0, st, (const uint32_t*)c->d_zslot_sz.p, (unsigned long long)nchunks, (unsigned long long*)c->d_zslot_off.p);
    ZencParams zp;
    zp.streams = (const uint8_t*)c->d_streams.p; zp.stream_off = (const uint64_t*)c->d_stream_off.p; zp.n = d_n; zp.nchunks = (uint32_t)nchunks; zp.codec = codec;
    zp.stream_size = (const uint32_t*)c->d_sizes.p;
    zp.hist = (uint32_t*)c->d_zhist.p; zp.frames = (ZencFrame*)c->d_zframes.p; zp.slots = (uint8_t*)c->d_zslots.p; zp.slot_off = (const uint64_t*)c->d_zslot_off.p;
    zp.out_size = (uint32_t*)c->d_zout_sz.p; zp.status = d_status;
    // one CTA per chunk for the per-chunk kernels: the chunks are ragged (1 .. 102400 samples) and a fixed stride over a
    // persistent grid left the CTAs with the most samples running alone at the end (emit: 50 % warps active of 75 % possible)
#ifdef RNSC_ZE_STATIC_GRID
    const int wide = (int)std::min<size_t>(nchunks, (size_t)c->sm_count * 8), emit_grid = (int)std::min<size_t>(nchunks, (size_t)c->sm_count * 6);
#else
    const int wide = (int)nchunks, emit_grid = (int)nchunks;
#endif
    RNSC_LAUNCH(k_zenc_histogram, wide, ZE_NT, 0, st, zp);
    // one round (ZE_BUILD_SLOTS consecutive frames) per warp, dealt out by the block scheduler: a round costs anything between
    // nothing (raw frames) and four 256-symbol trees (a persistent grid with a fixed stride: 2.5 % slower on the 20k-read C5 set.
    // Tried: rounds of one frame KIND, so that the four serial lanes are all busy with trees of one size -- no gain)
    RNSC_LAUNCH(k_zenc_build, (int)((nchunks * nf + ZE_BUILD_WARPS * ZE_BUILD_SLOTS - 1) / (ZE_BUILD_WARPS * ZE_BUILD_SLOTS)), 32 * ZE_BUILD_WARPS, ZE_BUILD_SMEM, st, zp);
    RNSC_LAUNCH(k_zenc_emit, emit_grid, ZE_NT, 0, st, zp);
    RNSC_LAUNCH(k_exclusive_scan_u32, 1, NT, 0, st, (const uint32_t*)c->d_zout_sz.p, (unsigned long long)nchunks, (unsigned long long*)d_comp_off);
    RNSC_LAUNCH(k_zenc_gather, wide, ZE_NT, 0, st, (const uint8_t*)c->d_zslots.p, (const uint64_t*)c->d_zslot_off.p, (const uint64_t*)d_comp_off, (uint32_t)nchunks, d_comp,
                (unsigned long long)comp_cap, d_status);
    c->launches += 7;
    CUDA_OK(cudaGetLastError());
    return RNSC_OK;
}

int rnsc_dev_svb_decode(rnsc_ctx* c, int codec, const uint8_t* d_streams, const uint64_t* d_stream_off, const uint32_t* d_n, const uint64_t* d_sample_off,
                        size_t nchunks, int16_t* d_samples, int32_t* d_status, void* stream)
{
    if (!c || !svb_codec_ok(codec) || (nchunks && (!d_streams || !d_stream_off || !d_n || !d_sample_off || !d_samples || !d_status))) return fail(RNSC_ERR_BAD_ARG, "rnsc_dev_svb_decode: bad argument");
    if (use_device(c)) return RNSC_ERR_NO_DEVICE;
    if (!nchunks) return RNSC_OK;
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const int grid = (int)nchunks;                                        // one CTA per chunk: ragged sizes, balanced by the block scheduler
    RNSC_LAUNCH(k_svb_unpack, grid, NT, 0, st, codec, d_streams, d_stream_off, d_n, d_sample_off, (uint32_t)nchunks, d_samples, d_status);
    c->launches += 1;
    CUDA_OK(cudaGetLastError());
    return RNSC_OK;
}

int rnsc_dev_rc_encode(rnsc_ctx* c, const int16_t* d_samples, const uint64_t* d_sample_off, const uint32_t* d_n, size_t nchunks, uint8_t pore_type,
                       uint8_t* d_comp, const uint64_t* d_slot_off, uint32_t* d_comp_size, int32_t* d_status, void* stream)
{
    if (!c || (nchunks && (!d_samples || !d_sample_off || !d_n || !d_comp || !d_slot_off || !d_comp_size || !d_status))) return fail(RNSC_ERR_BAD_ARG, "rnsc_dev_rc_encode: bad argument");
    if (use_device(c)) return RNSC_ERR_NO_DEVICE;
    if (!nchunks) return RNSC_OK;
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const int grid = (int)((nchunks + RC_WARPS - 1) / RC_WARPS);      // one warp per chunk, the block scheduler deals the ragged chunks out
    RNSC_LAUNCH(k_rc_encode, grid, 32 * RC_WARPS, 0, st, d_samples, d_sample_off, d_n, (uint32_t)nchunks, pore_type, d_comp, d_slot_off, d_comp_size, d_status);
    c->launches += 1;
    CUDA_OK(cudaGetLastError());
    return RNSC_OK;
}

int rnsc_dev_rc_decode(rnsc_ctx* c, const uint8_t* d_comp, const uint64_t* d_comp_off, const uint32_t* d_n, const uint64_t* d_sample_off, size_t nchunks,
                       int16_t* d_samples, int32_t* d_status, void* stream)
{
    if (!c || (nchunks && (!d_comp || !d_comp_off || !d_n || !d_sample_off || !d_samples || !d_status))) return fail(RNSC_ERR_BAD_ARG, "rnsc_dev_rc_decode: bad argument");
    if (use_device(c)) return RNSC_ERR_NO_DEVICE;
    if (!nchunks) return RNSC_OK;
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const int grid = (int)((nchunks + RC_WARPS - 1) / RC_WARPS);      // one warp per chunk, the block scheduler deals the ragged chunks out
    RNSC_LAUNCH(k_rc_decode, grid, 32 * RC_WARPS, 0, st, d_comp, d_comp_off, d_n, d_sample_off, (uint32_t)nchunks, d_samples, d_status);
    c->launches += 1;
    CUDA_OK(cudaGetLastError());
    return RNSC_OK;
}

// ---------------------------------------------------------------------------------------------------
// host-buffer batch entry points
// ---------------------------------------------------------------------------------------------------
int rnsc_decompress_batch(rnsc_ctx* c, int codec, const uint8_t* comp, const uint64_t* comp_off, const uint32_t* n, size_t nchunks,
                          int16_t* samples, const uint64_t* sample_off, int32_t* status)
{
    if (!c || !codec_ok(codec) || (nchunks && (!comp || !comp_off || !n || !samples || !sample_off || !status))) return fail(RNSC_ERR_BAD_ARG, "rnsc_decompress_batch: bad argument");
    if (!nchunks) return RNSC_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    if (use_device(c)) return RNSC_ERR_NO_DEVICE;
    for (size_t i = 0; i < nchunks; ++i) if (comp_off[i + 1] < comp_off[i]) return fail(RNSC_ERR_BAD_ARG, "comp_off must be non-decreasing");
    // The batch is cut into slices of about slice_samples() samples that flow through two lanes of device staging:
    // H2D of slice i+1, decode of slice i and D2H of slice i-1 overlap (three streams chained by events). Chunk i may
    // live anywhere in `comp`/`samples`; inside a lane the chunks are staged back to back (the kernels accept any
    // byte/sample alignment) with dense offset tables.
    // The slice loop runs inside a lambda so that EVERY exit (CUDA error, allocation failure, kernel launch failure) passes
    // the clean-up below: copies into the caller's buffers may still be in flight and lanes marked used.
    auto run = [&]() -> int {
    size_t first = 0; int slice = 0; int rc = RNSC_OK;
    while (first < nchunks && rc == RNSC_OK) {
        size_t last = first; uint64_t cb = 0, sb = 0; uint32_t max_n = 0;
        while (last < nchunks && (last == first || sb + n[last] <= slice_samples())) { cb += comp_off[last + 1] - comp_off[last]; sb += n[last]; if (n[last] > max_n) max_n = n[last]; ++last; }
        const size_t m = last - first;
        Lane& L = c->lane[slice & 1];
        if (L.used) {                                             // the lane's previous slice has left the device: its pinned tables are free, its statuses go to the caller
            CUDA_OK(cudaEventSynchronize(L.ev_out));
            memcpy(status + L.st_first, L.h_status.p, L.st_count * 4); L.st_count = 0;
        }
        if (L.comp.reserve(cb + 64) || L.samples.reserve(sb * 2 + 64) || L.off_a.reserve((m + 1) * 8) || L.off_b.reserve((m + 1) * 8) ||
            L.n.reserve(m * 4 + 16) || L.status.reserve(m * 4 + 16) || L.h_off.reserve(2 * (m + 1) * 8) || L.h_status.reserve(m * 4 + 16))
            return fail(RNSC_ERR_NO_DEVICE, "cudaMalloc(staging) failed");
        uint64_t* coff = (uint64_t*)L.h_off.p; uint64_t* soff = coff + (m + 1);
        bool contiguous = true, dense = true;
        { uint64_t a = 0, b = 0;
          for (size_t i = 0; i < m; ++i) {
              coff[i] = a; soff[i] = b;
              if (comp_off[first + i] - comp_off[first] != a) contiguous = false;
              if (sample_off[first + i] - sample_off[first] != b) dense = false;
              a += comp_off[first + i + 1] - comp_off[first + i]; b += n[first + i];
          }
          coff[m] = a; soff[m] = b; }
        // ---- H2D ----
        if (L.used) { CUDA_OK(cudaStreamWaitEvent(c->h2d, L.ev_dec, 0)); }
        if (contiguous) CUDA_OK(cudaMemcpyAsync(L.comp.p, comp + comp_off[first], cb, cudaMemcpyHostToDevice, c->h2d));
        else for (size_t i = 0; i < m; ++i) CUDA_OK(cudaMemcpyAsync((uint8_t*)L.comp.p + coff[i], comp + comp_off[first + i], comp_off[first + i + 1] - comp_off[first + i], cudaMemcpyHostToDevice, c->h2d));
        CUDA_OK(cudaMemsetAsync((uint8_t*)L.comp.p + cb, 0, 32, c->h2d));
        CUDA_OK(cudaMemcpyAsync(L.off_a.p, coff, (m + 1) * 8, cudaMemcpyHostToDevice, c->h2d));
        CUDA_OK(cudaMemcpyAsync(L.off_b.p, soff, (m + 1) * 8, cudaMemcpyHostToDevice, c->h2d));
        CUDA_OK(cudaMemcpyAsync(L.n.p, n + first, m * 4, cudaMemcpyHostToDevice, c->h2d));
        CUDA_OK(cudaEventRecord(L.ev_in, c->h2d));
        // ---- decode ----
        cudaStream_t st = c->stream;
        CUDA_OK(cudaStreamWaitEvent(st, L.ev_in, 0));
        if (L.used) CUDA_OK(cudaStreamWaitEvent(st, L.ev_out, 0));
        if (codec == RNSC_RC) rc = rnsc_dev_rc_decode(c, (const uint8_t*)L.comp.p, (const uint64_t*)L.off_a.p, (const uint32_t*)L.n.p, (const uint64_t*)L.off_b.p, m, (int16_t*)L.samples.p, (int32_t*)L.status.p, st);
        else rc = launch_decode(c, codec, (const uint8_t*)L.comp.p, (const uint64_t*)L.off_a.p, (const uint32_t*)L.n.p, (const uint64_t*)L.off_b.p, m, max_n, (int16_t*)L.samples.p, (int32_t*)L.status.p, st, cb);
        if (rc) break;
        CUDA_OK(cudaEventRecord(L.ev_dec, st));
        // ---- D2H ----
        CUDA_OK(cudaStreamWaitEvent(c->d2h, L.ev_dec, 0));
        CUDA_OK(cudaMemcpyAsync(L.h_status.p, L.status.p, m * 4, cudaMemcpyDeviceToHost, c->d2h));
        L.st_first = first; L.st_count = m;
        if (dense) CUDA_OK(cudaMemcpyAsync(samples + sample_off[first], L.samples.p, sb * 2, cudaMemcpyDeviceToHost, c->d2h));
        else for (size_t i = 0; i < m; ++i) if (n[first + i]) CUDA_OK(cudaMemcpyAsync(samples + sample_off[first + i], (int16_t*)L.samples.p + soff[i], (size_t)n[first + i] * 2, cudaMemcpyDeviceToHost, c->d2h));
        CUDA_OK(cudaEventRecord(L.ev_out, c->d2h));
        L.used = true;
        first = last; ++slice;
    }
    return rc;
    };
    int rc = run();
    const cudaError_t e1 = cudaStreamSynchronize(c->h2d), e2 = cudaStreamSynchronize(c->stream), e3 = cudaStreamSynchronize(c->d2h);
    for (int k = 0; k < 2; ++k) {
        Lane& L = c->lane[k];
        if (L.st_count && e3 == cudaSuccess) memcpy(status + L.st_first, L.h_status.p, L.st_count * 4);      // the last slices' statuses
        L.st_count = 0; L.used = false;
    }
    if (rc == RNSC_OK && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess))
        rc = fail(RNSC_ERR_NO_DEVICE, "rnsc_decompress_batch: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
    return rc;
}

int rnsc_compress_batch(rnsc_ctx* c, int codec, const int16_t* samples, const uint64_t* sample_off, const uint32_t* n, size_t nchunks,
                        uint8_t* comp, size_t comp_cap, uint64_t* comp_off, int32_t* status)
{
    if (!c || !codec_ok(codec) || !comp_off || (nchunks && (!samples || !sample_off || !n || !comp || !status))) return fail(RNSC_ERR_BAD_ARG, "rnsc_compress_batch: bad argument");
    comp_off[0] = 0;
    if (!nchunks) return RNSC_OK;
    if (codec != RNSC_RC && c->entropy_stage == 0 && !zstd().ok) return fail(RNSC_ERR_NO_ZSTD, "libzstd.so.1 not loadable: %s", dlerror());
    std::lock_guard<std::mutex> lk(c->mu);
    if (use_device(c)) return RNSC_ERR_NO_DEVICE;
    cudaStream_t st = c->stream;
    if (codec == RNSC_RC) {
        std::vector<uint64_t> soff(nchunks + 1);
        uint64_t sb = 0;
        for (size_t i = 0; i < nchunks; ++i) { soff[i] = sb; sb += n[i]; }
        soff[nchunks] = sb;
        if (c->d_samples.reserve(sb * 2 + 64) || c->d_off_b.reserve((nchunks + 1) * 8) || c->d_n.reserve(nchunks * 4 + 16) || c->d_status.reserve(nchunks * 4 + 16))
            return fail(RNSC_ERR_NO_DEVICE, "cudaMalloc(staging) failed");
        bool dense = true;
        for (size_t i = 0; i < nchunks; ++i) if (sample_off[i] - sample_off[0] != soff[i]) { dense = false; break; }
        if (dense) CUDA_OK(cudaMemcpyAsync(c->d_samples.p, samples + sample_off[0], sb * 2, cudaMemcpyHostToDevice, st));
        else for (size_t i = 0; i < nchunks; ++i) if (n[i]) CUDA_OK(cudaMemcpyAsync((int16_t*)c->d_samples.p + soff[i], samples + sample_off[i], (size_t)n[i] * 2, cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaMemcpyAsync(c->d_off_b.p, soff.data(), (nchunks + 1) * 8, cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaMemcpyAsync(c->d_n.p, n, nchunks * 4, cudaMemcpyHostToDevice, st));
        std::vector<uint64_t> slot(nchunks + 1);
        uint64_t tot = 0;
        for (size_t i = 0; i < nchunks; ++i) { slot[i] = tot; tot += (rnsc_max_compressed_size(RNSC_RC, n[i]) + 24 + 15) & ~(size_t)15; }
        slot[nchunks] = tot;
        if (c->d_comp.reserve(tot + 64) || c->d_off_a.reserve((nchunks + 1) * 8) || c->d_sizes.reserve(nchunks * 4 + 16)) return fail(RNSC_ERR_NO_DEVICE, "cudaMalloc(staging) failed");
        CUDA_OK(cudaMemcpyAsync(c->d_off_a.p, slot.data(), (nchunks + 1) * 8, cudaMemcpyHostToDevice, st));
        int r = rnsc_dev_rc_encode(c, (const int16_t*)c->d_samples.p, (const uint64_t*)c->d_off_b.p, (const uint32_t*)c->d_n.p, nchunks, 0, (uint8_t*)c->d_comp.p,
                                   (const uint64_t*)c->d_off_a.p, (uint32_t*)c->d_sizes.p, (int32_t*)c->d_status.p, st);
        if (r) return r;
        std::vector<uint32_t> csz(nchunks);
        CUDA_OK(cudaMemcpyAsync(csz.data(), c->d_sizes.p, nchunks * 4, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaMemcpyAsync(status, c->d_status.p, nchunks * 4, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        uint64_t op = 0;
        for (size_t i = 0; i < nchunks; ++i) {
            comp_off[i] = op;
            if (status[i] == RNSC_OK) {
                if (op + csz[i] > comp_cap) { cudaStreamSynchronize(st); return fail(RNSC_ERR_DST_TOO_SMALL, "output buffer too small"); }   // earlier copies into `comp` are in flight
                CUDA_OK(cudaMemcpyAsync(comp + op, (uint8_t*)c->d_comp.p + slot[i], csz[i], cudaMemcpyDeviceToHost, st));
                op += csz[i];
            }
        }
        comp_off[nchunks] = op;
        CUDA_OK(cudaStreamSynchronize(st));
        return RNSC_OK;
    }

    if (c->entropy_stage == 1) {
        // ---- GPU entropy stage: samples in, finished chunks out; nothing but the copies touches the host. Slices flow through
        //      the two staging lanes like in rnsc_decompress_batch: H2D of slice k+1, encode of slice k and D2H of slice k-1
        //      overlap on three streams. A slice's place in the caller's buffer is known once the sizes of the slices before
        //      it are, so its offset table comes back first (a small copy), the host adds it up, and the payload copy follows.
        struct Pending { size_t first = 0, m = 0; Lane* L = nullptr; };
        Pending pend; bool have = false;
        uint64_t op = 0;
        int rc = RNSC_OK;
        auto finish = [&](const Pending& q) -> int {      // slice q's table is on its way: place it, start its payload copy
            CUDA_OK(cudaEventSynchronize(q.L->ev_tab));
            memcpy(status + q.first, q.L->h_status.p, q.m * 4);
            const uint64_t* coff = (const uint64_t*)q.L->h_off.p + (q.m + 1);
            if (op + coff[q.m] > comp_cap) return fail(RNSC_ERR_DST_TOO_SMALL, "output buffer too small: need %llu bytes, have %zu", (unsigned long long)(op + coff[q.m]), comp_cap);
            if (coff[q.m]) CUDA_OK(cudaMemcpyAsync(comp + op, q.L->comp.p, coff[q.m], cudaMemcpyDeviceToHost, c->d2h));
            CUDA_OK(cudaEventRecord(q.L->ev_out, c->d2h));
            for (size_t i = 0; i < q.m; ++i) comp_off[q.first + i] = op + coff[i];
            op += coff[q.m];
            return RNSC_OK;
        };
        auto run = [&]() -> int {
            size_t first = 0; int slice = 0;
            while (first < nchunks) {
                size_t last = first; uint64_t sb = 0;
                while (last < nchunks && (last == first || sb + n[last] <= slice_samples())) { sb += n[last]; ++last; }
                const size_t m = last - first;
                Lane& L = c->lane[slice & 1];
                // the lane's previous slice (k-2) was finished in the previous iteration: its tables are free, its payload copy is queued
                const size_t ccap = 2 * sb + sb / 4 + 1104 * m + 64;
                if (L.samples.reserve(sb * 2 + 64) || L.off_b.reserve((m + 1) * 8) || L.n.reserve(m * 4 + 16) || L.comp.reserve(ccap + 64) || L.off_a.reserve((m + 1) * 8) ||
                    L.status.reserve(m * 4 + 16) || L.h_off.reserve(2 * (m + 1) * 8) || L.h_status.reserve(m * 4 + 16))
                    return fail(RNSC_ERR_NO_DEVICE, "cudaMalloc(staging) failed");
                uint64_t* soff = (uint64_t*)L.h_off.p;          // [0, m]: sample offsets in the lane; [m+1, 2m+1]: the slice's chunk offsets, filled by the device
                bool dense = true; uint64_t a = 0;
                for (size_t i = 0; i < m; ++i) { soff[i] = a; if (sample_off[first + i] - sample_off[first] != a) dense = false; a += n[first + i]; }
                soff[m] = a;
                // ---- H2D (the lane's sample buffer is free once its previous encode has run) ----
                if (L.used) CUDA_OK(cudaStreamWaitEvent(c->h2d, L.ev_dec, 0));
                if (dense) CUDA_OK(cudaMemcpyAsync(L.samples.p, samples + sample_off[first], sb * 2, cudaMemcpyHostToDevice, c->h2d));
                else for (size_t i = 0; i < m; ++i) if (n[first + i]) CUDA_OK(cudaMemcpyAsync((int16_t*)L.samples.p + soff[i], samples + sample_off[first + i], (size_t)n[first + i] * 2, cudaMemcpyHostToDevice, c->h2d));
                CUDA_OK(cudaMemcpyAsync(L.off_b.p, soff, (m + 1) * 8, cudaMemcpyHostToDevice, c->h2d));
                CUDA_OK(cudaMemcpyAsync(L.n.p, n + first, m * 4, cudaMemcpyHostToDevice, c->h2d));
                CUDA_OK(cudaEventRecord(L.ev_in, c->h2d));
                // ---- encode (the lane's output buffer is free once its previous payload has left) ----
                CUDA_OK(cudaStreamWaitEvent(st, L.ev_in, 0));
                if (L.used) CUDA_OK(cudaStreamWaitEvent(st, L.ev_out, 0));
                const int r = rnsc_dev_encode(c, codec, (const int16_t*)L.samples.p, (const uint64_t*)L.off_b.p, (const uint32_t*)L.n.p, m, sb, (uint8_t*)L.comp.p, ccap,
                                              (uint64_t*)L.off_a.p, (int32_t*)L.status.p, st);
                if (r) return r;
                CUDA_OK(cudaEventRecord(L.ev_dec, st));
                // ---- the slice's offset table and statuses ----
                CUDA_OK(cudaStreamWaitEvent(c->d2h, L.ev_dec, 0));
                CUDA_OK(cudaMemcpyAsync(soff + (m + 1), L.off_a.p, (m + 1) * 8, cudaMemcpyDeviceToHost, c->d2h));
                CUDA_OK(cudaMemcpyAsync(L.h_status.p, L.status.p, m * 4, cudaMemcpyDeviceToHost, c->d2h));
                CUDA_OK(cudaEventRecord(L.ev_tab, c->d2h));
                L.used = true;
                // ---- the previous slice: place it, start its payload copy (this slice's H2D and encode are already queued) ----
                if (have) { const int r2 = finish(pend); if (r2) return r2; }
                pend.first = first; pend.m = m; pend.L = &L; have = true;
                first = last; ++slice;
            }
            if (have) { const int r2 = finish(pend); if (r2) return r2; }
            return RNSC_OK;
        };
        rc = run();
        const cudaError_t e1 = cudaStreamSynchronize(c->h2d), e2 = cudaStreamSynchronize(st), e3 = cudaStreamSynchronize(c->d2h);
        for (int k = 0; k < 2; ++k) c->lane[k].used = false;
        if (rc == RNSC_OK && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess))
            rc = fail(RNSC_ERR_NO_DEVICE, "rnsc_compress_batch: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
        if (rc == RNSC_OK) comp_off[nchunks] = op;
        return rc;
    }

    // ---- sliced pipeline: while the host threads run the zstd frame stage of slice k, this thread stages slice k+1
    //      (H2D, splitter kernels, D2H of the streams into the other pinned buffer) ----
    const int nf = codec_frames(codec);
    Zstd& z = zstd();
    struct SliceJob { size_t c0 = 0, c1 = 0; std::vector<uint64_t> stoff, slot; uint8_t* arena = nullptr; std::thread th; };
    std::vector<size_t> cut{0};
    { uint64_t acc = 0; for (size_t i = 0; i < nchunks; ++i) { if (acc && acc + n[i] > slice_samples()) { cut.push_back(i); acc = 0; } acc += n[i]; } }
    cut.push_back(nchunks);
    const size_t nsl = cut.size() - 1;
    std::vector<SliceJob> jobs(nsl);
    std::vector<uint32_t> csize(nchunks, 0);
    PinBuf* hb[2] = {&c->h_a, &c->h_b};
    auto join_all = [&] { for (auto& j : jobs) { if (j.th.joinable()) j.th.join(); } };
    auto free_all = [&] { for (auto& j : jobs) { free(j.arena); j.arena = nullptr; } };
    int rc = RNSC_OK;
    for (size_t k = 0; k < nsl && rc == RNSC_OK; ++k) {
        SliceJob& J = jobs[k];
        J.c0 = cut[k]; J.c1 = cut[k + 1];
        const size_t m = J.c1 - J.c0;
        const uint32_t* nn = n + J.c0;
        std::vector<uint64_t> soff(m + 1);
        uint64_t sb = 0, scap = 0;
        bool dense = true;
        for (size_t i = 0; i < m; ++i) { soff[i] = sb; sb += nn[i]; scap += stream_bound(codec, nn[i]); if (sample_off[J.c0 + i] - sample_off[J.c0] != soff[i]) dense = false; }
        soff[m] = sb;
        if (c->d_samples.reserve(sb * 2 + 64) || c->d_off_b.reserve((m + 1) * 8) || c->d_n.reserve(m * 4 + 16) || c->d_streams.reserve(scap + 64) || c->d_stream_off.reserve((5 * m + 1) * 8)) {
            rc = fail(RNSC_ERR_NO_DEVICE, "cudaMalloc(staging) failed"); break;
        }
        cudaError_t e = cudaSuccess;
        if (dense) e = cudaMemcpyAsync(c->d_samples.p, samples + sample_off[J.c0], sb * 2, cudaMemcpyHostToDevice, st);
        else for (size_t i = 0; i < m && e == cudaSuccess; ++i) if (nn[i]) e = cudaMemcpyAsync((int16_t*)c->d_samples.p + soff[i], samples + sample_off[J.c0 + i], (size_t)nn[i] * 2, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_off_b.p, soff.data(), (m + 1) * 8, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_n.p, nn, m * 4, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { rc = fail(RNSC_ERR_NO_DEVICE, "cudaMemcpyAsync: %s", cudaGetErrorString(e)); break; }
        // splitter stage on the device
        rc = rnsc_dev_svb_encode(c, codec, (const int16_t*)c->d_samples.p, (const uint64_t*)c->d_off_b.p, (const uint32_t*)c->d_n.p, m,
                                 (uint8_t*)c->d_streams.p, scap, (uint64_t*)c->d_stream_off.p, st);
        if (rc) break;
        J.stoff.resize(5 * m + 1);
        e = cudaMemcpyAsync(J.stoff.data(), c->d_stream_off.p, J.stoff.size() * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { rc = fail(RNSC_ERR_NO_DEVICE, "stream offsets: %s", cudaGetErrorString(e)); break; }
        const uint64_t total = J.stoff[5 * m];
        if (k >= 2 && jobs[k - 2].th.joinable()) jobs[k - 2].th.join();          // its pinned buffer is free again
        PinBuf* H = hb[k & 1];
        if (H->reserve(total + 64)) { rc = fail(RNSC_ERR_NO_DEVICE, "cudaMallocHost(streams) failed"); break; }
        e = cudaMemcpyAsync(H->p, c->d_streams.p, total, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { rc = fail(RNSC_ERR_NO_DEVICE, "streams D2H: %s", cudaGetErrorString(e)); break; }
        // one slot per chunk in the slice's arena (worst-case frame sizes): no per-chunk allocation
        J.slot.resize(m + 1);
        uint64_t need_all = 0;
        for (size_t i = 0; i < m; ++i) {
            J.slot[i] = need_all;
            if (nf == 1) need_all += rnsc_max_compressed_size(codec, nn[i]);
            else for (int f = 0; f < nf; ++f) need_all += z.bound((size_t)(J.stoff[5 * i + f + 1] - J.stoff[5 * i + f])) + 8;
        }
        J.slot[m] = need_all;
        J.arena = (uint8_t*)malloc(need_all + 64);
        if (!J.arena) { rc = fail(RNSC_ERR_NO_DEVICE, "malloc(frame arena %llu) failed", (unsigned long long)need_all); break; }
        // zstd frame stage on host threads (byte-identical with the reference's ZSTD_compress(.., 1))
        const uint8_t* hs = (const uint8_t*)H->p;
        const int threads = c->host_threads;
        auto work = [&J, &z, hs, nn, nf, codec, status, &csize, threads, m]() {
            parallel_for(m, threads, [&](size_t i) {
                const uint64_t* so = &J.stoff[5 * i];
                uint8_t* o = J.arena + J.slot[i];
                const size_t need = (size_t)(J.slot[i + 1] - J.slot[i]);
                const size_t cap = rnsc_max_compressed_size(codec, nn[i]);
                const size_t gi = J.c0 + i;
                status[gi] = RNSC_OK;
                if (nf == 1) {       // keys||data are adjacent in the stream buffer: one frame
                    size_t cs = z.compress(o, cap, hs + so[0], (size_t)(so[2] - so[0]), 1);
                    if (z.is_error(cs)) status[gi] = RNSC_ERR_ZSTD_ENCODE; else csize[gi] = (uint32_t)cs;
                    return;
                }
                size_t op = 0;
                for (int f = 0; f < nf; ++f) {
                    const size_t len = (size_t)(so[f + 1] - so[f]);
                    const size_t hdr = f + 1 < nf ? 8 : 0;
                    size_t cs = z.compress(o + op + hdr, need - op - hdr, hs + so[f], len, 1);
                    if (z.is_error(cs)) { status[gi] = RNSC_ERR_ZSTD_ENCODE; return; }
                    if (hdr) { uint64_t v = cs; memcpy(o + op, &v, 8); }
                    op += hdr + cs;
                }
                // destination capacity max(2n+26,1024): C4/C5 report it (C5.hpp:420-427); C1-C3 overrun in the
                // reference, which is reported as the same error here.
                if (op > cap) { status[gi] = RNSC_ERR_DST_TOO_SMALL; return; }
                csize[gi] = (uint32_t)op;
            });
        };
        if (nsl == 1) work(); else J.th = std::thread(work);
    }
    join_all();
    if (rc) { free_all(); return rc; }
    uint64_t op = 0;
    for (size_t i = 0; i < nchunks; ++i) { comp_off[i] = op; op += csize[i]; }
    if (op > comp_cap) { free_all(); return fail(RNSC_ERR_DST_TOO_SMALL, "output buffer too small: need %llu bytes, have %zu", (unsigned long long)op, comp_cap); }
    // parallel compaction of the slots into the caller's buffer
    const size_t group = 64;                                         // chunks per copy task
    std::vector<uint32_t> slice_of(nchunks);
    for (size_t k = 0; k < nsl; ++k) for (size_t i = jobs[k].c0; i < jobs[k].c1; ++i) slice_of[i] = (uint32_t)k;
    parallel_for((nchunks + group - 1) / group, c->host_threads, [&](size_t g) {
        for (size_t i = g * group; i < std::min(nchunks, (g + 1) * group); ++i) {
            const SliceJob& J = jobs[slice_of[i]];
            if (csize[i]) memcpy(comp + comp_off[i], J.arena + J.slot[i - J.c0], csize[i]);
        }
    });
    free_all();
    comp_off[nchunks] = op;
    return RNSC_OK;
}

// ---------------------------------------------------------------------------------------------------
// per-chunk entry points
// ---------------------------------------------------------------------------------------------------
int rnsc_compress_chunk(rnsc_ctx* c, int codec, const int16_t* samples, size_t n, uint8_t* out, size_t* inout_size)
{
    if (!c || !codec_ok(codec) || !out || !inout_size || (n && !samples) || n > 0xFFFFFFFFull) return fail(RNSC_ERR_BAD_ARG, "rnsc_compress_chunk: bad argument");
    static const int16_t dummy = 0;
    const uint64_t soff = 0; const uint32_t nn = (uint32_t)n; uint64_t coff[2]; int32_t st = 0;
    std::vector<uint8_t> tmp(rnsc_max_compressed_size(codec, n) + 64);
    int r = rnsc_compress_batch(c, codec, n ? samples : &dummy, &soff, &nn, 1, tmp.data(), tmp.size(), coff, &st);
    if (r) return r;
    if (st) return fail(st, "%s", rnsc_status_string(st));
    if (coff[1] > *inout_size) return fail(RNSC_ERR_DST_TOO_SMALL, "Compressed signal size (%llu) is greater than provided buffer size (%zu)", (unsigned long long)coff[1], *inout_size);
    memcpy(out, tmp.data(), coff[1]);
    *inout_size = coff[1];
    return RNSC_OK;
}

int rnsc_decompress_chunk(rnsc_ctx* c, int codec, const uint8_t* in, size_t in_size, size_t n, int16_t* out)
{
    static int16_t dummy_out[8];
    if (!c || !codec_ok(codec) || !in || (n && !out) || n > 0xFFFFFFFFull) return fail(RNSC_ERR_BAD_ARG, "rnsc_decompress_chunk: bad argument");
    const uint64_t coff[2] = {0, in_size}; const uint64_t soff = 0; const uint32_t nn = (uint32_t)n; int32_t st = 0;
    int r = rnsc_decompress_batch(c, codec, in, coff, &nn, 1, n ? out : dummy_out, &soff, &st);
    if (r) return r;
    if (st) return fail(st, "%s", rnsc_status_string(st));
    return RNSC_OK;
}

// ---------------------------------------------------------------------------------------------------
// pod5-named entry points (c_api.h:678-711) on one process-wide context
// ---------------------------------------------------------------------------------------------------
// One context per calling thread, created on first use on the current device and handed to another thread when its
// owner exits (AsyncSignalLoader-style worker pools call the per-chunk functions concurrently, async_signal_loader.cpp:
// 176-220; a single mutexed context would serialise them). Contexts are parked, never destroyed: at process exit the
// CUDA runtime may already be gone when thread-local destructors run.
static std::mutex g_pool_mu;
static std::vector<rnsc_ctx*> g_ctx_pool;
static std::atomic<int> g_variant{RNSC_C5};
namespace {
struct ThreadCtx {
    rnsc_ctx* c = nullptr;
    ~ThreadCtx() { if (c) { std::lock_guard<std::mutex> lk(g_pool_mu); g_ctx_pool.push_back(c); } }
};
thread_local ThreadCtx t_ctx;
}  // namespace
rnsc_ctx* rnsc_thread_context(void)
{
    if (t_ctx.c) return t_ctx.c;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (!g_ctx_pool.empty()) { t_ctx.c = g_ctx_pool.back(); g_ctx_pool.pop_back(); }
    }
    if (!t_ctx.c) { t_ctx.c = rnsc_create(-1); if (t_ctx.c) t_ctx.c->host_threads = 1; }     // a per-chunk caller compresses its frames on its own thread
    return t_ctx.c;
}
static rnsc_ctx* proc_ctx() { return rnsc_thread_context(); }
int rnsc_set_pgnano_variant(int codec) { if (codec < RNSC_VBZ || codec > RNSC_C5) return fail(RNSC_ERR_BAD_ARG, "variant must be one of VBZ, VBZ0, C1..C5"); g_variant = codec; return RNSC_OK; }
int rnsc_get_pgnano_variant(void) { return g_variant; }

size_t rnsc_pod5_vbz_compressed_signal_max_size(size_t sample_count) { return rnsc_max_compressed_size(RNSC_VBZ, sample_count); }

rnsc_pod5_error_t rnsc_pod5_vbz_compress_signal(const int16_t* signal, size_t signal_size, char* out, size_t* size)
{
    if (!signal || !out || !size) return fail(RNSC_ERR_BAD_ARG, "null passed to C API");
    rnsc_ctx* c = proc_ctx(); if (!c) return RNSC_ERR_NO_DEVICE;
    return rnsc_compress_chunk(c, RNSC_VBZ, signal, signal_size, (uint8_t*)out, size);
}
rnsc_pod5_error_t rnsc_pod5_pinanoraw_compress_signal(const int16_t* signal, size_t signal_size, char* out, size_t* size)
{
    if (!signal || !out || !size) return fail(RNSC_ERR_BAD_ARG, "null passed to C API");
    rnsc_ctx* c = proc_ctx(); if (!c) return RNSC_ERR_NO_DEVICE;
    return rnsc_compress_chunk(c, g_variant, signal, signal_size, (uint8_t*)out, size);
}
rnsc_pod5_error_t rnsc_pod5_vbz_decompress_signal(const char* in, size_t in_size, size_t sample_count, short* out)
{
    if (!in || !out) return fail(RNSC_ERR_BAD_ARG, "null passed to C API");
    rnsc_ctx* c = proc_ctx(); if (!c) return RNSC_ERR_NO_DEVICE;
    return rnsc_decompress_chunk(c, RNSC_VBZ, (const uint8_t*)in, in_size, sample_count, out);
}
rnsc_pod5_error_t rnsc_pod5_pinanoraw_decompress_signal(const char* in, size_t in_size, size_t sample_count, short* out)
{
    if (!in || !out) return fail(RNSC_ERR_BAD_ARG, "null passed to C API");
    rnsc_ctx* c = proc_ctx(); if (!c) return RNSC_ERR_NO_DEVICE;
    return rnsc_decompress_chunk(c, g_variant, (const uint8_t*)in, in_size, sample_count, out);
}

}  // extern "C"
