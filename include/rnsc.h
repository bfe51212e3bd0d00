/* rnsc.h -- C ABI of librnsc.so, the B200-native drop-in for the per-chunk raw-signal codec hot path of
 * tomas-gr/RawNanoporeSignalCompression.
 *
 * Plain pointers and sizes only; no exceptions cross this boundary; every entry point is thread-safe
 * when each thread uses its own rnsc_ctx (the pod5-named wrappers use one context per calling thread).
 * The library needs a CUDA device: rnsc_create() fails with RNSC_ERR_NO_DEVICE otherwise -- there is
 * no CPU fallback.
 *
 * Codec ids. The reference selects the pgnano variant at COMPILE time (pgnano/pgnano.cpp:1,72-125); a
 * POD5 file only records "pgnano.signal" (pod5_format/types.h:90), so here it is a runtime parameter.
 */
#ifndef RNSC_H
#define RNSC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum rnsc_codec {
    RNSC_VBZ = 0,  /* stock VBZ: pod5_format/signal_compression.cpp:37-141 (binary VBZ1, flag --VBZ)      */
    RNSC_VBZ0 = 1, /* pgnano/svb16/VBZ_0.hpp:316-423  (2-bit keys, one nibble stream, one frame)        */
    RNSC_C1 = 2,   /* pgnano/svb16/C1.hpp:202-382     (VBZ keys and data in separate frames)            */
    RNSC_C2 = 3,   /* pgnano/svb16/C2.hpp:195-463     (keys, low byte of every sample, high of large)   */
    RNSC_C3 = 4,   /* pgnano/svb16/C3.hpp:212-503     (keys, ll, lh, h)                                 */
    RNSC_C4 = 5,   /* pgnano/svb16/C4.hpp:274-679     (2-bit keys, S/M/Llow/Lhigh, no offsets)          */
    RNSC_C5 = 6,   /* pgnano/svb16/C5.hpp:282-683     (as C4 with offsets 1/17/273; reference default)  */
    RNSC_RC = 7    /* legacy range coder: pgnano/compressor.cpp:39-88, decompressor.cpp:45-73           */
} rnsc_codec;

/* Status codes (per call and per chunk). The reference reports arrow::Status::Invalid(msg); the message
 * each code stands for is returned by rnsc_status_string(). */
typedef enum rnsc_status {
    RNSC_OK = 0,
    RNSC_ERR_NOT_ZSTD = 1,      /* "Input data not compressed by zstd"       signal_compression.cpp:105-111 */
    RNSC_ERR_ZSTD_DECODE = 2,   /* "Input data failed to decompress using zstd"              ibid.:122-129 */
    RNSC_ERR_DST_TOO_SMALL = 3, /* "Not enough space in destination buffer" (C5.hpp:420-427) / "Compressed
                                   signal size is greater than provided buffer size" (c_api.cpp:1201-1209)   */
    RNSC_ERR_REMAINING = 4,     /* "Remaining data at end of signal buffer"          signal_compression.cpp:138 */
    RNSC_ERR_CONTAINER = 5,     /* Cx size prefixes run past the chunk (the reference reads out of bounds)   */
    RNSC_ERR_ZSTD_ENCODE = 6,   /* "Failed to compress data"                          signal_compression.cpp:67 */
    RNSC_ERR_RC_RAW = 7,        /* range coder would take the reference's broken raw fallback (compressor.cpp:60-65) */
    RNSC_ERR_BAD_ARG = 8,       /* null pointer / unknown codec (check_not_null, c_api.cpp:1190-1195)        */
    RNSC_ERR_NO_DEVICE = 9,     /* no CUDA device / CUDA runtime error (message has the CUDA error string)   */
    RNSC_ERR_NO_ZSTD = 10       /* libzstd.so.1 (host entropy stage of the encoders) could not be loaded     */
} rnsc_status;

typedef struct rnsc_ctx rnsc_ctx;

/* ---- library / context ------------------------------------------------------------------------- */
const char* rnsc_version(void);
int rnsc_device_count(void);
const char* rnsc_status_string(int status);
/* Last error text of the calling thread (replaces pod5_get_error_string, c_api.cpp:49-58, which is
 * process-global and not thread-safe in the reference). */
const char* rnsc_last_error(void);
/* Context = one CUDA device + one stream + reusable device/pinned staging. `device` < 0: current device. */
rnsc_ctx* rnsc_create(int device);
void rnsc_destroy(rnsc_ctx* ctx);
/* Host threads used for the zstd frame stage of the encoders (default: hardware concurrency). */
void rnsc_set_host_threads(rnsc_ctx* ctx, int n);

/* Entropy stage of rnsc_compress_batch / rnsc_compress_chunk (codecs VBZ .. C5):
 *   RNSC_ENTROPY_LIBZSTD (default): the device splitter feeds ZSTD_compress(.., 1) of the image's libzstd.so.1 on host
 *     threads -- the bytes are identical to the reference's (same library, same call: C5.hpp:337-407,
 *     signal_compression.cpp:65);
 *   RNSC_ENTROPY_GPU: the frames are written on the device (csrc/zstd_encode.cuh: raw / RLE / Huffman-literal blocks, no
 *     sequences). Valid Zstandard frames that libzstd, the reference's decoders and this library decode to the same
 *     samples; the BYTES differ from libzstd's (decode parity, not byte parity) and the ratio is that of an order-0
 *     coder (no LZ matches: identical on C2-C5 in practice, a little larger on VBZ). No host work, no libzstd needed. */
#define RNSC_ENTROPY_LIBZSTD 0
#define RNSC_ENTROPY_GPU 1
int rnsc_set_entropy_stage(rnsc_ctx* ctx, int mode);
int rnsc_get_entropy_stage(const rnsc_ctx* ctx);

/* ---- sizes -------------------------------------------------------------------------------------- */
/* VBZ: compressed_signal_max_size (signal_compression.cpp:30-35) = ZSTD_compressBound(ceil(n/8)+2n).
 * Cx/RC: Compressor::compressed_signal_max_size (pgnano/compressor.h:39-45) = max(2n+26, 1024). */
size_t rnsc_max_compressed_size(int codec, size_t n);
/* Upper bound of the pre-zstd streams of one chunk (sum over streams). */
size_t rnsc_max_stream_bytes(int codec, size_t n);
/* Number of streams (= zstd frames for Cx) of a codec: VBZ/VBZ0 2 streams in 1 frame; C1 2; C2 3; C3 4; C4/C5 5. */
int rnsc_stream_count(int codec);

/* ---- per-chunk, host buffers (replaces pod5::compress_signal / decompress_signal,
 *      pgnano::compress_signal / decompress_signal; signal_compression.h:17-36, pgnano/pgnano.h:13-23) - */
/* *inout_size: capacity of `out` on entry, bytes written on success (c_api.h:684-700 convention). */
int rnsc_compress_chunk(rnsc_ctx* ctx, int codec, const int16_t* samples, size_t n, uint8_t* out, size_t* inout_size);
/* The decoder needs n from the signal table's `samples` column (signal_table_reader.cpp:115-120). */
int rnsc_decompress_chunk(rnsc_ctx* ctx, int codec, const uint8_t* in, size_t in_size, size_t n, int16_t* out);

/* ---- batch, host buffers (the POD5 hooks a batched codec sits behind: pod5_add_reads_data_pre_compressed
 *      c_api.h:665-673 on write, SignalTableReader::extract_samples_inplace signal_table_reader.cpp:320-340
 *      on read). Chunk i covers samples[sample_off[i] .. +n[i]) and comp[comp_off[i] .. comp_off[i+1]).
 *      status[i] receives the per-chunk rnsc_status. Returns RNSC_OK if the call itself succeeded.
 *      Both calls cut the batch into slices (~128 MiB of int16) that flow through two lanes of device staging on three
 *      streams: H2D of slice k+1, kernels of slice k and D2H of slice k-1 overlap when the caller's buffers are pinned
 *      (cudaHostAlloc / cudaHostRegister) -- decode, and encode with RNSC_ENTROPY_GPU, then run at the PCIe rate; with
 *      RNSC_ENTROPY_LIBZSTD the host threads compress slice k while slice k+1 is staged. ------ */
int rnsc_compress_batch(rnsc_ctx* ctx, int codec, const int16_t* samples, const uint64_t* sample_off,
                        const uint32_t* n, size_t nchunks, uint8_t* comp, size_t comp_cap,
                        uint64_t* comp_off /* nchunks+1 */, int32_t* status);
int rnsc_decompress_batch(rnsc_ctx* ctx, int codec, const uint8_t* comp, const uint64_t* comp_off /* nchunks+1 */,
                          const uint32_t* n, size_t nchunks, int16_t* samples, const uint64_t* sample_off,
                          int32_t* status);

/* ---- device-resident batch (all d_* pointers are device memory; `stream` is a cudaStream_t or NULL for the
 *      context's own stream; calls are asynchronous on that stream unless stated) ---------------------- */
/* Whole-chunk decode: zstd frames -> streams -> int16, one fused launch group. d_comp must be 16-byte aligned
 * and have 16 readable bytes after the last chunk. max_n = max over n[] (sizes per-CTA scratch).
 * Calls on one context share its planner scratch: they run one after another on the device even when they
 * are enqueued on different streams (each call waits for the previous call's completion event); use one
 * context per stream for decodes that should overlap. Scratch is sized from the data and grows with what
 * the context has seen: a batch with far more LZ sequences per chunk than any before it is still decoded
 * correctly, its overflow by the general decoder (rnsc_decode_class_counts out[3]). */
int rnsc_dev_decode(rnsc_ctx* ctx, int codec, const uint8_t* d_comp, const uint64_t* d_comp_off,
                    const uint32_t* d_n, const uint64_t* d_sample_off, size_t nchunks, uint32_t max_n,
                    int16_t* d_samples, int32_t* d_status, void* stream);
/* Whole-chunk encode on the device with the GPU entropy stage (RNSC_ENTROPY_GPU above): int16 samples -> finished chunks
 * (container + zstd frames), nothing leaves HBM. total_samples = sum of n[] (sizes the scratch). Chunk i is written to
 * d_comp[d_comp_off[i] .. d_comp_off[i+1]) (d_comp_off: nchunks+1 entries, produced by a device prefix scan of the chunk
 * sizes); d_status[i] is RNSC_ERR_DST_TOO_SMALL when a chunk exceeds the reference's output capacity
 * (rnsc_max_compressed_size) or comp_cap. Replaces pgnano::compress_signal / pod5::compress_signal for a caller that keeps
 * signal on the GPU. Asynchronous on `stream`. */
int rnsc_dev_encode(rnsc_ctx* ctx, int codec, const int16_t* d_samples, const uint64_t* d_sample_off, const uint32_t* d_n,
                    size_t nchunks, size_t total_samples, uint8_t* d_comp, size_t comp_cap, uint64_t* d_comp_off /* nchunks+1 */,
                    int32_t* d_status, void* stream);
/* Splitter stage only (stops before zstd): int16 -> streams. Stream k of chunk i is
 * d_streams[d_stream_off[5*i+k] .. d_stream_off[5*i+k+1]); d_stream_off has 5*nchunks+1 entries and is
 * produced by a device prefix scan of the stream sizes. Returns RNSC_ERR_DST_TOO_SMALL (after a stream
 * synchronisation) when streams_cap is smaller than the total. */
int rnsc_dev_svb_encode(rnsc_ctx* ctx, int codec, const int16_t* d_samples, const uint64_t* d_sample_off,
                        const uint32_t* d_n, size_t nchunks, uint8_t* d_streams, size_t streams_cap,
                        uint64_t* d_stream_off, void* stream);
/* The same in ONE pass over the samples, for consumers on the device that do not need the streams dense: chunk i owns a
 * slab sized by the streams' upper bounds (rnsc_svb_slab_bytes(codec, n[i]), slabs back to back in d_streams), stream k of
 * chunk i is the d_stream_size[5*i+k] bytes at d_streams[d_stream_start[5*i+k]] (keys first; a one-frame codec's payload
 * follows its keys directly). No size pass, no scan of the stream sizes, no synchronisation: the samples are read from HBM
 * once (3.06 instead of 5.06 bytes of DRAM traffic per C5 sample). d_stream_start / d_stream_size: 5*nchunks entries.
 * streams_cap must hold the sum of the slabs (checked on the host from total_samples: 4*total + 128*nchunks + 64 always
 * suffices). This is the splitter rnsc_dev_encode runs in front of its entropy stage. */
size_t rnsc_svb_slab_bytes(int codec, uint32_t n);
int rnsc_dev_svb_encode_slabs(rnsc_ctx* ctx, int codec, const int16_t* d_samples, const uint64_t* d_sample_off,
                              const uint32_t* d_n, size_t nchunks, size_t total_samples, uint8_t* d_streams, size_t streams_cap,
                              uint64_t* d_stream_start /* 5*nchunks */, uint32_t* d_stream_size /* 5*nchunks */, void* stream);
/* Inverse splitter stage only (starts after zstd): streams -> int16. */
int rnsc_dev_svb_decode(rnsc_ctx* ctx, int codec, const uint8_t* d_streams, const uint64_t* d_stream_off,
                        const uint32_t* d_n, const uint64_t* d_sample_off, size_t nchunks, int16_t* d_samples,
                        int32_t* d_status, void* stream);
/* Range coder, device resident. Output slot of chunk i starts at d_comp + d_slot_off[i] and must hold
 * rnsc_max_compressed_size(RNSC_RC, n[i]) bytes; d_comp_size[i] receives the bytes written (0 + status
 * RNSC_ERR_RC_RAW when the reference would have fallen back to raw). pore_type: header byte 9
 * (known_pore_types.h:8-14; benchmark.cpp:207 uses 0 = R10.4.1). */
int rnsc_dev_rc_encode(rnsc_ctx* ctx, const int16_t* d_samples, const uint64_t* d_sample_off, const uint32_t* d_n,
                       size_t nchunks, uint8_t pore_type, uint8_t* d_comp, const uint64_t* d_slot_off,
                       uint32_t* d_comp_size, int32_t* d_status, void* stream);
int rnsc_dev_rc_decode(rnsc_ctx* ctx, const uint8_t* d_comp, const uint64_t* d_comp_off /* nchunks+1 */,
                       const uint32_t* d_n, const uint64_t* d_sample_off, size_t nchunks, int16_t* d_samples,
                       int32_t* d_status, void* stream);
/* Exclusive prefix scan on the device (assigns compressed-output offsets): out[0]=0, out[i]=sum(in[0..i)). */
int rnsc_dev_exclusive_scan_u32(rnsc_ctx* ctx, const uint32_t* d_in, size_t count, uint64_t* d_out /* count+1 */, void* stream);
/* How the last whole-chunk decode of this context routed its chunks: out[0..2] = fast path with four / two / one CTA
 * per SM (small / medium / large chunks), out[3] = general decoder (frames the planner does not cover).
 * Synchronises `stream` (NULL: the context's stream). Diagnostics for tests and bench.py. */
int rnsc_decode_class_counts(rnsc_ctx* ctx, uint32_t out[4], void* stream);
/* Tuning aid: when on, the fast decode kernels add SM cycles per phase (thread 0 of every CTA) to 3 x 8 counters
 * (size class S, A, B) x (stage+tables, Huffman pass 1, chain check+scan, Huffman pass 2, fills+sequences, splitter
 * inverse, bookkeeping, chunks). rnsc_get_phase_cycles synchronises the device and returns the running totals. */
int rnsc_set_profiling(rnsc_ctx* ctx, int on);
int rnsc_get_phase_cycles(rnsc_ctx* ctx, uint64_t out[24]);
/* Same switch: every CTA of the fast decode kernels also records {start ns, end ns (%globaltimer), chunks decoded}.
 * out = 3 classes (S, A, B) x 1024 CTAs x 3 words of the last decode; unused slots are zero. Shows how the kernels of a
 * call share the SMs over time (tools/span_probe.py). */
#define RNSC_SPAN_WORDS (3 * 1024 * 3)
int rnsc_get_cta_spans(rnsc_ctx* ctx, uint64_t out[RNSC_SPAN_WORDS]);
/* Kernels launched by this context since creation (bench.py's gpu_launches). */
uint64_t rnsc_launch_count(const rnsc_ctx* ctx);

/* The calling thread's own context (created on first use on the current CUDA device, handed on to another thread when
 * its owner exits). What the pod5-named entry points below and libpod5_rnsc.so (include/pod5_rnsc.h) run on: concurrent
 * callers -- the reference's AsyncSignalLoader workers, async_signal_loader.cpp:176-220 -- do not serialise. NULL without
 * a CUDA device (rnsc_last_error() says why). */
rnsc_ctx* rnsc_thread_context(void);

/* ---- names the reference's C API exports for this path (pod5_format/c_api.h:678-711), same signatures, rnsc_ prefix
 *      (the un-prefixed names with pod5_error_t values live in libpod5_rnsc.so, include/pod5_rnsc.h: link that library in
 *      front of the reference's to take the calls over); each runs on rnsc_thread_context(); errors via rnsc_last_error(). */
typedef int rnsc_pod5_error_t; /* 0 = POD5_OK */
size_t rnsc_pod5_vbz_compressed_signal_max_size(size_t sample_count);                      /* c_api.cpp:1177 */
rnsc_pod5_error_t rnsc_pod5_vbz_compress_signal(const int16_t* signal, size_t signal_size, char* compressed_signal_out,
                                                size_t* compressed_signal_size);           /* c_api.cpp:1183 */
rnsc_pod5_error_t rnsc_pod5_pinanoraw_compress_signal(const int16_t* signal, size_t signal_size, char* compressed_signal_out,
                                                      size_t* compressed_signal_size);     /* c_api.cpp:1217 */
rnsc_pod5_error_t rnsc_pod5_vbz_decompress_signal(const char* compressed_signal, size_t compressed_signal_size,
                                                  size_t sample_count, short* signal_out); /* c_api.cpp:1255 */
/* The reference has no C entry point for Cx decompression (only C++ pgnano::decompress_signal). */
rnsc_pod5_error_t rnsc_pod5_pinanoraw_decompress_signal(const char* compressed_signal, size_t compressed_signal_size,
                                                        size_t sample_count, short* signal_out);
/* Which Cx variant the two pinanoraw entry points use (default RNSC_C5, the reference's checked-in build). */
int rnsc_set_pgnano_variant(int codec);
int rnsc_get_pgnano_variant(void);

#ifdef __cplusplus
}
#endif
#endif /* RNSC_H */
