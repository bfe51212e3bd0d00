// oracle/ref_cx_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Thin extern "C" wrapper that compiles the *reference's own* VBZ and Cx codec headers, where they lie
// under /root/reference, into oracle/_ref/libref_oracle.so (recipe: oracle/Makefile).
// No reference source is copied into this repository; this file only includes it at build time.
//   VBZ  : pod5_format/svb16/{encode,decode}.hpp + the zstd calls of signal_compression.cpp:37-141
//   Cx   : pod5_format/pgnano/svb16/{VBZ_0,C1,C2,C3,C4,C5}.hpp (compress_signal_* / decompress_signal_*)
#include "pod5_format/types.h"
#include "pod5_format/read_table_utils.h"
#include <zstd.h>
#include <cstring>
#include <string>
#include <vector>

// Global counters the reference headers declare `extern` (defined in its copy.cpp:64-85).
long full_size_keys, full_size_S, full_size_M, full_size_Llow, full_size_Lhigh, full_size_data;
long comp_size_keys, comp_size_S, comp_size_M, comp_size_Llow, comp_size_Lhigh, comp_size_data;
long number_small, number_medium, number_large, total_samples;
double compression_time, decompression_time;

// (The stock VBZ codec is wrapped in a second translation unit, ref_vbz_harness.cpp, because both
// trees define namespace svb16 with identical names.)
#include "pod5_format/pgnano/svb16/svb16.h"
#include "pod5_format/pgnano/svb16/decode.hpp"
#include "pod5_format/pgnano/svb16/encode.hpp"
#include "pod5_format/pgnano/svb16/VBZ_0.hpp"
#include "pod5_format/pgnano/svb16/C1.hpp"
#include "pod5_format/pgnano/svb16/C2.hpp"
#include "pod5_format/pgnano/svb16/C3.hpp"
#include "pod5_format/pgnano/svb16/C4.hpp"
#include "pod5_format/pgnano/svb16/C5.hpp"

static thread_local std::string g_err;

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// codec ids match include/rnsc.h: 0 VBZ (other TU), 1 VBZ0, 2 C1, 3 C2, 4 C3, 5 C4, 6 C5, 7 RC (other TU)
int ref_cx_encode(int codec, const int16_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_size)
{
    auto pool = arrow::default_memory_pool();
    pod5::ReadData rd;
    gsl::span<const int16_t> s(in, n);
    gsl::span<uint8_t> d(out, cap);
    arrow::Result<std::size_t> r = arrow::Status::Invalid("bad codec");
    switch (codec) {
        case 1: r = pgnano::compress_signal_VBZ1(s, pool, d, rd, false); break;
        case 2: r = pgnano::compress_signal_KD(s, pool, d, rd, false); break;
        case 3: r = pgnano::compress_signal_lh(s, pool, d, rd, false); break;
        case 4: r = pgnano::compress_signal_ll_lh(s, pool, d, rd, false); break;
        case 5: r = pgnano::compress_signal_N02(s, pool, d, rd, false); break;
        case 6: r = pgnano::compress_signal_N01(s, pool, d, rd, false); break;
        default: break;
    }
    if (!r.ok()) { g_err = r.status().ToString(); return 1; }
    *out_size = *r;
    return 0;
}

int ref_cx_decode(int codec, const uint8_t* in, size_t clen, size_t n, int16_t* out)
{
    auto pool = arrow::default_memory_pool();
    pgnano::PGNanoReaderState st;
    gsl::span<const uint8_t> s(in, clen);
    gsl::span<int16_t> d(out, n);
    arrow::Status r = arrow::Status::Invalid("bad codec");
    switch (codec) {
        case 1: r = pgnano::decompress_signal_VBZ1(s, pool, d, st); break;
        case 2: r = pgnano::decompress_signal_KD(s, pool, d, st); break;
        case 3: r = pgnano::decompress_signal_lh(s, pool, d, st); break;
        case 4: r = pgnano::decompress_signal_ll_lh(s, pool, d, st); break;
        case 5: r = pgnano::decompress_signal_N02(s, pool, d, st); break;
        case 6: r = pgnano::decompress_signal_N01(s, pool, d, st); break;
        default: break;
    }
    if (!r.ok()) { g_err = r.ToString(); return 1; }
    return 0;
}

// Pre-zstd streams of the reference's splitters. `streams` must hold 5 regions of (2n+16) bytes each,
// laid out back to back; sizes[0..4] receive the byte count of each stream in wire order.
int ref_cx_streams(int codec, const int16_t* in, size_t n, uint8_t* streams, size_t* sizes)
{
    size_t stride = 2 * n + 16;
    uint8_t* b0 = streams; uint8_t* b1 = b0 + stride; uint8_t* b2 = b1 + stride;
    uint8_t* b3 = b2 + stride; uint8_t* b4 = b3 + stride;
    for (int i = 0; i < 5; ++i) sizes[i] = 0;
    switch (codec) {
        case 1: { auto r = svb16::encode_scalar_VBZ1<int16_t, true, true>(in, b0, b1, (uint32_t)n);
                  sizes[0] = r.keys_sz; sizes[1] = r.data_sz; break; }
        case 2: { auto r = svb16::encode_scalar_KD<int16_t, true, true>(in, b0, b1, (uint32_t)n);
                  sizes[0] = r.keys_sz; sizes[1] = r.data_sz; break; }
        case 3: { auto r = svb16::encode_scalar_lh<int16_t, true, true>(in, b0, b1, b2, (uint32_t)n);
                  sizes[0] = r.keys_sz; sizes[1] = r.data_l_sz; sizes[2] = r.data_h_sz; break; }
        case 4: { auto r = svb16::encode_scalar_ll_lh<int16_t, true, true>(in, b0, b1, b2, b3, (uint32_t)n);
                  sizes[0] = r.keys_sz; sizes[1] = r.data_ll_sz; sizes[2] = r.data_lh_sz; sizes[3] = r.data_h_sz; break; }
        case 5: { auto r = svb16::encode_scalar_N02<int16_t, true, true>(in, b0, b1, b2, b3, b4, (uint32_t)n);
                  sizes[0] = r.keys_sz; sizes[1] = r.data_S_sz; sizes[2] = r.data_M_sz; sizes[3] = r.data_Llow_sz; sizes[4] = r.data_Lhigh_sz; break; }
        case 6: { auto r = svb16::encode_scalar_N01<int16_t, true, true>(in, b0, b1, b2, b3, b4, (uint32_t)n);
                  sizes[0] = r.keys_sz; sizes[1] = r.data_S_sz; sizes[2] = r.data_M_sz; sizes[3] = r.data_Llow_sz; sizes[4] = r.data_Lhigh_sz; break; }
        default: g_err = "bad codec"; return 1;
    }
    return 0;
}

}  // extern "C"
