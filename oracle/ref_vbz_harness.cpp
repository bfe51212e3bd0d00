// oracle/ref_vbz_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the reference's *stock* StreamVByte16 headers (pod5_format/svb16/{encode,decode}.hpp, incl.
// the SSE4.1 decoder decode_x64.hpp) where they lie under /root/reference, and restates the ~20 lines
// of zstd glue of pod5::compress_signal / decompress_signal (signal_compression.cpp:37-78,96-141)
// around them without Arrow. Built into oracle/_ref/libref_oracle.so by oracle/Makefile.
// Both trees define `namespace svb16` with same-named inline functions whose bodies differ (e.g.
// decode_input_buffer_padding_byte_count: 16 here, 0 in the pgnano copy); rename this copy so the two
// translation units cannot be merged by the linker.
#define svb16 svb16_stock
#include "pod5_format/svb16/decode.hpp"
#include "pod5_format/svb16/encode.hpp"
#include <zstd.h>
#include <cstring>
#include <string>
#include <vector>

static thread_local std::string g_err;

extern "C" {

const char* ref_vbz_last_error() { return g_err.c_str(); }

// signal_compression.cpp:30-35
size_t ref_vbz_max_size(size_t n) { return ZSTD_compressBound(svb16_max_encoded_length((uint32_t)n)); }

// svb16::encode<int16,true,true> (encode.hpp:13-24): keys || data into `out` (cap >= ceil(n/8)+2n).
size_t ref_vbz_svb_encode(const int16_t* in, size_t n, uint8_t* out)
{
    return svb16::encode<int16_t, true, true>(in, out, (uint32_t)n);
}

// svb16::decode<int16,true,true> (decode.hpp:25-37); `in` must have 16 readable padding bytes.
size_t ref_vbz_svb_decode(const uint8_t* in, size_t in_size, size_t n, int16_t* out, int force_scalar)
{
    gsl::span<int16_t> o(out, n);
    gsl::span<const uint8_t> i(in, in_size);
    if (force_scalar) {
        auto kl = ::svb16_key_length((uint32_t)n);
        return svb16::decode_scalar<int16_t, true, true>(o, i.subspan(0, kl), i.subspan(kl)) - i.begin();
    }
    return svb16::decode<int16_t, true, true>(o, i);
}

// pod5::compress_signal (signal_compression.cpp:37-78)
int ref_vbz_encode(const int16_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_size)
{
    std::vector<uint8_t> inter(svb16_max_encoded_length((uint32_t)n) + 1);
    size_t enc = svb16::encode<int16_t, true, true>(in, inter.data(), (uint32_t)n);
    size_t bound = ZSTD_compressBound(enc);
    if (ZSTD_isError(bound)) { g_err = "Failed to find zstd max size for data"; return 1; }
    size_t c = ZSTD_compress(out, cap, inter.data(), enc, 1);
    if (ZSTD_isError(c)) { g_err = "Failed to compress data"; return 1; }
    *out_size = c;
    return 0;
}

// pod5::decompress_signal (signal_compression.cpp:96-141)
int ref_vbz_decode(const uint8_t* in, size_t clen, size_t n, int16_t* out)
{
    unsigned long long fcs = ZSTD_getFrameContentSize(in, clen);
    if (ZSTD_isError(fcs)) { g_err = "Input data not compressed by zstd"; return 1; }
    size_t pad = svb16::decode_input_buffer_padding_byte_count();
    std::vector<uint8_t> inter(fcs + pad);
    size_t r = ZSTD_decompress(inter.data(), inter.size(), in, clen);
    if (ZSTD_isError(r)) { g_err = "Input data failed to decompress using zstd"; return 1; }
    gsl::span<int16_t> o(out, n);
    size_t consumed = svb16::decode<int16_t, true, true>(o, gsl::make_span(inter.data(), inter.size()));
    if (consumed + pad != inter.size()) { g_err = "Remaining data at end of signal buffer"; return 1; }
    return 0;
}

unsigned ref_zstd_version() { return ZSTD_versionNumber(); }

}  // extern "C"
