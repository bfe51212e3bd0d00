// pod5_shim.cpp -- libpod5_rnsc.so: the reference's per-chunk C API names (include/pod5_rnsc.h) over librnsc.
// Plain C++ (no CUDA in this file): built with g++ and linked against librnsc.so next to it.
#include "../../include/pod5_rnsc.h"
#include "../../include/rnsc.h"

#include <string>

namespace {

thread_local pod5_error_t t_errno = POD5_OK;
thread_local std::string t_errstr;

void reset_error() { t_errno = POD5_OK; t_errstr.clear(); }

// arrow::Status::ToString() of the reference: "<code name>: <message>" (c_api.cpp:52-56)
pod5_error_t set_error(pod5_error_t code, const std::string& msg)
{
    t_errno = code;
    t_errstr = (code == POD5_ERROR_INVALID ? "Invalid: " : "Unknown error: ") + msg;
    return code;
}

pod5_error_t from_rnsc(int r)
{
    if (r == RNSC_OK) return POD5_OK;
    const char* last = rnsc_last_error();
    const std::string msg = (last && *last) ? last : rnsc_status_string(r);
    return set_error((r == RNSC_ERR_NO_DEVICE || r == RNSC_ERR_NO_ZSTD) ? POD5_ERROR_UNKNOWNERROR : POD5_ERROR_INVALID, msg);
}

bool check_not_null(const void* p) { if (!p) { set_error(POD5_ERROR_INVALID, "null passed to C API"); return false; } return true; }                       // c_api.cpp:104-111
bool check_output_not_null(const void* p) { if (!p) { set_error(POD5_ERROR_INVALID, "null output parameter passed to C API"); return false; } return true; } // c_api.cpp:122-129

pod5_error_t compress(int codec, int16_t const* signal, size_t n, char* out, size_t* size)
{
    reset_error();
    if (!check_not_null(signal) || !check_output_not_null(out) || !check_output_not_null(size)) return t_errno;
    rnsc_ctx* c = rnsc_thread_context();
    if (!c) return from_rnsc(RNSC_ERR_NO_DEVICE);
    return from_rnsc(rnsc_compress_chunk(c, codec, signal, n, (uint8_t*)out, size));
}

pod5_error_t decompress(int codec, char const* in, size_t in_size, size_t n, short* out)
{
    reset_error();
    if (!check_not_null(in) || !check_output_not_null(out)) return t_errno;
    rnsc_ctx* c = rnsc_thread_context();
    if (!c) return from_rnsc(RNSC_ERR_NO_DEVICE);
    return from_rnsc(rnsc_decompress_chunk(c, codec, (const uint8_t*)in, in_size, n, out));
}

}  // namespace

extern "C" {

pod5_error_t pod5_get_error_no(void) { return t_errno; }
char const* pod5_get_error_string(void) { return t_errstr.c_str(); }

size_t pod5_vbz_compressed_signal_max_size(size_t sample_count) { reset_error(); return rnsc_max_compressed_size(RNSC_VBZ, sample_count); }

pod5_error_t pod5_vbz_compress_signal(int16_t const* signal, size_t signal_size, char* out, size_t* size) { return compress(RNSC_VBZ, signal, signal_size, out, size); }
pod5_error_t pod5_pinanoraw_compress_signal(int16_t const* signal, size_t signal_size, char* out, size_t* size) { return compress(rnsc_get_pgnano_variant(), signal, signal_size, out, size); }
pod5_error_t pod5_vbz_decompress_signal(char const* in, size_t in_size, size_t sample_count, short* out) { return decompress(RNSC_VBZ, in, in_size, sample_count, out); }
pod5_error_t pod5_pinanoraw_decompress_signal(char const* in, size_t in_size, size_t sample_count, short* out) { return decompress(rnsc_get_pgnano_variant(), in, in_size, sample_count, out); }

}  // extern "C"
