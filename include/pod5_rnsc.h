/* pod5_rnsc.h -- the per-chunk signal-compression entry points of the reference's POD5 C API, under their ORIGINAL
 * names and error convention, implemented on librnsc (B200). Exported by libpod5_rnsc.so, a thin library separate from
 * librnsc.so so that taking these names over is opt-in at link time: link it (or LD_PRELOAD it) in front of the
 * reference's libpod5_format and benchmark.cpp:145,175, the pybind module and any other caller of these functions run on
 * the GPU codec unchanged.
 *
 * Replaces, in tomas-gr/RawNanoporeSignalCompression:
 *   pod5/c++/pod5_format/c_api.h:31-53    enum pod5_error, pod5_get_error_no, pod5_get_error_string
 *   pod5/c++/pod5_format/c_api.h:678-711  pod5_vbz_compressed_signal_max_size, pod5_vbz_compress_signal,
 *                                         pod5_pinanoraw_compress_signal, pod5_vbz_decompress_signal
 *   (bodies: c_api.cpp:49-58,104-129,1177-1273)
 * Error convention as there: the call returns a pod5_error_t; number and text of the most recent error are kept and read
 * back with pod5_get_error_no() / pod5_get_error_string() ("Invalid: <message>", arrow::Status::ToString()). Every codec
 * error of the reference is arrow::Status::Invalid, i.e. POD5_ERROR_INVALID; a missing CUDA device / CUDA failure, which
 * the reference cannot have, is POD5_ERROR_UNKNOWNERROR. The record is per thread here (process-global and racy in the
 * reference, c_api.cpp:49-58). */
#ifndef POD5_RNSC_H
#define POD5_RNSC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum pod5_error {               /* c_api.h:31-47 (values of arrow::StatusCode) */
    POD5_OK = 0,
    POD5_ERROR_OUTOFMEMORY = 1,
    POD5_ERROR_KEYERROR = 2,
    POD5_ERROR_TYPEERROR = 3,
    POD5_ERROR_INVALID = 4,
    POD5_ERROR_IOERROR = 5,
    POD5_ERROR_CAPACITYERROR = 6,
    POD5_ERROR_INDEXERROR = 7,
    POD5_ERROR_CANCELLED = 8,
    POD5_ERROR_UNKNOWNERROR = 9,
    POD5_ERROR_NOTIMPLEMENTED = 10,
    POD5_ERROR_SERIALIZATIONERROR = 11,
    POD5_ERROR_STRING_NOT_LONG_ENOUGH = 11
};
typedef enum pod5_error pod5_error_t;

pod5_error_t pod5_get_error_no(void);                                                       /* c_api.h:50 */
char const* pod5_get_error_string(void);                                                    /* c_api.h:53 */

size_t pod5_vbz_compressed_signal_max_size(size_t sample_count);                            /* c_api.h:681, c_api.cpp:1177 */
/* *compressed_signal_size: capacity of compressed_signal_out on entry, bytes written on success. */
pod5_error_t pod5_vbz_compress_signal(int16_t const* signal, size_t signal_size, char* compressed_signal_out,
                                      size_t* compressed_signal_size);                      /* c_api.h:688, c_api.cpp:1183 */
/* The pgnano (Cx) encoder; the variant is rnsc_set_pgnano_variant()'s (default C5, the reference's checked-in build). */
pod5_error_t pod5_pinanoraw_compress_signal(int16_t const* signal, size_t signal_size, char* compressed_signal_out,
                                            size_t* compressed_signal_size);                /* c_api.h:699, c_api.cpp:1217 */
pod5_error_t pod5_vbz_decompress_signal(char const* compressed_signal, size_t compressed_signal_size, size_t sample_count,
                                        short* signal_out);                                 /* c_api.h:709, c_api.cpp:1255 */
/* Not in the reference's C API (its only Cx decoder is C++, pgnano::decompress_signal): the inverse of
 * pod5_pinanoraw_compress_signal, same conventions. */
pod5_error_t pod5_pinanoraw_decompress_signal(char const* compressed_signal, size_t compressed_signal_size,
                                              size_t sample_count, short* signal_out);

#ifdef __cplusplus
}
#endif
#endif /* POD5_RNSC_H */
