// Build shim (test infrastructure): the container ships libzstd.so.1 (1.5.5) without headers.
// Prototypes of the stable one-shot API, as published in zstd's public documentation.
#pragma once
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
size_t ZSTD_compress(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int compressionLevel);
size_t ZSTD_decompress(void* dst, size_t dstCapacity, const void* src, size_t compressedSize);
unsigned long long ZSTD_getFrameContentSize(const void* src, size_t srcSize);
size_t ZSTD_compressBound(size_t srcSize);
unsigned ZSTD_isError(size_t code);
const char* ZSTD_getErrorName(size_t code);
unsigned ZSTD_versionNumber(void);
const char* ZSTD_versionString(void);
#ifdef __cplusplus
}
#endif
