"""Frames from OTHER zstd encoder settings than the reference's `ZSTD_compress(..., 1)` (signal_compression.cpp:65,
C5.hpp:337-407): other levels (different match finders, repeat offsets, treeless literals, FSE table modes), the frame
checksum flag, and contents > 128 KiB (two blocks). The reference decodes all of these because it calls
`ZSTD_decompress` (signal_compression.cpp:117, C5.hpp:588-655); a drop-in decoder has to as well. Test helper only:
uses the image's libzstd.so.1 through ctypes as the frame WRITER."""
import ctypes as C
import struct

import numpy as np

import recipes

_z = C.CDLL("libzstd.so.1")
_z.ZSTD_compressBound.restype = C.c_size_t
_z.ZSTD_createCCtx.restype = C.c_void_p
_z.ZSTD_compress2.restype = C.c_size_t
_z.ZSTD_CCtx_setParameter.restype = C.c_size_t
_z.ZSTD_isError.restype = C.c_uint
_LEVEL, _CHECKSUM = 100, 201  # ZSTD_c_compressionLevel, ZSTD_c_checksumFlag (stable enum values of zstd.h)

SETTINGS = [(1, True), (3, False), (9, True), (19, False), (-5, True), (22, True)]


def frame(src: bytes, level: int, checksum: bool) -> bytes:
    cap = _z.ZSTD_compressBound(C.c_size_t(len(src)))
    dst = C.create_string_buffer(cap)
    cctx = C.c_void_p(_z.ZSTD_createCCtx())
    try:
        assert not _z.ZSTD_isError(C.c_size_t(_z.ZSTD_CCtx_setParameter(cctx, _LEVEL, level)))
        assert not _z.ZSTD_isError(C.c_size_t(_z.ZSTD_CCtx_setParameter(cctx, _CHECKSUM, int(checksum))))
        n = _z.ZSTD_compress2(cctx, dst, C.c_size_t(cap), src, C.c_size_t(len(src)))
        assert not _z.ZSTD_isError(C.c_size_t(n))
    finally:
        _z.ZSTD_freeCCtx(cctx)
    return dst.raw[:n]


def signals():
    """(name, int16 chunk): a smooth walk (Huffman literals, few sequences), a periodic signal (long matches, repeat
    offsets), near-uniform noise at full chunk size (SVB bytes > 128 KiB: two blocks, raw literals), a short chunk."""
    per = np.tile(recipes.walk(4099, 21, 25.0), 24)[:98000].copy()
    per[::977] += 300
    return [("walk", recipes.walk(102400, 3, 12.0)), ("periodic", per.astype(np.int16)),
            ("uniform", recipes.uniform16(102400, 11)), ("short", recipes.walk(257, 5, 40.0))]


def vbz_frame(oracle, x: np.ndarray, level: int, checksum: bool) -> bytes:
    """Stock VBZ chunk = one frame over keys||data (signal_compression.cpp:37-78)."""
    return frame(b"".join(oracle.streams(0, x)), level, checksum)


def c5_container(oracle, x: np.ndarray, level: int, checksum: bool) -> bytes:
    """C5 chunk = [u64 ck][keys][u64 cS][S][u64 cM][M][u64 cLl][Llow][Lhigh], one frame per stream (C5.hpp:282-474)."""
    fr = [frame(s, level, checksum) for s in oracle.streams(6, x)]
    out = b""
    for k, f in enumerate(fr):
        out += (struct.pack("<Q", len(f)) if k < 4 else b"") + f
    return out
