"""ctypes binding of librnsc.so (the C ABI of include/rnsc.h).

The shared object is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a). There is no CPU fallback:
if the library is missing, or no CUDA device is visible when a context is created, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "librnsc.so")

VBZ, VBZ0, C1, C2, C3, C4, C5, RC = range(8)
CODEC_NAMES = {"VBZ": VBZ, "VBZ1": VBZ, "VBZ0": VBZ0, "C1": C1, "C2": C2, "C3": C3, "C4": C4, "C5": C5, "RC": RC}
CODEC_LABEL = {v: k for k, v in CODEC_NAMES.items() if k != "VBZ1"}

OK = 0
ERR_NOT_ZSTD, ERR_ZSTD_DECODE, ERR_DST_TOO_SMALL, ERR_REMAINING, ERR_CONTAINER = 1, 2, 3, 4, 5
ERR_ZSTD_ENCODE, ERR_RC_RAW, ERR_BAD_ARG, ERR_NO_DEVICE, ERR_NO_ZSTD = 6, 7, 8, 9, 10

# every symbol include/rnsc.h declares (tests check that the built library exports all of them)
EXPORTS = [
    "rnsc_version", "rnsc_device_count", "rnsc_status_string", "rnsc_last_error", "rnsc_create", "rnsc_destroy",
    "rnsc_set_host_threads", "rnsc_max_compressed_size", "rnsc_max_stream_bytes", "rnsc_stream_count",
    "rnsc_compress_chunk", "rnsc_decompress_chunk", "rnsc_compress_batch", "rnsc_decompress_batch",
    "rnsc_dev_decode", "rnsc_dev_svb_encode", "rnsc_dev_svb_encode_slabs", "rnsc_svb_slab_bytes", "rnsc_dev_svb_decode", "rnsc_dev_rc_encode", "rnsc_dev_rc_decode",
    "rnsc_dev_exclusive_scan_u32", "rnsc_launch_count", "rnsc_decode_class_counts", "rnsc_set_profiling", "rnsc_get_phase_cycles", "rnsc_get_cta_spans",
    "rnsc_pod5_vbz_compressed_signal_max_size", "rnsc_pod5_vbz_compress_signal", "rnsc_pod5_pinanoraw_compress_signal",
    "rnsc_pod5_vbz_decompress_signal", "rnsc_pod5_pinanoraw_decompress_signal", "rnsc_set_pgnano_variant",
    "rnsc_get_pgnano_variant", "rnsc_thread_context", "rnsc_set_entropy_stage", "rnsc_get_entropy_stage", "rnsc_dev_encode",
]
# libpod5_rnsc.so (include/pod5_rnsc.h): the reference's own C API names for this path
POD5_EXPORTS = [
    "pod5_get_error_no", "pod5_get_error_string", "pod5_vbz_compressed_signal_max_size", "pod5_vbz_compress_signal",
    "pod5_pinanoraw_compress_signal", "pod5_vbz_decompress_signal", "pod5_pinanoraw_decompress_signal",
]


class RnscError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"rnsc status {status}: {message}")
        self.status = status


def load(path: Optional[str] = None) -> C.CDLL:
    path = path or DEFAULT_LIB
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). librnsc has no CPU fallback.")
    lib = C.CDLL(path)
    vp, sz, i32, u8, u32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint8, C.c_uint32, C.c_uint64
    sig = {
        "rnsc_version": (C.c_char_p, []),
        "rnsc_device_count": (i32, []),
        "rnsc_status_string": (C.c_char_p, [i32]),
        "rnsc_last_error": (C.c_char_p, []),
        "rnsc_create": (vp, [i32]),
        "rnsc_destroy": (None, [vp]),
        "rnsc_set_host_threads": (None, [vp, i32]),
        "rnsc_max_compressed_size": (sz, [i32, sz]),
        "rnsc_max_stream_bytes": (sz, [i32, sz]),
        "rnsc_stream_count": (i32, [i32]),
        "rnsc_compress_chunk": (i32, [vp, i32, vp, sz, vp, C.POINTER(sz)]),
        "rnsc_decompress_chunk": (i32, [vp, i32, vp, sz, sz, vp]),
        "rnsc_compress_batch": (i32, [vp, i32, vp, vp, vp, sz, vp, sz, vp, vp]),
        "rnsc_decompress_batch": (i32, [vp, i32, vp, vp, vp, sz, vp, vp, vp]),
        "rnsc_dev_decode": (i32, [vp, i32, vp, vp, vp, vp, sz, u32, vp, vp, vp]),
        "rnsc_dev_svb_encode": (i32, [vp, i32, vp, vp, vp, sz, vp, sz, vp, vp]),
        "rnsc_dev_svb_encode_slabs": (i32, [vp, i32, vp, vp, vp, sz, sz, vp, sz, vp, vp, vp]),
        "rnsc_svb_slab_bytes": (sz, [i32, u32]),
        "rnsc_dev_svb_decode": (i32, [vp, i32, vp, vp, vp, vp, sz, vp, vp, vp]),
        "rnsc_dev_rc_encode": (i32, [vp, vp, vp, vp, sz, u8, vp, vp, vp, vp, vp]),
        "rnsc_dev_rc_decode": (i32, [vp, vp, vp, vp, vp, sz, vp, vp, vp]),
        "rnsc_dev_exclusive_scan_u32": (i32, [vp, vp, sz, vp, vp]),
        "rnsc_launch_count": (u64, [vp]),
        "rnsc_decode_class_counts": (i32, [vp, vp, vp]),
        "rnsc_set_profiling": (i32, [vp, i32]),
        "rnsc_get_phase_cycles": (i32, [vp, vp]),
        "rnsc_get_cta_spans": (i32, [vp, vp]),
        "rnsc_pod5_vbz_compressed_signal_max_size": (sz, [sz]),
        "rnsc_pod5_vbz_compress_signal": (i32, [vp, sz, vp, C.POINTER(sz)]),
        "rnsc_pod5_pinanoraw_compress_signal": (i32, [vp, sz, vp, C.POINTER(sz)]),
        "rnsc_pod5_vbz_decompress_signal": (i32, [vp, sz, sz, vp]),
        "rnsc_pod5_pinanoraw_decompress_signal": (i32, [vp, sz, sz, vp]),
        "rnsc_set_pgnano_variant": (i32, [i32]),
        "rnsc_get_pgnano_variant": (i32, []),
        "rnsc_thread_context": (vp, []),
        "rnsc_set_entropy_stage": (i32, [vp, i32]),
        "rnsc_get_entropy_stage": (i32, [vp]),
        "rnsc_dev_encode": (i32, [vp, i32, vp, vp, vp, sz, sz, vp, sz, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def load_pod5(path: Optional[str] = None) -> C.CDLL:
    """libpod5_rnsc.so: pod5_vbz_compress_signal & co. under the reference's names and pod5_error_t convention."""
    path = path or os.path.join(os.path.dirname(DEFAULT_LIB), "libpod5_rnsc.so")
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    load()                                     # librnsc.so first (same directory; libpod5_rnsc.so finds it through $ORIGIN)
    lib = C.CDLL(path)
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    for name, (res, args) in {
        "pod5_get_error_no": (i32, []), "pod5_get_error_string": (C.c_char_p, []),
        "pod5_vbz_compressed_signal_max_size": (sz, [sz]),
        "pod5_vbz_compress_signal": (i32, [vp, sz, vp, C.POINTER(sz)]), "pod5_pinanoraw_compress_signal": (i32, [vp, sz, vp, C.POINTER(sz)]),
        "pod5_vbz_decompress_signal": (i32, [vp, sz, sz, vp]), "pod5_pinanoraw_decompress_signal": (i32, [vp, sz, sz, vp]),
    }.items():
        fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
    return lib
