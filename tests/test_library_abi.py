"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/rnsc.h declares.
No compute calls here (there is no GPU in the build container)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rnsc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rnsc_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from rawnanoporesignalcompression_b200 import _lib
    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(librnsc_path):
    lib = C.CDLL(librnsc_path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"librnsc.so does not export {s}"


def test_sizes_without_device(librnsc_path):
    from rawnanoporesignalcompression_b200 import _lib
    lib = _lib.load(librnsc_path)
    # pgnano/compressor.h:39-45: max(2n+26, 1024)
    assert lib.rnsc_max_compressed_size(_lib.C5, 0) == 1024
    assert lib.rnsc_max_compressed_size(_lib.C5, 102400) == 204826
    # signal_compression.cpp:30-35: ZSTD_compressBound(ceil(n/8) + 2n)
    n = 102400
    src = n // 8 + 2 * n
    assert lib.rnsc_max_compressed_size(_lib.VBZ, n) == src + (src >> 8)
    assert [lib.rnsc_stream_count(c) for c in range(8)] == [2, 2, 2, 3, 4, 5, 5, 0]
    assert lib.rnsc_status_string(4) == b"Remaining data at end of signal buffer"


def test_no_cpu_fallback_without_device(librnsc_path):
    """Creating a context without a CUDA device must fail loudly (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        return
    from rawnanoporesignalcompression_b200 import _lib
    from rawnanoporesignalcompression_b200.codec import Codec
    try:
        Codec(0, lib_path=librnsc_path)
    except _lib.RnscError as e:
        assert e.status == _lib.ERR_NO_DEVICE
    else:
        raise AssertionError("Codec() succeeded without a CUDA device")


def pod5_declared_symbols():
    text = open(os.path.join(ROOT, "include", "pod5_rnsc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pod5_[a-z0-9_]+)\s*\(", text)))


def test_pod5_shim_exports_the_reference_names():
    """libpod5_rnsc.so exports the un-prefixed names of c_api.h:49-53,678-711 and reports errors the reference's way
    (pod5_error_t POD5_ERROR_INVALID = 4, text "Invalid: null passed to C API"; c_api.cpp:104-111) -- no compute here."""
    from rawnanoporesignalcompression_b200 import _lib
    assert pod5_declared_symbols() == sorted(_lib.POD5_EXPORTS)
    lib = _lib.load_pod5()
    for s in pod5_declared_symbols():
        assert hasattr(lib, s), f"libpod5_rnsc.so does not export {s}"
    n = 102400
    src = n // 8 + 2 * n
    assert lib.pod5_vbz_compressed_signal_max_size(n) == src + (src >> 8)
    size = C.c_size_t(16)
    out = C.create_string_buffer(16)
    assert lib.pod5_vbz_compress_signal(None, 4, out, C.byref(size)) == 4
    assert lib.pod5_get_error_no() == 4 and lib.pod5_get_error_string() == b"Invalid: null passed to C API"
    sig = (C.c_int16 * 4)(1, 2, 3, 4)
    assert lib.pod5_pinanoraw_compress_signal(sig, 4, None, C.byref(size)) == 4
    assert lib.pod5_get_error_string() == b"Invalid: null output parameter passed to C API"
    assert lib.pod5_vbz_compressed_signal_max_size(0) >= 0 and lib.pod5_get_error_no() == 0      # every call resets the record (c_api.cpp:1179)
