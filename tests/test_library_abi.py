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
