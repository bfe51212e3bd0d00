"""The drop-in seams of INTEGRATION.md on the GPU: the reference's C API names (libpod5_rnsc.so, include/pod5_rnsc.h) and,
when integration/_build was built here, the C++ functions + the reference's per-chunk benchmark loop (seam_bench)."""
import ctypes as C
import json
import os
import subprocess
import threading

import numpy as np
import pytest

import recipes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pod5lib():
    from rawnanoporesignalcompression_b200 import _lib
    return _lib.load_pod5()


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def test_c_api_names_roundtrip_and_bytes(pod5lib, ref_oracle):
    """pod5_vbz_compress_signal / pod5_pinanoraw_compress_signal produce the reference's bytes; pod5_vbz_decompress_signal
    and the added pod5_pinanoraw_decompress_signal return the samples; *size is capacity in, bytes out (c_api.cpp:1183-1253)."""
    from rawnanoporesignalcompression_b200 import synth
    x = synth.generate_reads(np.array([70000]), 5, synth.R10_4_1).numpy()
    for fn_c, fn_d, codec in ((pod5lib.pod5_vbz_compress_signal, pod5lib.pod5_vbz_decompress_signal, 0),
                              (pod5lib.pod5_pinanoraw_compress_signal, pod5lib.pod5_pinanoraw_decompress_signal, 6)):
        cap = 4 * x.size
        out = np.zeros(cap, np.uint8); size = C.c_size_t(cap)
        assert fn_c(_ptr(x), x.size, _ptr(out), C.byref(size)) == 0, pod5lib.pod5_get_error_string()
        st, want = ref_oracle.encode_status(codec, x)
        assert st == 0 and out[: size.value].tobytes() == want
        back = np.zeros(x.size, np.int16)
        assert fn_d(_ptr(out), size.value, x.size, _ptr(back)) == 0
        assert (back == x).all()
        # capacity too small: POD5_ERROR_INVALID and the reference's message (c_api.cpp:1201-1209)
        small = C.c_size_t(16)
        assert fn_c(_ptr(x), x.size, _ptr(out), C.byref(small)) == 4
        assert pod5lib.pod5_get_error_no() == 4 and pod5lib.pod5_get_error_string().startswith(b"Invalid: Compressed signal size")
        # wrong sample count: "Remaining data at end of signal buffer" (signal_compression.cpp:138)
        assert fn_d(_ptr(out), size.value, x.size - 8, _ptr(back)) == 4
        assert pod5lib.pod5_get_error_string() == b"Invalid: Remaining data at end of signal buffer"


def test_c_api_concurrent_callers_use_their_own_contexts(pod5lib):
    """AsyncSignalLoader-style workers (async_signal_loader.cpp:176-220): several threads decode at once, each on its own
    librnsc context; results stay correct and the error record is per thread."""
    xs = [recipes.walk(30000 + 1000 * i, 100 + i, 12.0) for i in range(6)]
    comp = []
    for x in xs:
        out = np.zeros(4 * x.size, np.uint8); size = C.c_size_t(out.size)
        assert pod5lib.pod5_pinanoraw_compress_signal(_ptr(x), x.size, _ptr(out), C.byref(size)) == 0
        comp.append(out[: size.value].copy())
    bad = []

    def worker(i):
        for _ in range(5):
            back = np.zeros(xs[i].size, np.int16)
            r = pod5lib.pod5_pinanoraw_decompress_signal(_ptr(comp[i]), comp[i].size, xs[i].size, _ptr(back))
            if r != 0 or not (back == xs[i]).all():
                bad.append(i)

    th = [threading.Thread(target=worker, args=(i,)) for i in range(len(xs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not bad


def test_cxx_seam_and_per_chunk_benchmark_loop(tmp_path):
    """integration/_build/seam_bench = the reference's benchmark.cpp:134-190 loop through pod5::/pgnano::compress_signal,
    decompress_signal (C++ seam, arrow buffers) and through the C API names; every chunk must round-trip."""
    exe = os.path.join(ROOT, "integration", "_build", "seam_bench")
    if not os.path.exists(exe):
        pytest.skip("integration/_build not built (needs the reference's gsl-lite + pyarrow at build time)")
    from rawnanoporesignalcompression_b200 import synth
    lengths = synth.read_lengths(12, 20261018)
    sig = synth.generate_reads(lengths, 20261018, synth.R10_4_1).numpy()
    off, n, _ = synth.chunk_table(lengths)
    path = str(tmp_path / "chunks.bin")
    synth.write_chunk_file(path, sig, off, n)
    r = subprocess.run([exe, path, "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    res = json.loads(r.stdout)
    for leg in ("cxx_pgnano", "cxx_vbz", "c_api_pinanoraw", "c_api_vbz", "c_api_pinanoraw_concurrent"):
        assert res[leg]["ok"], (leg, res[leg])
    assert res["cxx_pgnano"]["chunks"] == len(n) and res["cxx_pgnano"]["samples"] == int(n.sum())
    assert abs(res["cxx_pgnano"]["bits_per_symbol_mean"] - res["c_api_pinanoraw"]["bits_per_symbol_mean"]) < 1e-9
