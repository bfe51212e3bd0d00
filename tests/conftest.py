import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_binding import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref_oracle():
    from oracle_binding import RefOracle
    if not RefOracle.available():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    return RefOracle()


@pytest.fixture(scope="session")
def enc_oracle(oracle):
    """The checker of the GPU parity tests: the reference's own code (oracle/_ref, prebuilt and shipped to the GPU box)
    when it exists, else the C port that is pinned to it by the CPU suite."""
    from oracle_binding import RefOracle
    return RefOracle() if RefOracle.available() else oracle


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def real_chunks():
    """(stored VBZ frame bytes, sample count) of the reference's pod5 fixture (committed copy)."""
    from rawnanoporesignalcompression_b200.pod5_container import read_pod5
    sig = read_pod5(os.path.join(ROOT, "tests", "golden", "multi_fast5_zip_v3.pod5")).signal
    return list(zip(sig.column("signal").to_pylist(), sig.column("samples").to_pylist()))


@pytest.fixture(scope="session")
def librnsc_path():
    """Builds (nvcc cross-compiles without a GPU) or finds the in-tree CUDA library."""
    import __graft_entry__ as g
    return g.build_cuda()


@pytest.fixture(scope="session")
def emul_lib():
    """The kernel sources compiled with g++ against tests/emul (CUDA-on-CPU fibers). Test infrastructure."""
    out_dir = os.path.join(ROOT, "tests", "emul", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "librnsc_emul.so")
    csrc = os.path.join(ROOT, "rawnanoporesignalcompression_b200", "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(ROOT, "tests", "emul", f) for f in ("cuda_emul.h", "cuda_emul.cpp")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-DRNSC_EMUL_BUILD", "-x", "c++",
                        "-I" + os.path.join(ROOT, "tests", "emul"), "-I" + os.path.join(ROOT, "include"),
                        os.path.join(csrc, "api.cu"), os.path.join(ROOT, "tests", "emul", "cuda_emul.cpp"),
                        "-o", so, "-ldl", "-lpthread"], check=True)
    return so


@pytest.fixture(scope="session")
def gpu_codec(librnsc_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rawnanoporesignalcompression_b200.codec import Codec
    cd = Codec(0)
    yield cd
    cd.close()
