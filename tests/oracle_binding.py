"""ctypes binding of the CPU oracle (oracle/liboracle.so) and, when built, of the reference's own code
(oracle/_ref/libref_oracle.so). TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_oracle.so")

_P = lambda a: C.c_void_p(a.ctypes.data)
_S = C.c_size_t


def ensure_built():
    if not os.path.exists(ORACLE_SO):
        subprocess.run(["make", "-f", os.path.join(ROOT, "oracle", "Makefile"), ORACLE_SO], check=True, cwd=ROOT)


class Oracle:
    """Plain-C restatement (oracle/rnsc_oracle.c, zstd_decode_oracle.c)."""

    def __init__(self):
        ensure_built()
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.orc_max_size.restype = C.c_size_t
        self.lib.orc_rc_encode.restype = C.c_size_t
        self.lib.orc_zstd_frame_content_size.restype = C.c_uint64
        self.lib.orc_zstd_version.restype = C.c_uint

    def zstd_version(self) -> int:
        return int(self.lib.orc_zstd_version())

    def max_size(self, codec: int, n: int) -> int:
        return int(self.lib.orc_max_size(codec, _S(n)))

    def encode_status(self, codec: int, x: np.ndarray):
        x = np.ascontiguousarray(x, dtype=np.int16)
        n = x.size
        cap = self.max_size(codec, n)
        out = np.zeros(cap + 64, np.uint8)
        if codec == 7:
            m = self.lib.orc_rc_encode(_P(x), _S(n), _P(out), C.c_uint8(0))
            return (0 if m else 7), out[:m].tobytes()
        sz = _S(0)
        r = self.lib.orc_encode(codec, _P(x), _S(n), _P(out), _S(cap), C.byref(sz))
        return r, out[: sz.value].tobytes()

    def encode(self, codec: int, x: np.ndarray) -> bytes:
        r, b = self.encode_status(codec, x)
        if r:
            raise RuntimeError(f"oracle encode status {r}")
        return b

    def decode_status(self, codec: int, b: bytes, n: int, own_zstd: bool = False):
        out = np.zeros(max(n, 1), np.int16)
        bb = np.frombuffer(bytes(b) + b"\0" * 16, np.uint8)
        if codec == 7:
            r = self.lib.orc_rc_decode(_P(bb), _S(len(b)), _S(n), _P(out))
        else:
            r = self.lib.orc_decode(codec, _P(bb), _S(len(b)), _S(n), _P(out), int(own_zstd))
        return r, out[:n]

    def decode(self, codec: int, b: bytes, n: int, own_zstd: bool = False) -> np.ndarray:
        r, o = self.decode_status(codec, b, n, own_zstd)
        if r:
            raise RuntimeError(f"oracle decode status {r}")
        return o

    def streams(self, codec: int, x: np.ndarray):
        x = np.ascontiguousarray(x, dtype=np.int16)
        n = x.size
        stride = 2 * n + 16
        buf = np.zeros(5 * stride + 16, np.uint8)
        sizes = (C.c_size_t * 5)()
        assert self.lib.orc_streams(codec, _P(x), _S(n), _P(buf), sizes) == 0
        return [buf[k * stride : k * stride + sizes[k]].tobytes() for k in range(5)]

    def zstd_decode(self, frame: bytes) -> bytes:
        fb = np.frombuffer(bytes(frame) + b"\0" * 16, np.uint8)
        fcs = int(self.lib.orc_zstd_frame_content_size(_P(fb), _S(len(frame))))
        if fcs >= (1 << 64) - 2:
            raise RuntimeError("not a zstd frame / unknown content size")
        out = np.zeros(fcs + 16, np.uint8)
        got = _S(0)
        r = self.lib.orc_zstd_decode_frame(_P(fb), _S(len(frame)), _P(out), _S(fcs), C.byref(got))
        if r:
            raise RuntimeError(f"oracle zstd decode status {r}")
        return out[: got.value].tobytes()


class RefOracle:
    """The reference's own codec sources compiled from /root/reference (oracle/_ref). Optional."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def __init__(self):
        self.lib = C.CDLL(REF_SO)
        self.lib.ref_vbz_max_size.restype = C.c_size_t
        self.lib.ref_rc_encode.restype = C.c_size_t
        self.lib.ref_vbz_svb_encode.restype = C.c_size_t
        self.lib.ref_vbz_svb_decode.restype = C.c_size_t

    def encode_status(self, codec: int, x: np.ndarray):
        x = np.ascontiguousarray(x, dtype=np.int16)
        n = x.size
        cap = int(self.lib.ref_vbz_max_size(_S(n))) if codec == 0 else max(2 * n + 26, 1024)
        out = np.zeros(4 * n + 4096, np.uint8)   # the reference overruns its capacity for C1-C3
        sz = _S(0)
        if codec == 0:
            r = self.lib.ref_vbz_encode(_P(x), _S(n), _P(out), _S(cap), C.byref(sz))
            r = 6 if r else 0
        elif codec == 7:
            m = self.lib.ref_rc_encode(_P(x), _S(n), _P(out), b"R10.4.1")
            return (0 if m else 7), out[:m].tobytes()
        else:
            r = self.lib.ref_cx_encode(codec, _P(x), _S(n), _P(out), _S(cap), C.byref(sz))
            if r:
                r = 3 if codec in (5, 6) else 6
            elif sz.value > cap:
                r = 3
        return r, out[: sz.value].tobytes()

    def encode(self, codec: int, x: np.ndarray) -> bytes:
        r, b = self.encode_status(codec, x)
        if r:
            raise RuntimeError(f"reference encode status {r}")
        return b

    def decode_status(self, codec: int, b: bytes, n: int):
        out = np.zeros(max(n, 1), np.int16)
        bb = np.frombuffer(bytes(b) + b"\0" * 16, np.uint8)
        if codec == 0:
            r = self.lib.ref_vbz_decode(_P(bb), _S(len(b)), _S(n), _P(out))
        elif codec == 7:
            self.lib.ref_rc_decode(_P(bb), _S(len(b)), _P(out)); r = 0
        else:
            r = self.lib.ref_cx_decode(codec, _P(bb), _S(len(b)), _S(n), _P(out))
        return r, out[:n]
