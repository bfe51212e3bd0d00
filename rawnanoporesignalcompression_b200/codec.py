"""Host-side mirror of the reference's per-chunk codec interface over librnsc (CUDA, sm_100a).

Names follow the reference:
  * `compressed_signal_max_size`, `compress_signal`, `decompress_signal`  -- pod5::/pgnano:: C++ functions
    (pod5_format/signal_compression.h:17-36, pgnano/pgnano.h:13-23); the pgnano variant, compile-time in the
    reference (pgnano/pgnano.cpp:1), is the runtime `codec` argument.
  * `vbz_compress_signal`, `vbz_decompress_signal`, `vbz_compress_signal_chunked`,
    `vbz_decompress_signal_chunked` -- the numpy wrappers of pod5's pybind (python/pod5 signal_tools.py:14-131).
  * `Codec.compress_batch` / `decompress_batch` -- what the batch hooks pod5_add_reads_data_pre_compressed /
    extract_samples_inplace need (c_api.h:665-673, signal_table_reader.cpp:320-340).
  * `Codec.dev_*` -- device-resident variants taking torch CUDA tensors (no host copies).
Errors: the reference returns arrow::Status::Invalid(msg); here `RnscError(status, msg)` is raised.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import CODEC_NAMES, RnscError, VBZ, VBZ0, C1, C2, C3, C4, C5, RC  # noqa: F401

DEFAULT_SIGNAL_CHUNK_SIZE = 102_400  # pod5_format/file_writer.h:22


def codec_id(codec) -> int:
    if isinstance(codec, str):
        try:
            return CODEC_NAMES[codec.upper()]
        except KeyError:
            raise ValueError(f"unknown codec {codec!r}; one of {sorted(CODEC_NAMES)}") from None
    c = int(codec)
    if not 0 <= c <= 7:
        raise ValueError(f"unknown codec id {c}")
    return c


def _ptr(a) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


class Codec:
    """One librnsc context (one CUDA device, one stream, reusable staging)."""

    def __init__(self, device: int = -1, lib_path: Optional[str] = None):
        self.lib = _lib.load(lib_path)
        self.ctx = self.lib.rnsc_create(int(device))
        if not self.ctx:
            raise RnscError(_lib.ERR_NO_DEVICE, self.lib.rnsc_last_error().decode())

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.rnsc_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- helpers -------------------------------------------------------------------------------
    def _check(self, r: int):
        if r != 0:
            raise RnscError(r, self.lib.rnsc_last_error().decode() or self.lib.rnsc_status_string(r).decode())

    def status_string(self, s: int) -> str:
        return self.lib.rnsc_status_string(int(s)).decode()

    def max_compressed_size(self, codec, n: int) -> int:
        return int(self.lib.rnsc_max_compressed_size(codec_id(codec), int(n)))

    def set_host_threads(self, n: int):
        self.lib.rnsc_set_host_threads(self.ctx, int(n))

    def set_entropy_stage(self, mode):
        """'libzstd' / 0 (default: host libzstd level 1, bytes identical to the reference) or 'gpu' / 1 (frames written on the
        device: valid zstd, decode parity, no host work)."""
        m = {"libzstd": 0, "gpu": 1}.get(mode, mode) if isinstance(mode, str) else int(mode)
        self._check(self.lib.rnsc_set_entropy_stage(self.ctx, int(m)))

    def decode_class_counts(self, stream=None):
        """(fast 4/SM, fast 2/SM, fast 1/SM, general decoder) chunk counts of the last whole-chunk decode."""
        out = (C.c_uint32 * 4)()
        self._check(self.lib.rnsc_decode_class_counts(self.ctx, out, C.c_void_p(int(stream.cuda_stream)) if stream is not None else C.c_void_p(0)))
        return tuple(int(v) for v in out)

    def set_profiling(self, on: bool):
        self._check(self.lib.rnsc_set_profiling(self.ctx, int(bool(on))))

    def cta_spans(self):
        """(3, 1024, 3) uint64: per size class (S, A, B) and CTA {start ns, end ns, chunks} of the last decode (profiling on)."""
        out = np.zeros(3 * 1024 * 3, dtype=np.uint64)
        self._check(self.lib.rnsc_get_cta_spans(self.ctx, out.ctypes.data))
        return out.reshape(3, 1024, 3)

    def phase_cycles(self):
        """{class: {phase: cycles}} accumulated by the fast decode kernels since set_profiling(True)."""
        out = (C.c_uint64 * 24)()
        self._check(self.lib.rnsc_get_phase_cycles(self.ctx, out))
        names = ["stage_tables", "huf_pass1", "huf_scan", "huf_pass2", "fills_sequences", "unpack", "bookkeeping", "chunks"]
        return {cls: {nm: int(out[8 * k + i]) for i, nm in enumerate(names)} for k, cls in enumerate(("S_256x4", "A_512x2", "B_1024x1"))}

    @property
    def launch_count(self) -> int:
        return int(self.lib.rnsc_launch_count(self.ctx))

    # ---- per chunk -----------------------------------------------------------------------------
    def compress_signal(self, samples: np.ndarray, codec=VBZ) -> bytes:
        cid = codec_id(codec)
        s = np.ascontiguousarray(samples, dtype=np.int16)
        cap = self.max_compressed_size(cid, s.size)
        out = np.empty(cap, dtype=np.uint8)
        size = C.c_size_t(cap)
        self._check(self.lib.rnsc_compress_chunk(self.ctx, cid, _ptr(s), s.size, _ptr(out), C.byref(size)))
        return out[: size.value].tobytes()

    def decompress_signal(self, compressed, sample_count: int, codec=VBZ) -> np.ndarray:
        cid = codec_id(codec)
        b = np.frombuffer(bytes(compressed), dtype=np.uint8) if not isinstance(compressed, np.ndarray) else np.ascontiguousarray(compressed, dtype=np.uint8)
        out = np.empty(int(sample_count), dtype=np.int16)
        self._check(self.lib.rnsc_decompress_chunk(self.ctx, cid, _ptr(b), b.size, int(sample_count), _ptr(out)))
        return out

    # ---- batch, host buffers -------------------------------------------------------------------
    def compress_batch(self, chunks: Sequence[np.ndarray], codec=VBZ) -> Tuple[List[bytes], np.ndarray]:
        """Returns (compressed bytes per chunk, int32 status per chunk)."""
        cid = codec_id(codec)
        n = np.array([len(c) for c in chunks], dtype=np.uint32)
        soff = np.zeros(len(chunks) + 1, dtype=np.uint64)
        np.cumsum(n, out=soff[1:])
        flat = np.concatenate([np.ascontiguousarray(c, dtype=np.int16) for c in chunks]) if len(chunks) else np.zeros(0, np.int16)
        out, coff, status = self.compress_flat(flat, soff[:-1].copy(), n, cid)
        return [out[int(coff[i]) : int(coff[i + 1])].tobytes() for i in range(len(chunks))], status

    def compress_flat(self, samples: np.ndarray, sample_off: np.ndarray, n: np.ndarray, codec=VBZ):
        cid = codec_id(codec)
        samples = np.ascontiguousarray(samples, dtype=np.int16)
        sample_off = np.ascontiguousarray(sample_off, dtype=np.uint64)
        n = np.ascontiguousarray(n, dtype=np.uint32)
        uniq, counts = np.unique(n, return_counts=True)                # one ABI call per distinct chunk length
        cap = int(sum(self.max_compressed_size(cid, int(k)) * int(m) for k, m in zip(uniq, counts))) + 64
        out = np.empty(cap, dtype=np.uint8)
        coff = np.zeros(n.size + 1, dtype=np.uint64)
        status = np.zeros(n.size, dtype=np.int32)
        if samples.size == 0:
            samples = np.zeros(1, np.int16)
        self._check(self.lib.rnsc_compress_batch(self.ctx, cid, _ptr(samples), _ptr(sample_off), _ptr(n), n.size, _ptr(out), cap, _ptr(coff), _ptr(status)))
        return out[: int(coff[-1])], coff, status

    def decompress_batch(self, chunks: Sequence[bytes], sample_counts: Sequence[int], codec=VBZ) -> Tuple[List[np.ndarray], np.ndarray]:
        cid = codec_id(codec)
        n = np.asarray(sample_counts, dtype=np.uint32)
        coff = np.zeros(len(chunks) + 1, dtype=np.uint64)
        np.cumsum([len(c) for c in chunks], out=coff[1:])
        flat = np.frombuffer(b"".join(bytes(c) for c in chunks) + b"\0" * 16, dtype=np.uint8)
        out, soff, status = self.decompress_flat(flat, coff, n, cid)
        return [out[int(soff[i]) : int(soff[i]) + int(n[i])] for i in range(len(chunks))], status

    def decompress_flat(self, comp: np.ndarray, comp_off: np.ndarray, n: np.ndarray, codec=VBZ):
        cid = codec_id(codec)
        comp = np.ascontiguousarray(comp, dtype=np.uint8)
        comp_off = np.ascontiguousarray(comp_off, dtype=np.uint64)
        n = np.ascontiguousarray(n, dtype=np.uint32)
        soff = np.zeros(n.size + 1, dtype=np.uint64)
        np.cumsum(n, out=soff[1:])
        out = np.empty(max(int(soff[-1]), 1), dtype=np.int16)
        status = np.zeros(n.size, dtype=np.int32)
        self._check(self.lib.rnsc_decompress_batch(self.ctx, cid, _ptr(comp), _ptr(comp_off), _ptr(n), n.size, _ptr(out), _ptr(soff), _ptr(status)))
        return out[: int(soff[-1])], soff, status

    # ---- device resident (torch CUDA tensors) --------------------------------------------------
    @staticmethod
    def _dp(t) -> C.c_void_p:
        return C.c_void_p(int(t.data_ptr())) if t is not None else C.c_void_p(0)

    @staticmethod
    def _stream(stream) -> C.c_void_p:
        if stream is None:
            import torch
            stream = torch.cuda.current_stream()
        return C.c_void_p(int(stream.cuda_stream))

    def dev_decode(self, codec, comp, comp_off, n, sample_off, max_n: int, samples, status, stream=None):
        """comp:u8, comp_off:i64/u64 [nchunks+1], n:u32 (stored as int32 tensor), sample_off:i64 [>=nchunks],
        samples:int16 out, status:int32 out. Asynchronous on `stream` (default: torch's current stream)."""
        self._check(self.lib.rnsc_dev_decode(self.ctx, codec_id(codec), self._dp(comp), self._dp(comp_off), self._dp(n), self._dp(sample_off),
                                             int(status.numel()), int(max_n), self._dp(samples), self._dp(status), self._stream(stream)))

    def dev_encode(self, codec, samples, sample_off, n, total_samples: int, comp, comp_off, status, stream=None):
        """Whole-chunk encode with the GPU entropy stage. samples:int16, sample_off:i64, n:int32 [nchunks], comp:u8 out,
        comp_off:i64 out [nchunks+1], status:int32 out. Asynchronous on `stream`."""
        self._check(self.lib.rnsc_dev_encode(self.ctx, codec_id(codec), self._dp(samples), self._dp(sample_off), self._dp(n), int(n.numel()), int(total_samples),
                                             self._dp(comp), int(comp.numel()), self._dp(comp_off), self._dp(status), self._stream(stream)))

    def dev_svb_encode(self, codec, samples, sample_off, n, streams, stream_off, stream=None):
        self._check(self.lib.rnsc_dev_svb_encode(self.ctx, codec_id(codec), self._dp(samples), self._dp(sample_off), self._dp(n), int(n.numel()),
                                                 self._dp(streams), int(streams.numel()), self._dp(stream_off), self._stream(stream)))

    def dev_svb_encode_slabs(self, codec, samples, sample_off, n, total_samples: int, streams, stream_start, stream_size, stream=None):
        """One-pass splitter: stream k of chunk i = stream_size[5i+k] bytes at streams[stream_start[5i+k]] (per-chunk slabs)."""
        self._check(self.lib.rnsc_dev_svb_encode_slabs(self.ctx, codec_id(codec), self._dp(samples), self._dp(sample_off), self._dp(n), int(n.numel()), int(total_samples),
                                                       self._dp(streams), int(streams.numel()), self._dp(stream_start), self._dp(stream_size), self._stream(stream)))

    def dev_svb_decode(self, codec, streams, stream_off, n, sample_off, samples, status, stream=None):
        self._check(self.lib.rnsc_dev_svb_decode(self.ctx, codec_id(codec), self._dp(streams), self._dp(stream_off), self._dp(n), self._dp(sample_off),
                                                 int(n.numel()), self._dp(samples), self._dp(status), self._stream(stream)))

    def dev_rc_encode(self, samples, sample_off, n, pore_type: int, comp, slot_off, comp_size, status, stream=None):
        self._check(self.lib.rnsc_dev_rc_encode(self.ctx, self._dp(samples), self._dp(sample_off), self._dp(n), int(n.numel()), int(pore_type),
                                                self._dp(comp), self._dp(slot_off), self._dp(comp_size), self._dp(status), self._stream(stream)))

    def dev_rc_decode(self, comp, comp_off, n, sample_off, samples, status, stream=None):
        self._check(self.lib.rnsc_dev_rc_decode(self.ctx, self._dp(comp), self._dp(comp_off), self._dp(n), self._dp(sample_off), int(n.numel()),
                                                self._dp(samples), self._dp(status), self._stream(stream)))

    def dev_exclusive_scan_u32(self, values, out, stream=None):
        self._check(self.lib.rnsc_dev_exclusive_scan_u32(self.ctx, self._dp(values), int(values.numel()), self._dp(out), self._stream(stream)))


# ---- module-level convenience mirroring pod5's Python helpers (signal_tools.py:14-131) ---------------
_default: Optional[Codec] = None


def default_codec() -> Codec:
    global _default
    if _default is None:
        _default = Codec()
    return _default


def compressed_signal_max_size(sample_count: int, codec=VBZ) -> int:
    return default_codec().max_compressed_size(codec, sample_count)


def compress_signal(samples: np.ndarray, codec=VBZ) -> bytes:
    return default_codec().compress_signal(samples, codec)


def decompress_signal(compressed, sample_count: int, codec=VBZ) -> np.ndarray:
    return default_codec().decompress_signal(compressed, sample_count, codec)


def vbz_compress_signal(signal: np.ndarray) -> bytes:
    return compress_signal(signal, VBZ)


def vbz_decompress_signal(compressed, sample_count: int) -> np.ndarray:
    return decompress_signal(compressed, sample_count, VBZ)


def vbz_compress_signal_chunked(signal: np.ndarray, signal_chunk_size: int = DEFAULT_SIGNAL_CHUNK_SIZE):
    """(list of compressed chunks, list of chunk sample counts) -- signal_tools.py:89-131."""
    signal = np.ascontiguousarray(signal, dtype=np.int16)
    pieces = [signal[i : i + signal_chunk_size] for i in range(0, len(signal), signal_chunk_size)]
    comp, status = default_codec().compress_batch(pieces, VBZ)
    if status.any():
        raise RnscError(int(status[status != 0][0]), "chunk failed to compress")
    return comp, [len(p) for p in pieces]


def vbz_decompress_signal_chunked(compressed_chunks: Sequence[bytes], sample_counts: Sequence[int]) -> np.ndarray:
    out, status = default_codec().decompress_batch(compressed_chunks, sample_counts, VBZ)
    if status.any():
        raise RnscError(int(status[status != 0][0]), "chunk failed to decompress")
    return np.concatenate(out) if out else np.zeros(0, np.int16)
