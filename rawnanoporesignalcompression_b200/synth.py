"""Synthetic nanopore-shaped int16 signal; new code, the reference ships none.

Signal model "levels + noise": per read, a sequence of events; event k has an i.i.d. level
l_k ~ mu + sigma_L * g and a dwell ~ 1 + Geometric(mean = f_s / bps); sample = clip(l_k + sigma_n * g',
-4096, 4095) with rare (p ~ 1e-4) +-(500..3000) spikes that populate the >= 273 zigzag class.

Everything is integer arithmetic on a counter-based hash (SplitMix64 finaliser keyed by
(seed, read index, sample/event index)), written with torch int64 ops only, so the same function gives
bit-identical samples on the CPU and on any GPU and every rank can regenerate exactly its own shard.

Chunking follows the reference writer: reads are cut into consecutive pieces of `chunk_size` = 102,400
samples, last piece shorter (pod5/c++/pod5_format/file_writer.cpp:119-143, file_writer.h:22).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

DEFAULT_CHUNK_SIZE = 102_400  # file_writer.h:22

_M1 = -0x61C8864680B583EB  # 0x9E3779B97F4A7C15 as int64
_M2 = -0x40A7B892E31B1A47  # 0xBF58476D1CE4E5B9
_M3 = -0x6B2FB644ECCEEE15  # 0x94D049BB133111EB


@dataclass(frozen=True)
class PoreModel:
    name: str
    sample_rate: int      # Hz
    bases_per_s: int
    mu: int               # mean level (ADC counts)
    sigma_level_q: int    # level sigma, Q10 fixed point
    sigma_noise_q: int    # noise sigma, Q10 fixed point
    spike_p_q32: int      # spike probability * 2^32

    @property
    def event_p_q32(self) -> int:
        # P(new event at a sample) = bps / f_s  (mean dwell f_s / bps)
        return int((self.bases_per_s << 32) // self.sample_rate)


# sigma's calibrated against the real R9.4.1 test file (class mix 2.4/35.5/61.5/0.6 %, median |delta| 12,
# VBZ 6.75 bits/sample), then frozen: R10.4.1 gives 2.3/34.6/62.5/0.7 %, |delta| 12, VBZ 6.66 b/sample.
R10_4_1 = PoreModel("R10.4.1", 5000, 400, 380, 55 << 10, int(11.5 * 1024), int(1e-4 * (1 << 32)))
R9_4_1 = PoreModel("R9.4.1", 4000, 450, 380, 55 << 10, int(12.0 * 1024), int(1e-4 * (1 << 32)))
MODELS = {"R10.4.1": R10_4_1, "R9.4.1": R9_4_1}


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    """Logical shift right of int64 (torch's >> is arithmetic)."""
    return (x >> s) & ((1 << (64 - s)) - 1)


def _mix(x: torch.Tensor) -> torch.Tensor:
    """SplitMix64 finaliser on int64 tensors (two's-complement wrap-around == uint64 arithmetic)."""
    x = (x ^ _lsr(x, 30)) * _M2
    x = (x ^ _lsr(x, 27)) * _M3
    return x ^ _lsr(x, 31)


def _hash(seed: int, stream: int, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    k = (seed * 0x2545F4914F6CDD1D + stream * 0x9E3779B97F4A7C15) & ((1 << 64) - 1)
    if k >= 1 << 63:
        k -= 1 << 64
    return _mix(_mix(a * _M1 + k) ^ (b * _M2))


def _gauss_q(h: torch.Tensor) -> torch.Tensor:
    """Irwin-Hall(4) over the four 16-bit fields of h: mean 0, sigma = 65536/sqrt(3) = 37837.2."""
    return (h & 0xFFFF) + (_lsr(h, 16) & 0xFFFF) + (_lsr(h, 32) & 0xFFFF) + _lsr(h, 48) - 131070


def read_lengths(n_reads: int, seed: int, first_read: int = 0, median: int = 30_000,
                 sigma: float = 0.9, lo: int = 2_000, hi: int = 3_000_000) -> np.ndarray:
    """Read lengths ~ LogNormal(ln median, sigma) clipped to [lo, hi]; counter-based per read index
    (numpy float64 on the host only: lengths are metadata, generated once and shared with every rank)."""
    idx = np.arange(first_read, first_read + n_reads, dtype=np.uint64)
    x = idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64((seed * 0xD1342543DE82EF95) & ((1 << 64) - 1))
    def mix(v):
        v = (v ^ (v >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        v = (v ^ (v >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return v ^ (v >> np.uint64(31))
    with np.errstate(over="ignore"):
        h1 = mix(x)
        h2 = mix(h1 ^ np.uint64(0xA5A5A5A5A5A5A5A5))
    u1 = ((h1 >> np.uint64(11)).astype(np.float64) + 0.5) / float(1 << 53)
    u2 = ((h2 >> np.uint64(11)).astype(np.float64) + 0.5) / float(1 << 53)
    g = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    ln = np.exp(np.log(float(median)) + sigma * g)
    return np.clip(np.rint(ln), lo, hi).astype(np.int64)


def generate_reads(lengths: np.ndarray, seed: int, model: PoreModel = R10_4_1, first_read: int = 0,
                   device: Optional[torch.device] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Concatenated int16 samples of reads first_read .. first_read+len(lengths)-1."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    lens = torch.as_tensor(np.asarray(lengths, dtype=np.int64), device=device)
    total = int(np.asarray(lengths, dtype=np.int64).sum())
    if out is None:
        out = torch.empty(total, dtype=torch.int16, device=device)
    if total == 0:
        return out
    starts = torch.cumsum(lens, 0) - lens
    rid = torch.repeat_interleave(torch.arange(first_read, first_read + lens.numel(), device=device), lens,
                                  output_size=total)
    pos = torch.arange(total, device=device) - torch.repeat_interleave(starts, lens, output_size=total)
    # events: Bernoulli(bps/f_s) per sample, forced at the first sample of each read
    h_ev = _hash(seed, 1, rid, pos)
    ev = (_lsr(h_ev, 32) < model.event_p_q32) | (pos == 0)
    k = torch.cumsum(ev.to(torch.int64), 0)
    k = k - torch.repeat_interleave(k[starts] - 1, lens, output_size=total)   # event index inside read
    del h_ev, ev
    level = model.mu * 1024 + ((_gauss_q(_hash(seed, 2, rid, k)) * model.sigma_level_q) // 37837)
    del k
    h_n = _hash(seed, 3, rid, pos)
    val = level + ((_gauss_q(h_n) * model.sigma_noise_q) // 37837)
    del level
    val = (val + 512) >> 10
    h_s = _hash(seed, 4, rid, pos)
    spike = _lsr(h_s, 32) < model.spike_p_q32
    mag = 500 + (h_s & 0xFFFF) % 2501
    mag = torch.where((h_s & 0x10000) != 0, mag, -mag)
    val = torch.where(spike, val + mag, val)
    out.copy_(torch.clamp(val, -4096, 4095).to(torch.int16))
    return out


def chunk_table(lengths: np.ndarray, chunk_size: int = DEFAULT_CHUNK_SIZE) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(chunk_offset[int64, in samples], chunk_n[uint32], chunk_read[int64]) of the signal-table rows the
    reference writer would produce for reads of these lengths (file_writer.cpp:119-143)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    per_read = np.maximum((lengths + chunk_size - 1) // chunk_size, 0)
    read_of = np.repeat(np.arange(lengths.size, dtype=np.int64), per_read)
    first = np.cumsum(per_read) - per_read
    j = np.arange(read_of.size, dtype=np.int64) - first[read_of]
    read_start = np.cumsum(lengths) - lengths
    off = read_start[read_of] + j * chunk_size
    n = np.minimum(lengths[read_of] - j * chunk_size, chunk_size).astype(np.uint32)
    return off, n, read_of


def write_chunk_file(path: str, samples: np.ndarray, off: np.ndarray, n: np.ndarray) -> None:
    """`[u32 n][int16 x n]...` chunk file of the reference's benchmark (src/c++/benchmark.cpp:26-49)."""
    with open(path, "wb") as f:
        for o, c in zip(off, n):
            f.write(np.uint32(c).tobytes())
            f.write(np.ascontiguousarray(samples[o : o + c], dtype=np.int16).tobytes())


def read_chunk_file(path: str):
    data = np.fromfile(path, dtype=np.uint8)
    pos, chunks = 0, []
    while pos < data.size:
        c = int(data[pos : pos + 4].view(np.uint32)[0]); pos += 4
        chunks.append(data[pos : pos + 2 * c].view(np.int16).copy()); pos += 2 * c
    return chunks
