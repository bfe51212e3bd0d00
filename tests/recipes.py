"""Deterministic input recipes mirroring the reference's own unit tests plus edge cases.
Pure numpy / python ints so they reproduce bit-for-bit on any box."""
from __future__ import annotations

import numpy as np

EDGE_LENGTHS = [0, 1, 3, 4, 5, 7, 8, 9, 63, 64, 65, 4095, 4096, 4097]
FULL_LENGTHS = [102_399, 102_400, 102_401]


def iota(n: int) -> np.ndarray:
    """signal_compression_tests.cpp:17-18 / signal_table_tests.cpp:33 (std::iota into int16, wraps)."""
    return (np.arange(n, dtype=np.int64) & 0xFFFF).astype(np.uint16).view(np.int16)


def ones(n: int) -> np.ndarray:
    """signal_table_tests.cpp:35 (10000 x 1)."""
    return np.ones(n, dtype=np.int16)


def minstd_uniform(n: int, seed: int = 1) -> np.ndarray:
    """svb16_scalar_tests.cpp:18-23: minstd_rand (x <- 48271 x mod 2^31-1) drawing uniform int16. libstdc++'s
    uniform_int_distribution mapping is implementation-defined; here the low 16 bits of each draw are used."""
    out = np.empty(n, dtype=np.uint16)
    x = seed
    for i in range(n):
        x = (x * 48271) % 2147483647
        out[i] = x & 0xFFFF
    return out.view(np.int16)


def mt19937_64_mod256(n: int, seed: int) -> np.ndarray:
    """compression_decompression.cpp:76-113: std::mt19937_64(seed)() % 256."""
    NN, MM = 312, 156
    MATRIX_A, UM, LM, MASK = 0xB5026F5AA96619E9, 0xFFFFFFFF80000000, 0x7FFFFFFF, (1 << 64) - 1
    mt = [0] * NN
    mt[0] = seed & MASK
    for i in range(1, NN):
        mt[i] = (6364136223846793005 * (mt[i - 1] ^ (mt[i - 1] >> 62)) + i) & MASK
    out = np.empty(n, dtype=np.int16)
    idx = NN
    for k in range(n):
        if idx >= NN:
            for i in range(NN):
                x = (mt[i] & UM) | (mt[(i + 1) % NN] & LM)
                mt[i] = mt[(i + MM) % NN] ^ (x >> 1) ^ (MATRIX_A if x & 1 else 0)
            idx = 0
        x = mt[idx]; idx += 1
        x ^= (x >> 29) & 0x5555555555555555
        x ^= (x << 17) & 0x71D67FFFEDA60000
        x ^= (x << 37) & 0xFFF7EEE000000000
        x ^= x >> 43
        out[k] = x % 256
    return out


def uniform16(n: int, seed: int) -> np.ndarray:
    """Incompressible chunk: uniform int16 (the reference errors on C4/C5 for a full chunk, Appendix A)."""
    return np.random.default_rng(seed).integers(-32768, 32768, n).astype(np.int16)


def extremes(n: int, seed: int) -> np.ndarray:
    """+-32767/-32768 steps: exercises the mod-2^16 wrap of delta and zigzag."""
    return np.random.default_rng(seed).choice(np.array([-32768, 32767, 0, 1, -1], dtype=np.int16), n)


def walk(n: int, seed: int, sigma: float = 30.0) -> np.ndarray:
    d = np.random.default_rng(seed).normal(0.0, sigma, n)
    return np.cumsum(np.rint(d)).astype(np.int64).astype(np.int16)


def named_recipes():
    """(name, samples) pairs used for the committed golden vectors."""
    r = [("iota10000", iota(10000)), ("iota100000", iota(100000)), ("ones10000", ones(10000)),
         ("minstd1024", minstd_uniform(1024)), ("minstd1000", minstd_uniform(1000)), ("minstd20000", minstd_uniform(20000)),
         ("mt64_1024_s1", mt19937_64_mod256(1024, 1)), ("mt64_4096_s2", mt19937_64_mod256(4096, 2)), ("mt64_16384_s3", mt19937_64_mod256(16384, 3)),
         ("uniform_full", uniform16(102400, 11)), ("extremes5000", extremes(5000, 12)), ("walk_full", walk(102400, 13, 8.0)),
         ("walk_noisy_full", walk(102400, 14, 300.0)), ("walk_over", walk(102401, 15, 30.0))]
    for n in EDGE_LENGTHS:
        r.append((f"walk{n}", walk(n, 100 + n)))
    return r
