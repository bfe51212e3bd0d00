"""Read-sharded split of a signal table across the GPUs of one box.

Every chunk is an independent codec unit (predictor reset per chunk, file_writer.cpp:131-141), so the table shards
naturally: contiguous ranges of READS (a read's chunks stay together, file order is preserved), boundaries on the
prefix sum of per-read sample counts at the k/G quantiles. Each rank runs its own batcher and kernels; the only thing
that crosses ranks is (chunk count, byte count) per rank, gathered once to turn rank-local compressed offsets into
global ones -- no collective on the data path.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np


def partition_reads(read_samples: Sequence[int], world: int) -> np.ndarray:
    """Boundaries b[0..world] (read indices) so that rank r owns reads [b[r], b[r+1]) and every rank gets about the
    same number of samples. Deterministic, so every rank computes the same split without communicating."""
    n = np.asarray(read_samples, dtype=np.int64)
    if world < 1:
        raise ValueError("world must be >= 1")
    csum = np.concatenate([[0], np.cumsum(n)])
    total = int(csum[-1])
    b = np.zeros(world + 1, dtype=np.int64)
    for r in range(1, world):
        # first read whose start is at or past the r/world quantile of the samples
        b[r] = int(np.searchsorted(csum[:-1], (total * r) // world, side="left"))
    b[world] = n.size
    return np.maximum.accumulate(b)


def shard_chunks(read_of_chunk: np.ndarray, bounds: np.ndarray, rank: int) -> Tuple[int, int]:
    """Chunk index range [c0, c1) of the reads of `rank` (chunks are stored in read order)."""
    c0 = int(np.searchsorted(read_of_chunk, bounds[rank], side="left"))
    c1 = int(np.searchsorted(read_of_chunk, bounds[rank + 1], side="left"))
    return c0, c1


def gather_layout(local_chunks: int, local_bytes: int, group=None) -> Tuple[np.ndarray, np.ndarray]:
    """All-gathers (chunks, compressed bytes) of every rank -- 2 integers per rank -- and returns the exclusive
    prefix sums (chunk base, byte base) per rank with the totals appended (arrays of world+1)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return np.array([0, local_chunks], np.int64), np.array([0, local_bytes], np.int64)
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor([local_chunks, local_bytes], dtype=torch.int64, device=dev)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    a = torch.stack(allv).cpu().numpy()
    cb = np.concatenate([[0], np.cumsum(a[:, 0])]).astype(np.int64)
    bb = np.concatenate([[0], np.cumsum(a[:, 1])]).astype(np.int64)
    return cb, bb


def global_offsets(local_comp_off: np.ndarray, byte_base: int) -> np.ndarray:
    """Rank-local compressed offsets (nchunks+1, starting at 0) -> offsets in the concatenated global stream."""
    return np.asarray(local_comp_off, dtype=np.int64) + int(byte_base)


class ShardedCodec:
    """One rank's view of a read-sharded signal table: which reads/chunks it owns and the codec that processes them."""

    def __init__(self, codec, read_samples: Sequence[int], rank: int, world: int, chunk_size: Optional[int] = None):
        from . import synth
        self.codec, self.rank, self.world = codec, rank, world
        self.read_samples = np.asarray(read_samples, dtype=np.int64)
        self.bounds = partition_reads(self.read_samples, world)
        self.chunk_off, self.chunk_n, self.read_of_chunk = synth.chunk_table(self.read_samples, chunk_size or synth.DEFAULT_CHUNK_SIZE)
        self.c0, self.c1 = shard_chunks(self.read_of_chunk, self.bounds, rank)
        self.sample0 = int(self.chunk_off[self.c0]) if self.c1 > self.c0 else 0

    @property
    def local_reads(self) -> Tuple[int, int]:
        return int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])

    def local_table(self) -> Tuple[np.ndarray, np.ndarray]:
        """(sample offsets relative to this rank's first sample, sample counts) of this rank's chunks."""
        return (self.chunk_off[self.c0 : self.c1] - self.sample0).astype(np.uint64), self.chunk_n[self.c0 : self.c1].astype(np.uint32)

    def compress_local(self, local_samples: np.ndarray, codec_id):
        """Encodes this rank's chunks; returns (bytes, local offsets, status, (chunk base, byte base, totals))."""
        soff, n = self.local_table()
        out, coff, status = self.codec.compress_flat(local_samples, soff, n, codec_id)
        cb, bb = gather_layout(int(n.size), int(coff[-1]))
        return out, coff, status, (cb, bb)

    def decompress_local(self, comp: np.ndarray, comp_off: np.ndarray, codec_id):
        _, n = self.local_table()
        return self.codec.decompress_flat(comp, comp_off, n, codec_id)
