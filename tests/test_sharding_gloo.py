"""N > 1 path on CPU: two gloo ranks read-shard one signal table, each runs the batcher + kernels (the kernel sources
on the CUDA-on-CPU emulator, tests/emul) over its own reads, and only (chunk count, byte count) per rank is
all-gathered. The concatenation of the ranks' outputs must equal the single-process oracle encoding, chunk for chunk."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_reads_properties():
    from rawnanoporesignalcompression_b200.sharding import partition_reads, shard_chunks
    from rawnanoporesignalcompression_b200 import synth
    lengths = synth.read_lengths(500, 5)
    for world in (1, 2, 3, 8):
        b = partition_reads(lengths, world)
        assert b[0] == 0 and b[-1] == lengths.size and (np.diff(b) >= 0).all()
        per = [int(lengths[b[r] : b[r + 1]].sum()) for r in range(world)]
        assert sum(per) == int(lengths.sum())
        assert max(per) - min(per) <= 2 * int(lengths.max())          # balanced to within a read or two
        _, _, read_of = synth.chunk_table(lengths)
        spans = [shard_chunks(read_of, b, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == read_of.size and all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
    assert list(partition_reads([], 4)) == [0, 0, 0, 0, 0]


def _worker(rank, world, port, so, q):
    try:
        sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch.distributed as dist
        from rawnanoporesignalcompression_b200 import synth
        from rawnanoporesignalcompression_b200.codec import Codec
        from rawnanoporesignalcompression_b200.sharding import ShardedCodec, global_offsets
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        lengths = np.array([5000, 2100, 9000, 3000, 12000, 2500, 4000], dtype=np.int64)
        cd = Codec(0, lib_path=so)
        sh = ShardedCodec(cd, lengths, rank, world, chunk_size=4096)
        r0, r1 = sh.local_reads
        # every rank regenerates exactly its own reads (counter-based generator)
        mine = synth.generate_reads(lengths[r0:r1], 77, synth.R10_4_1, first_read=r0).numpy()
        res = {}
        for codec in (6, 0, 7):
            out, coff, status, (cb, bb) = sh.compress_local(mine, codec)
            assert not status.any()
            dec, _, st2 = sh.decompress_local(np.concatenate([out, np.zeros(64, np.uint8)]), coff, codec)
            assert not st2.any() and (dec == mine).all()
            res[codec] = (out.tobytes(), global_offsets(coff, bb[rank]).tolist(), cb.tolist(), bb.tolist())
        q.put((rank, (r0, r1), res, None))
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, None, None, traceback.format_exc()))


def test_two_rank_read_sharded_encode_matches_oracle(emul_lib, oracle):
    import multiprocessing as mp
    from rawnanoporesignalcompression_b200 import synth
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, emul_lib, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in procs:
        rank, reads, res, err = q.get(timeout=600)
        assert err is None, err
        got[rank] = (reads, res)
    for p in procs:
        p.join(timeout=60)
    lengths = np.array([5000, 2100, 9000, 3000, 12000, 2500, 4000], dtype=np.int64)
    sig = synth.generate_reads(lengths, 77, synth.R10_4_1).numpy()
    off, n, _ = synth.chunk_table(lengths, 4096)
    assert got[0][0][0] == 0 and got[0][0][1] == got[1][0][0] and got[1][0][1] == lengths.size
    for codec in (6, 0, 7):
        want = [oracle.encode(codec, sig[o : o + c]) for o, c in zip(off, n)]
        blob = got[0][1][codec][0] + got[1][1][codec][0]
        offs = got[0][1][codec][1][:-1] + got[1][1][codec][1]
        assert got[0][1][codec][2] == got[1][1][codec][2] and got[0][1][codec][2][-1] == len(want)      # same gathered layout on both ranks
        assert got[0][1][codec][3][-1] == len(blob)
        assert len(offs) == len(want) + 1
        for i, w in enumerate(want):
            assert blob[offs[i] : offs[i + 1]] == w, (codec, i)
