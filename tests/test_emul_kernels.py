"""Runs the *kernel sources* (csrc/*.cuh compiled with g++ on the CUDA-on-CPU fiber emulator of tests/emul)
through the C ABI against the oracle on small inputs. This exercises host logic (staging, offset tables,
container assembly, status codes) and the kernels' indexing without a GPU; it is not the product path --
the -m gpu tests run the nvcc-built library on the B200."""
import numpy as np
import pytest

import recipes
from rawnanoporesignalcompression_b200.codec import Codec


@pytest.fixture(scope="module")
def ecodec(emul_lib):
    cd = Codec(0, lib_path=emul_lib)
    yield cd
    cd.close()


@pytest.mark.parametrize("codec", range(8))
def test_small_roundtrip_all_codecs(ecodec, oracle, codec):
    for n in (0, 1, 5, 9, 65, 700, 4097):
        for x in (recipes.walk(n, n + 3), recipes.extremes(n, n + 4)):
            st, want = oracle.encode_status(codec, x)
            if st:
                continue
            assert ecodec.compress_signal(x, codec) == want
            assert (ecodec.decompress_signal(want, n, codec) == x).all()


def test_batch_offsets_and_status(ecodec, oracle):
    chunks = [recipes.walk(n, n) for n in (300, 0, 4096, 17, 9000)]
    for codec in (0, 6, 7):
        comp, st = ecodec.compress_batch(chunks, codec)
        assert not st.any() and comp == [oracle.encode(codec, c) for c in chunks]
        dec, st = ecodec.decompress_batch(comp, [len(c) for c in chunks], codec)
        assert not st.any() and all((a == b).all() for a, b in zip(dec, chunks))
        # one corrupted chunk does not disturb its neighbours
        bad = list(comp)
        bad[2] = bad[2][:20] + bytes([bad[2][20] ^ 0xFF]) + bad[2][21:]
        bad_n = [len(c) for c in chunks]
        bad_n[4] -= 8
        dec, st = ecodec.decompress_batch(bad, bad_n, codec)
        assert st[0] == 0 and st[1] == 0 and st[3] == 0 and (dec[0] == chunks[0]).all() and (dec[3] == chunks[3]).all()
        assert st[4] == 4       # "Remaining data at end of signal buffer" (RC: header sample count mismatch)
        if codec != 7:
            assert st[2] != 0   # corrupted zstd frame (the range coder has no integrity check, like the reference)


def test_huffman_and_sequences_path(ecodec, oracle):
    # 20000 samples: VBZ frame = Huffman 4-stream literals + FSE sequences; C5 = Huffman keys/M, raw S
    x = recipes.walk(20000, 77, 12.0)
    for codec in (0, 2, 6):
        b = oracle.encode(codec, x)
        assert (ecodec.decompress_signal(b, x.size, codec) == x).all()


def test_fast_path_routing_and_sequences(ecodec, oracle):
    """C5/C4/VBZ0 chunks (single-block frames, predefined/RLE sequence tables) must take the planner + fast kernel,
    including a full 102,400-sample chunk (one CTA per SM class) and key frames that carry LZ sequences."""
    from rawnanoporesignalcompression_b200 import synth
    lengths = np.array([3000, 30000, 102400 + 5000], dtype=np.int64)
    sig = synth.generate_reads(lengths, 99, synth.R10_4_1).numpy()
    off, n, _ = synth.chunk_table(lengths)
    chunks = [sig[o : o + c] for o, c in zip(off, n)]
    # long constant runs make zstd emit sequences in the key stream (code 0 repeated)
    flat = chunks[1].copy(); flat[5000:9000] = flat[4999]; flat[20000:26000] = 7
    chunks.append(flat)
    for codec in (6, 5, 1):
        comp = [oracle.encode(codec, c) for c in chunks]
        dec, st = ecodec.decompress_batch(comp, [len(c) for c in chunks], codec)
        assert not st.any() and all((a == b).all() for a, b in zip(dec, chunks))
        fs, fa, fb, slow = ecodec.decode_class_counts()
        assert fs + fa + fb + slow == len(chunks)
        if codec != 1:
            assert slow == 0 and fb >= 1 and fs >= 1, (codec, fs, fa, fb, slow)


def test_fuzz_corrupt_chunks_never_crash_and_agree_with_rfc_decoder(ecodec, oracle):
    """Random byte flips anywhere in a chunk (headers, Huffman tree, bitstreams, sequences): the decoder must return a
    status or the same samples as the oracle's RFC 8878 decoder -- never crash, hang or write out of bounds."""
    x = recipes.walk(6000, 31, 14.0)
    flat = x.copy(); flat[1000:1600] = flat[999]
    for seed, (codec, src) in [(sd, cs) for sd in (1234, 1, 2, 3) for cs in ((6, x), (0, x), (6, flat), (0, flat), (1, x))]:
        rng = np.random.default_rng(seed)
        good = oracle.encode(codec, src)
        trials = []
        for _ in range(24):
            b = bytearray(good)
            for _k in range(int(rng.integers(1, 4))):
                pos = int(rng.integers(0, len(b)))
                b[pos] ^= int(rng.integers(1, 256))
            trials.append(bytes(b))
        trials.append(good[: len(good) // 2])                       # truncated
        trials.append(good + b"\x00\x01\x02")                       # trailing garbage
        outs, st = ecodec.decompress_batch(trials + [good], [src.size] * (len(trials) + 1), codec)
        assert st[-1] == 0 and (outs[-1] == src).all()              # the intact chunk next to them is untouched
        for t, o, s in zip(trials, outs, st):
            ost, oy = oracle.decode_status(codec, t, src.size, own_zstd=True)
            # DESIGN.md section 2, second deliberate difference: corrupted keys can make ONE stream over-read while the last
            # one still ends where it should; librnsc bounds every stream (status 4, "Remaining data"), the reference checks
            # only the final pointer and returns OK with garbage
            if s == 4 and ost == 0:
                continue
            assert (s != 0) == (ost != 0), (seed, codec, s, ost)
            if ost == 0:
                assert (o == oy).all()


@pytest.mark.parametrize("codec", range(7))
def test_splitter_encode_ragged_alignments(ecodec, oracle, codec):
    """Second-generation splitter encode (svb_pack2.cuh): chunks back to back at every 16-byte phase, lengths around the
    16-sample thread granule and the 4096-sample tile, class mixes from all-zero deltas to bursts of large values."""
    lens = [1, 7, 8, 15, 16, 17, 33, 4095, 4096, 4097, 3, 12289, 2, 5000, 9, 8193]
    chunks = []
    for k, n in enumerate(lens):
        kind = k % 4
        if kind == 0: x = recipes.walk(n, 100 + k, 12.0)
        elif kind == 1: x = recipes.extremes(n, 200 + k)
        elif kind == 2: x = np.full(n, -7, np.int16)
        else: x = recipes.walk(n, 300 + k, 3.0)
        chunks.append(x)
    comp, st = ecodec.compress_batch(chunks, codec)
    want = [oracle.encode_status(codec, c) for c in chunks]
    for got, s, (ws, wb), c in zip(comp, st, want, chunks):
        if ws:
            assert s != 0
        else:
            assert s == 0 and got == wb, (codec, len(c))


def test_encode_pipeline_many_slices(emul_lib):
    """rnsc_compress_batch stages slice k+1 while host threads compress slice k: force ~10 slices on a small batch."""
    import os, subprocess, sys
    env = dict(os.environ, RNSC_SLICE_SAMPLES="6000")
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "slice_check.py"), emul_lib, "6,0,2,7", "gpu_entropy"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "slices ok" in r.stdout, r.stdout + r.stderr


def test_foreign_zstd_frames(ecodec, oracle):
    """Frames from other zstd levels / with checksums / with two blocks (tests/zstd_foreign.py) through the kernel sources."""
    import zstd_foreign as zf
    for name, x in zf.signals():
        for level, ck in zf.SETTINGS[::2]:
            assert (ecodec.decompress_signal(zf.vbz_frame(oracle, x, level, ck), x.size, 0) == x).all(), (name, level, ck)
        assert (ecodec.decompress_signal(zf.c5_container(oracle, x, 19, True), x.size, 6) == x).all(), name


def test_gpu_entropy_stage_frames_decode_everywhere(ecodec, oracle):
    """RNSC_ENTROPY_GPU (csrc/zstd_encode.cuh): the frames written by the device encoder are valid Zstandard -- libzstd, the
    oracle's RFC 8878 decoder, the reference's own decompress_signal_* (when built) and this library's decoder return the
    input samples; chunks that exceed the reference's output capacity report status 3 like the host path."""
    from oracle_binding import RefOracle
    from rawnanoporesignalcompression_b200 import synth
    ref = RefOracle() if RefOracle.available() else None
    lengths = np.array([3000, 30000, 102400 + 5000, 700, 64, 10, 0, 1], dtype=np.int64)
    sig = synth.generate_reads(lengths, 99, synth.R10_4_1).numpy()
    off, n, _ = synth.chunk_table(lengths)
    chunks = [sig[o : o + c] for o, c in zip(off, n)]
    chunks += [recipes.walk(20000, 77, 12.0), recipes.extremes(5000, 3), np.zeros(9000, np.int16), np.arange(10000, dtype=np.int16), recipes.uniform16(4000, 5)]
    ecodec.set_entropy_stage("gpu")
    try:
        for codec in range(7):
            comp, st = ecodec.compress_batch(chunks, codec)
            for c, b, s in zip(chunks, comp, st):
                host_st, _ = oracle.encode_status(codec, c)
                if host_st == 3 and codec in (5, 6):      # C4/C5 "Not enough space in destination buffer" (C5.hpp:420-427)
                    assert s == 3
                if s:
                    assert s == 3
                    continue
                for r, y in (oracle.decode_status(codec, b, len(c)), oracle.decode_status(codec, b, len(c), own_zstd=True)) + ((ref.decode_status(codec, b, len(c)),) if ref else ()):
                    assert r == 0 and (y == c).all(), (codec, len(c))
            ok = [i for i, s in enumerate(st) if s == 0]
            dec, st2 = ecodec.decompress_batch([comp[i] for i in ok], [len(chunks[i]) for i in ok], codec)
            assert not st2.any() and all((a == chunks[i]).all() for a, i in zip(dec, ok))
    finally:
        ecodec.set_entropy_stage("libzstd")
