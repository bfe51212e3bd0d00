"""Parity tests proper: the nvcc-built librnsc.so on a real GPU, through the C ABI, against the oracle and the
committed golden vectors (bit-exact: this is integer/byte work)."""
import hashlib

import numpy as np
import pytest

import recipes

pytestmark = pytest.mark.gpu
CODECS = list(range(8))
NAMES = ["VBZ", "VBZ0", "C1", "C2", "C3", "C4", "C5", "RC"]


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def test_golden_real_chunks(gpu_codec, golden, real_chunks, oracle):
    """Config 1: the reference's own pod5 fixture (22 stored VBZ chunks, 1,548,931 samples)."""
    frames = [f for f, _ in real_chunks]
    ns = [n for _, n in real_chunks]
    dec, st = gpu_codec.decompress_batch(frames, ns, 0)
    assert not st.any()
    for x, g in zip(dec, golden["real"]):
        assert sha(x.tobytes()) == g["samples_sha256"]
    for codec in CODECS:
        comp, st = gpu_codec.compress_batch(dec, codec)
        assert not st.any()
        for b, g in zip(comp, golden["real"]):
            ge = g["codecs"][str(codec)]
            assert len(b) == ge["size"] and sha(b) == ge["sha256"], NAMES[codec]
        back, st = gpu_codec.decompress_batch(comp, ns, codec)
        assert not st.any() and all((a == b).all() for a, b in zip(back, dec)), NAMES[codec]


@pytest.mark.parametrize("codec", CODECS, ids=NAMES)
def test_golden_recipes(gpu_codec, golden, codec):
    items = recipes.named_recipes()
    xs = [x for _, x in items]
    comp, st = gpu_codec.compress_batch(xs, codec)
    for (name, x), b, s in zip(items, comp, st):
        ge = golden["recipes"][name]["codecs"][str(codec)]
        assert int(s) == ge["status"], (name, int(s))
        if ge["status"] == 0:
            assert len(b) == ge["size"] and sha(b) == ge["sha256"], name
    ok = [i for i, s in enumerate(st) if s == 0]
    back, st2 = gpu_codec.decompress_batch([comp[i] for i in ok], [xs[i].size for i in ok], codec)
    assert not st2.any()
    for i, y in zip(ok, back):
        assert (y == xs[i]).all(), items[i][0]


@pytest.mark.parametrize("codec", CODECS, ids=NAMES)
def test_edge_lengths_vs_oracle(gpu_codec, enc_oracle, codec):
    for n in recipes.EDGE_LENGTHS + recipes.FULL_LENGTHS:
        for x in (recipes.walk(n, 1000 + n, 25.0), recipes.extremes(n, 2000 + n)):
            st, want = enc_oracle.encode_status(codec, x)
            if st:
                with pytest.raises(Exception):
                    gpu_codec.compress_signal(x, codec)
                continue
            assert gpu_codec.compress_signal(x, codec) == want, (n, NAMES[codec])
            assert (gpu_codec.decompress_signal(want, n, codec) == x).all(), (n, NAMES[codec])


def test_error_semantics(gpu_codec, oracle):
    from rawnanoporesignalcompression_b200._lib import RnscError
    x = recipes.walk(30000, 5)
    for codec in range(7):
        b = oracle.encode(codec, x)
        pos = 0 if codec < 2 else 8
        bad = b[:pos] + b"\x00\x00\x00\x00" + b[pos + 4:]
        with pytest.raises(RnscError) as e:
            gpu_codec.decompress_signal(bad, x.size, codec)
        assert e.value.status == 1 == oracle.decode_status(codec, bad, x.size)[0]
        with pytest.raises(RnscError) as e:
            gpu_codec.decompress_signal(b, x.size - 8, codec)
        assert e.value.status == 4 == oracle.decode_status(codec, b, x.size - 8)[0]
        # flip a byte in the middle of the entropy-coded payload: must not crash; status or output mismatch
        mid = len(b) // 2
        bad = b[:mid] + bytes([b[mid] ^ 0x5A]) + b[mid + 1:]
        outs, st = gpu_codec.decompress_batch([bad, b], [x.size, x.size], codec)
        assert st[1] == 0 and (outs[1] == x).all()
        # Arbiter for a corrupted entropy payload is RFC 8878 (the oracle's own frame decoder): "a bitstream
        # that is not entirely and exactly consumed is faulty". libzstd 1.5.5's 4-stream Huffman fast path
        # does not check that, so it accepts some corrupt frames the specification (and librnsc) rejects;
        # whatever libzstd rejects must be rejected here too, and what both accept must decode identically.
        ost, oy = oracle.decode_status(codec, bad, x.size, own_zstd=True)
        lst, ly = oracle.decode_status(codec, bad, x.size)
        assert (st[0] != 0) == (ost != 0), NAMES[codec]
        if lst != 0:
            assert st[0] != 0, NAMES[codec]
        if ost == 0:
            assert (outs[0] == oy).all()
            if lst == 0:
                assert (outs[0] == ly).all()
    u = recipes.uniform16(102400, 11)
    for codec in (5, 6):
        with pytest.raises(RnscError) as e:
            gpu_codec.compress_signal(u, codec)
        assert e.value.status == 3


def test_synthetic_batch_all_codecs(gpu_codec, enc_oracle):
    """Seeded synthetic nanopore-shaped reads (R10.4.1 and R9.4.1 models), ragged chunking; every chunk's bytes against
    the reference encoder."""
    oracle = enc_oracle
    from rawnanoporesignalcompression_b200 import synth
    for model, seed in ((synth.R10_4_1, 20261018), (synth.R9_4_1, 20261019)):
        lengths = synth.read_lengths(12, seed)
        sig = synth.generate_reads(lengths, seed, model).numpy()
        off, n, _ = synth.chunk_table(lengths)
        chunks = [sig[o : o + c] for o, c in zip(off, n)]
        for codec in CODECS:
            comp, st = gpu_codec.compress_batch(chunks, codec)
            assert not st.any()
            for c, b in zip(chunks, comp):
                assert b == oracle.encode(codec, c), NAMES[codec]
            back, st = gpu_codec.decompress_batch(comp, [len(c) for c in chunks], codec)
            assert not st.any() and all((a == b).all() for a, b in zip(back, chunks)), NAMES[codec]


def test_device_resident_api(gpu_codec, oracle):
    import torch
    from rawnanoporesignalcompression_b200 import synth
    dev = torch.device("cuda:0")
    lengths = synth.read_lengths(40, 7)
    sig = synth.generate_reads(lengths, 7, synth.R10_4_1, device=dev)
    assert (sig.cpu() == synth.generate_reads(lengths, 7, synth.R10_4_1)).all(), "generator must be device independent"
    off, n, _ = synth.chunk_table(lengths)
    nch = len(n)
    d_off = torch.as_tensor(off, device=dev)
    d_n = torch.as_tensor(n.astype(np.int32), device=dev)
    host = sig.cpu().numpy()
    for codec in range(7):
        cap = int(sum(gpu_codec.lib.rnsc_max_stream_bytes(codec, int(k)) for k in n))
        d_streams = torch.empty(cap, dtype=torch.uint8, device=dev)
        d_soff = torch.empty(5 * nch + 1, dtype=torch.int64, device=dev)
        gpu_codec.dev_svb_encode(codec, sig, d_off, d_n, d_streams, d_soff)
        torch.cuda.synchronize()
        soff = d_soff.cpu().numpy()
        st_host = d_streams.cpu().numpy()
        for i in range(nch):
            want = oracle.streams(codec, host[off[i] : off[i] + n[i]])
            got = [st_host[soff[5 * i + k] : soff[5 * i + k + 1]].tobytes() for k in range(5)]
            assert got == want, (NAMES[codec], i)
        out = torch.zeros_like(sig)
        status = torch.full((nch,), -1, dtype=torch.int32, device=dev)
        gpu_codec.dev_svb_decode(codec, d_streams, d_soff, d_n, d_off, out, status)
        torch.cuda.synchronize()
        assert not status.any() and (out == sig).all(), NAMES[codec]
        # the one-pass variant (per-chunk slabs): same stream bytes, start + size per stream
        total = int(n.sum())
        assert sum(gpu_codec.lib.rnsc_svb_slab_bytes(codec, int(k)) for k in n) <= 4 * total + 128 * nch + 64
        d_slabs = torch.zeros(4 * total + 128 * nch + 64, dtype=torch.uint8, device=dev)
        d_start = torch.empty(5 * nch, dtype=torch.int64, device=dev); d_size = torch.empty(5 * nch, dtype=torch.int32, device=dev)
        gpu_codec.dev_svb_encode_slabs(codec, sig, d_off, d_n, total, d_slabs, d_start, d_size)
        torch.cuda.synchronize()
        start, size, sl_host = d_start.cpu().numpy(), d_size.cpu().numpy(), d_slabs.cpu().numpy()
        for i in range(nch):
            want = oracle.streams(codec, host[off[i] : off[i] + n[i]])
            got = [sl_host[start[5 * i + k] : start[5 * i + k] + size[5 * i + k]].tobytes() for k in range(5)]
            assert got == want, (NAMES[codec], i, "slabs")
        with pytest.raises(Exception):
            gpu_codec.dev_svb_encode_slabs(codec, sig, d_off, d_n, total, d_slabs[: total], d_start, d_size)
    # whole-chunk device decode of oracle-produced bytes at unaligned offsets
    for codec in (0, 6):
        comp = [oracle.encode(codec, host[o : o + c]) for o, c in zip(off, n)]
        coff = np.zeros(nch + 1, dtype=np.int64)
        np.cumsum([len(b) for b in comp], out=coff[1:])
        d_comp = torch.as_tensor(np.frombuffer(b"".join(comp) + b"\0" * 64, dtype=np.uint8).copy(), device=dev)
        d_coff = torch.as_tensor(coff, device=dev)
        out = torch.zeros_like(sig)
        status = torch.full((nch,), -1, dtype=torch.int32, device=dev)
        gpu_codec.dev_decode(codec, d_comp, d_coff, d_n, d_off, int(n.max()), out, status)
        torch.cuda.synchronize()
        assert not status.any() and (out == sig).all(), NAMES[codec]
    # device prefix scan
    v = torch.randint(0, 100000, (123457,), dtype=torch.int32, device=dev)
    o = torch.empty(v.numel() + 1, dtype=torch.int64, device=dev)
    gpu_codec.dev_exclusive_scan_u32(v, o)
    torch.cuda.synchronize()
    ref = torch.zeros(v.numel() + 1, dtype=torch.int64, device=dev)
    ref[1:] = torch.cumsum(v.to(torch.int64), 0)
    assert (o == ref).all()


def test_pod5_named_entry_points(gpu_codec, oracle):
    import ctypes as C
    lib = gpu_codec.lib
    x = recipes.walk(50000, 9)
    for fn_c, fn_d, codec in ((lib.rnsc_pod5_vbz_compress_signal, lib.rnsc_pod5_vbz_decompress_signal, 0),
                              (lib.rnsc_pod5_pinanoraw_compress_signal, lib.rnsc_pod5_pinanoraw_decompress_signal, 6)):
        cap = x.size * 4   # benchmark.cpp:138-140 passes samples*4
        out = np.zeros(cap, np.uint8)
        size = C.c_size_t(cap)
        assert fn_c(C.c_void_p(x.ctypes.data), x.size, C.c_void_p(out.ctypes.data), C.byref(size)) == 0
        assert out[: size.value].tobytes() == oracle.encode(codec, x)
        y = np.zeros(x.size, np.int16)
        assert fn_d(C.c_void_p(out.ctypes.data), size.value, x.size, C.c_void_p(y.ctypes.data)) == 0
        assert (y == x).all()
        small = C.c_size_t(10)
        assert fn_c(C.c_void_p(x.ctypes.data), x.size, C.c_void_p(out.ctypes.data), C.byref(small)) == 3
    assert lib.rnsc_set_pgnano_variant(4) == 0 and lib.rnsc_get_pgnano_variant() == 4
    out = np.zeros(x.size * 4, np.uint8); size = C.c_size_t(out.size)
    assert lib.rnsc_pod5_pinanoraw_compress_signal(C.c_void_p(x.ctypes.data), x.size, C.c_void_p(out.ctypes.data), C.byref(size)) == 0
    assert out[: size.value].tobytes() == oracle.encode(4, x)
    lib.rnsc_set_pgnano_variant(6)


def test_full_size_round_trip_property(gpu_codec):
    """Size-independent property at scale: encode -> decode == identity on ~60 M samples, C5 and VBZ, with the
    zstd stage of the encoder on host threads (sha256 over everything instead of an element-wise oracle)."""
    from rawnanoporesignalcompression_b200 import synth
    lengths = synth.read_lengths(1500, 424242)
    sig = synth.generate_reads(lengths, 424242, synth.R10_4_1).numpy()
    off, n, _ = synth.chunk_table(lengths)
    for codec in (6, 0):
        comp, coff, st = gpu_codec.compress_flat(sig, off.astype(np.uint64), n, codec)
        assert not st.any()
        out, soff, st = gpu_codec.decompress_flat(np.concatenate([comp, np.zeros(64, np.uint8)]), coff, n, codec)
        assert not st.any()
        assert sha(out.tobytes()) == sha(sig.tobytes())
        bits = 8.0 * int(coff[-1]) / sig.size
        assert 5.5 < bits < 7.5, bits


@pytest.mark.gpu
def test_encode_pipeline_many_slices(gpu_codec):
    """Host-buffer encode across many pipeline slices (RNSC_SLICE_SAMPLES is read once per process -> subprocess)."""
    import os, subprocess, sys
    env = dict(os.environ, RNSC_SLICE_SAMPLES="6000")
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "slice_check.py"), "gpu", "0,1,2,3,4,5,6,7", "gpu_entropy"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "slices ok" in r.stdout, r.stdout + r.stderr


def test_foreign_zstd_frames(gpu_codec, oracle):
    """Chunks whose frames were NOT written by `ZSTD_compress(..., 1)`: other levels, checksummed frames, two-block
    frames. The reference reads them through ZSTD_decompress (signal_compression.cpp:117, C5.hpp:588-655); the device
    decoder must return the same samples, one by one and mixed into a batch with stock level-1 chunks."""
    import zstd_foreign as zf
    sigs = zf.signals()
    frames, ns, want = [], [], []
    for name, x in sigs:
        for level, ck in zf.SETTINGS:
            frames.append(zf.vbz_frame(oracle, x, level, ck)); ns.append(x.size); want.append(x)
        frames.append(oracle.encode(0, x)); ns.append(x.size); want.append(x)
    got, st = gpu_codec.decompress_batch(frames, ns, 0)
    assert not st.any(), st
    assert all((g == w).all() for g, w in zip(got, want))
    assert (gpu_codec.decompress_signal(frames[2], ns[2], 0) == want[2]).all()
    conts, ns, want = [], [], []
    for name, x in sigs:
        for level, ck in ((19, True), (3, False), (-5, True)):
            conts.append(zf.c5_container(oracle, x, level, ck)); ns.append(x.size); want.append(x)
    got, st = gpu_codec.decompress_batch(conts, ns, 6)
    assert not st.any(), st
    assert all((g == w).all() for g, w in zip(got, want))


@pytest.mark.parametrize("model_name", ["R10.4.1", "R9.4.1"])
def test_encode_byte_parity_1000_chunks(gpu_codec, enc_oracle, model_name):
    """BASELINE config 3 at test scale: more than 1000 chunks (full-size and ragged) of each signal model, every codec, the
    product encoder's bytes against the reference encoder's chunk by chunk; then the decoder returns the samples."""
    from rawnanoporesignalcompression_b200 import synth
    seed = 20261018 if model_name == "R10.4.1" else 20261019
    lengths = synth.read_lengths(1000, seed, median=24_000)
    sig = synth.generate_reads(lengths, seed, synth.MODELS[model_name]).numpy()
    off, n, _ = synth.chunk_table(lengths)
    assert n.size >= 1000 and (n == synth.DEFAULT_CHUNK_SIZE).any()
    for codec in range(7):
        comp, coff, st = gpu_codec.compress_flat(sig, off.astype(np.uint64), n, codec)
        assert not st.any(), NAMES[codec]
        for i, (o, c) in enumerate(zip(off, n)):
            want = enc_oracle.encode(codec, sig[o : o + c])
            got = comp[int(coff[i]) : int(coff[i + 1])].tobytes()
            assert got == want, (NAMES[codec], i, int(c))
        back, _, st = gpu_codec.decompress_flat(np.concatenate([comp, np.zeros(64, np.uint8)]), coff, n, codec)
        assert not st.any() and (back == sig).all(), NAMES[codec]
    # range coder: a tenth of the chunks (the reference's coder runs at ~30 MB/s)
    sel = np.arange(0, n.size, 10)
    chunks = [sig[off[i] : off[i] + n[i]] for i in sel]
    comp, st = gpu_codec.compress_batch(chunks, 7)
    assert not st.any()
    for c, b in zip(chunks, comp):
        assert b == enc_oracle.encode(7, c)


@pytest.mark.parametrize("codec", [6, 0, 1, 3], ids=["C5", "VBZ", "VBZ0", "C2"])
def test_gpu_entropy_stage_device_encode(gpu_codec, enc_oracle, oracle, codec):
    """rnsc_dev_encode (GPU entropy stage, csrc/zstd_encode.cuh): > 1000 chunks of synthetic signal encoded on the device;
    every chunk decodes back to its samples through the reference's decoder (oracle/_ref when present, else libzstd via the
    port) and, as one batch, through this library's fast decoder; sizes stay within 3 % of the host-libzstd encoder."""
    import torch
    from rawnanoporesignalcompression_b200 import synth
    dev = torch.device("cuda:0")
    lengths = synth.read_lengths(1000, 20261018, median=24_000)
    sig = synth.generate_reads(lengths, 20261018, synth.R10_4_1, device=dev)
    off, n, _ = synth.chunk_table(lengths)
    total = int(lengths.sum())
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev); d_n = torch.from_numpy(n.astype(np.int32)).to(dev)
    d_comp = torch.empty(2 * total + total // 4 + 1104 * n.size + 64, dtype=torch.uint8, device=dev)
    d_coff = torch.empty(n.size + 1, dtype=torch.int64, device=dev); d_st = torch.empty(n.size, dtype=torch.int32, device=dev)
    gpu_codec.dev_encode(codec, sig, d_off, d_n, total, d_comp, d_coff, d_st)
    torch.cuda.synchronize()
    assert not d_st.any().item()
    coff = d_coff.cpu().numpy(); comp = d_comp[: int(coff[-1]) + 64].cpu().numpy(); host = sig.cpu().numpy()
    for i in range(0, n.size, 3):                                  # every third chunk through the CPU reference decoder
        b = comp[int(coff[i]) : int(coff[i + 1])].tobytes()
        r, y = enc_oracle.decode_status(codec, b, int(n[i]))
        assert r == 0 and (y == host[off[i] : off[i] + n[i]]).all(), i
    d_out = torch.empty_like(sig); d_st2 = torch.empty_like(d_st)
    gpu_codec.dev_decode(codec, d_comp, d_coff, d_n, d_off, int(n.max()), d_out, d_st2)
    torch.cuda.synchronize()
    assert not d_st2.any().item() and torch.equal(d_out, sig)
    _, hcoff, hst = gpu_codec.compress_flat(host, off.astype(np.uint64), n, codec)          # host libzstd level 1 (the reference's bytes)
    assert not hst.any()
    ratio = float(coff[-1]) / float(hcoff[-1])
    assert ratio < (1.03 if codec != 0 and codec != 1 else 1.10), ratio
