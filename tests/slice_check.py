"""Helper run in a subprocess (RNSC_SLICE_SAMPLES is read once per process): a host-buffer batch that spans many
pipeline slices must give the same bytes as the oracle, chunk by chunk, and round-trip. argv: <lib path or 'gpu'> <codecs>"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recipes
from oracle_binding import Oracle
from rawnanoporesignalcompression_b200.codec import Codec

lib = sys.argv[1]
codecs = [int(c) for c in sys.argv[2].split(",")]
cd = Codec(0) if lib == "gpu" else Codec(0, lib_path=lib)
orc = Oracle()
lens = [300, 0, 4096, 17, 9000, 5000, 1, 2500, 7000, 64, 12000, 3]
chunks = [recipes.walk(n, 40 + k, 11.0) if k % 3 else recipes.extremes(n, 50 + k) for k, n in enumerate(lens)]
for codec in codecs:
    comp, st = cd.compress_batch(chunks, codec)
    for got, s, c in zip(comp, st, chunks):
        ws, wb = orc.encode_status(codec, c)
        assert (s != 0) == (ws != 0), (codec, len(c), s, ws)
        if not ws:
            assert got == wb, (codec, len(c))
    ok = [i for i in range(len(chunks)) if not st[i]]
    dec, st2 = cd.decompress_batch([comp[i] for i in ok], [len(chunks[i]) for i in ok], codec)
    assert not st2.any() and all((a == chunks[i]).all() for a, i in zip(dec, ok))
# the same batch with the GPU entropy stage: other bytes, but every chunk must decode to its samples through the reference-side
# decoder (the oracle) and through this library; chunk offsets must be consistent across the slices
if len(sys.argv) > 3 and sys.argv[3] == "gpu_entropy":
    cd.set_entropy_stage("gpu")
    for codec in [c for c in codecs if c != 7]:
        comp, st = cd.compress_batch(chunks, codec)
        for got, s, c in zip(comp, st, chunks):
            ws, _ = orc.encode_status(codec, c)
            assert (s != 0) == (ws != 0), (codec, len(c), s, ws)
            if not s:
                assert (orc.decode(codec, got, len(c)) == c).all(), (codec, len(c))
        ok = [i for i in range(len(chunks)) if not st[i]]
        dec, st2 = cd.decompress_batch([comp[i] for i in ok], [len(chunks[i]) for i in ok], codec)
        assert not st2.any() and all((a == chunks[i]).all() for a, i in zip(dec, ok))
    cd.set_entropy_stage("libzstd")
cd.close()
print("slices ok")
