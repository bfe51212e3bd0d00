"""Small decode/encode case for `compute-sanitizer --tool memcheck` (all codecs, ragged sizes, corrupt chunks)."""
import sys
import numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import recipes
from oracle_binding import Oracle
from rawnanoporesignalcompression_b200.codec import Codec
from rawnanoporesignalcompression_b200 import synth

o = Oracle(); cd = Codec(0, lib_path=sys.argv[1]) if len(sys.argv) > 1 else Codec(0)      # argv[1]: another build of the library (the ASAN build of the CPU emulator)
lengths = np.array([2500, 30000, 102400 + 700], dtype=np.int64)
sig = synth.generate_reads(lengths, 5, synth.R10_4_1).numpy()
off, n, _ = synth.chunk_table(lengths)
chunks = [sig[a:a + b] for a, b in zip(off, n)] + [recipes.walk(0, 1), recipes.walk(7, 2), recipes.extremes(3000, 3), recipes.iota(10000)]
rng = np.random.default_rng(7)
for codec in range(8):
    comp = [o.encode(codec, c) for c in chunks]
    got, st = cd.compress_batch(chunks, codec)
    assert not st.any() and got == comp, codec
    dec, st = cd.decompress_batch(comp, [len(c) for c in chunks], codec)
    assert not st.any() and all((a == b).all() for a, b in zip(dec, chunks)), codec
    if codec < 7:
        bad = []
        for c in comp[:3]:
            for _ in range(6):
                b = bytearray(c); pos = int(rng.integers(0, len(b))); b[pos] ^= int(rng.integers(1, 256)); bad.append(bytes(b))
        ns = [len(chunks[i]) for i in range(3) for _ in range(6)]
        cd.decompress_batch(bad, ns, codec)
# the GPU entropy stage (one-pass slab splitter, histogram, code build, emit, gather) through the host-buffer pipeline
cd.set_entropy_stage("gpu")
for codec in range(7):
    got, st = cd.compress_batch(chunks, codec)
    for g, s, c in zip(got, st, chunks):
        ws, _ = o.encode_status(codec, c)
        assert (s != 0) == (ws != 0), (codec, len(c), s, ws)
        if not s:
            assert (o.decode(codec, g, len(c)) == c).all(), (codec, len(c))
    ok = [i for i in range(len(chunks)) if not st[i]]
    dec, st2 = cd.decompress_batch([got[i] for i in ok], [len(chunks[i]) for i in ok], codec)
    assert not st2.any() and all((a == chunks[i]).all() for a, i in zip(dec, ok)), codec
cd.set_entropy_stage("libzstd")
print("sanitize case ok", cd.decode_class_counts())
