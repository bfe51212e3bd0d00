"""Small decode/encode case for `compute-sanitizer --tool memcheck` (all codecs, ragged sizes, corrupt chunks)."""
import sys
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import recipes
from oracle_binding import Oracle
from rawnanoporesignalcompression_b200.codec import Codec
from rawnanoporesignalcompression_b200 import synth

o = Oracle(); cd = Codec(0)
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
print("sanitize case ok", cd.decode_class_counts())
