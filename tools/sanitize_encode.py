"""Encode-only case for `compute-sanitizer --tool memcheck|racecheck`: the splitter kernels (svb_pack2.cuh) on ragged chunks
at every 16-byte phase, all seven splitter codecs, checked against the oracle."""
import sys
import numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import recipes
from oracle_binding import Oracle
from rawnanoporesignalcompression_b200.codec import Codec

o = Oracle(); cd = Codec(0)
lens = [1, 7, 8, 15, 16, 17, 33, 2047, 2048, 2049, 3, 6200, 2, 5000, 9, 4097]
chunks = [recipes.walk(n, 100 + k, 12.0) if k % 3 else recipes.extremes(n, 200 + k) for k, n in enumerate(lens)]
for codec in range(7):
    got, st = cd.compress_batch(chunks, codec)
    for g, s, c in zip(got, st, chunks):
        ws, wb = o.encode_status(codec, c)
        assert (s != 0) == (ws != 0)
        if not ws:
            assert g == wb, (codec, len(c))
print("sanitize encode ok")
