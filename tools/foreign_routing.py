"""Which decoder class the planner routes foreign zstd frames to (tests/zstd_foreign.py): fast 4/SM, 2/SM, 1/SM, general."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import zstd_foreign as zf
from oracle_binding import Oracle
from rawnanoporesignalcompression_b200.codec import Codec

o = Oracle(); cd = Codec(0)
for name, x in zf.signals():
    for level, ck in zf.SETTINGS:
        for codec, make in ((0, zf.vbz_frame), (6, zf.c5_container)):
            got, st = cd.decompress_batch([make(o, x, level, ck)] * 8, [x.size] * 8, codec)
            assert not st.any() and all((g == x).all() for g in got)
            print(name, "level", level, "checksum", ck, "VBZ" if codec == 0 else "C5", "classes", tuple(cd.decode_class_counts()), flush=True)
