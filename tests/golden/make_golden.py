"""Generates tests/golden/golden.json from the REFERENCE's own code (oracle/_ref, compiled from /root/reference
by oracle/Makefile) and copies the reference's POD5 fixture. Run in the build container only:

    python tests/golden/make_golden.py

Every entry stores sha256 digests (and sizes) of what the reference produced with this image's libzstd:
  * `real`: per stored VBZ chunk of pod5/test_data/multi_fast5_zip_v3.pod5: sha256 of the samples the reference
    decoder (SSE path) yields from the stored frame, and per codec the sha256 of the reference encoder's bytes.
  * `recipes`: the same for the synthetic recipes of tests/recipes.py (reference unit-test inputs + edge cases),
    including the status code when the reference refuses the chunk.
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle_binding import RefOracle  # noqa: E402
import recipes  # noqa: E402
from rawnanoporesignalcompression_b200.pod5_container import read_pod5  # noqa: E402

SRC = "/root/reference/pod5/test_data/multi_fast5_zip_v3.pod5"


def sha(b) -> str:
    return hashlib.sha256(bytes(b)).hexdigest()


def entry(ref, x):
    e = {"n": int(x.size), "samples_sha256": sha(np.ascontiguousarray(x, dtype=np.int16).tobytes()), "codecs": {}}
    for codec in range(8):
        st, b = ref.encode_status(codec, x)
        ce = {"status": int(st)}
        if st == 0:
            ce.update(size=len(b), sha256=sha(b))
            rs, y = ref.decode_status(codec, b, x.size)
            assert rs == 0 and (y == x).all(), "reference round trip failed"
        e["codecs"][str(codec)] = ce
    return e


def main():
    assert RefOracle.available(), "build oracle/_ref first (make -f oracle/Makefile)"
    ref = RefOracle()
    shutil.copyfile(SRC, os.path.join(HERE, "multi_fast5_zip_v3.pod5"))
    sig = read_pod5(SRC).signal
    out = {"zstd_version": int(ref.lib.ref_zstd_version()), "source": "oracle/_ref (reference sources + libzstd of this image)", "real": [], "recipes": {}}
    for b, n in zip(sig.column("signal").to_pylist(), sig.column("samples").to_pylist()):
        st, x = ref.decode_status(0, b, n)
        assert st == 0
        e = entry(ref, x)
        e["stored_frame_sha256"] = sha(b)
        out["real"].append(e)
    for name, x in recipes.named_recipes():
        out["recipes"][name] = entry(ref, x)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote golden.json:", len(out["real"]), "real chunks,", len(out["recipes"]), "recipes")


if __name__ == "__main__":
    main()
