"""Pins the CPU oracle (oracle/rnsc_oracle.c + zstd_decode_oracle.c) against golden vectors produced by the
reference's own code (tests/golden/make_golden.py) and, when built, against oracle/_ref directly."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import recipes

CODECS = list(range(8))


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def test_zstd_version_matches_golden(oracle, golden):
    # whole-chunk encode bytes are only defined against the libzstd that made the golden vectors
    assert oracle.zstd_version() == golden["zstd_version"]


def test_real_chunks_decode_and_encode(oracle, golden, real_chunks):
    assert len(real_chunks) == 22 == len(golden["real"])
    total = 0
    for (frame, n), g in zip(real_chunks, golden["real"]):
        assert sha(frame) == g["stored_frame_sha256"] and n == g["n"]
        for own in (False, True):  # libzstd and the RFC 8878 restatement
            x = oracle.decode(0, frame, n, own_zstd=own)
            assert sha(x.tobytes()) == g["samples_sha256"]
        # stored frame == zstd(svb16(samples)): SVB-stage bytes are version independent
        assert oracle.zstd_decode(frame) == b"".join(oracle.streams(0, x)[:2])
        for codec in CODECS:
            ge = g["codecs"][str(codec)]
            st, b = oracle.encode_status(codec, x)
            assert st == ge["status"]
            if st == 0:
                assert len(b) == ge["size"] and sha(b) == ge["sha256"], f"codec {codec}"
                for own in (False, True):
                    if codec == 7 and own:
                        continue
                    assert (oracle.decode(codec, b, n, own_zstd=own) == x).all()
        total += n
    assert total == 1_548_931


@pytest.mark.parametrize("name_x", recipes.named_recipes(), ids=lambda p: p[0])
def test_recipes(oracle, golden, name_x):
    name, x = name_x
    g = golden["recipes"][name]
    assert sha(x.tobytes()) == g["samples_sha256"], "recipe drifted"
    for codec in CODECS:
        ge = g["codecs"][str(codec)]
        st, b = oracle.encode_status(codec, x)
        assert st == ge["status"], f"codec {codec}"
        if st == 0:
            assert len(b) == ge["size"] and sha(b) == ge["sha256"], f"codec {codec}"
            assert (oracle.decode(codec, b, x.size, own_zstd=(codec != 7)) == x).all()


def test_against_reference_build(oracle, ref_oracle):
    rng = np.random.default_rng(2026)
    for n in [0, 1, 2, 7, 8, 9, 100, 1000, 5000, 30000]:
        for x in (recipes.walk(n, n + 1), rng.integers(-300, 300, n).astype(np.int16), recipes.extremes(n, n + 2)):
            for codec in CODECS:
                rs, rb = ref_oracle.encode_status(codec, x)
                os_, ob = oracle.encode_status(codec, x)
                assert rs == os_, (n, codec)
                if rs == 0:
                    assert rb == ob, (n, codec)
                    st, y = ref_oracle.decode_status(codec, ob, n)
                    assert st == 0 and (y == x).all()


def test_error_semantics(oracle):
    x = recipes.walk(3000, 5)
    for codec in range(7):
        b = bytearray(oracle.encode(codec, x))
        # not a zstd frame
        bad = bytes(b)
        pos = 0 if codec < 2 else 8
        bad = bad[:pos] + b"\x00\x00\x00\x00" + bad[pos + 4:]
        assert oracle.decode_status(codec, bad, x.size)[0] == 1
        # wrong sample count -> "Remaining data at end of signal buffer"
        assert oracle.decode_status(codec, bytes(b), x.size - 8)[0] == 4
    # empty chunk: 9-byte empty frames, C5 = 77 bytes
    assert len(oracle.encode(6, np.zeros(0, np.int16))) == 77
    assert len(oracle.encode(0, np.zeros(0, np.int16))) == 9
    # incompressible full chunk: C4/C5 refuse
    u = recipes.uniform16(102400, 11)
    assert oracle.encode_status(6, u)[0] == 3 and oracle.encode_status(5, u)[0] == 3


def test_oracle_zstd_decoder_on_foreign_frames(oracle):
    """The oracle's own RFC 8878 frame decoder against frames written at other levels / with checksums / two blocks,
    checked against libzstd's decode of the same bytes and the source samples."""
    import zstd_foreign as zf
    for name, x in zf.signals():
        for level, ck in zf.SETTINGS:
            f = zf.vbz_frame(oracle, x, level, ck)
            for own in (True, False):
                st, y = oracle.decode_status(0, f, x.size, own_zstd=own)
                assert st == 0 and (y == x).all(), (name, level, ck, own)
        c = zf.c5_container(oracle, x, 19, True)
        for own in (True, False):
            st, y = oracle.decode_status(6, c, x.size, own_zstd=own)
            assert st == 0 and (y == x).all(), (name, "C5", own)
