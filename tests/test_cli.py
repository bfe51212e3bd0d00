"""COMPRESSOR CLI + POD5 container (BASELINE config 1: round trip of the reference's pod5/test_data fixture through
--pgnano and back through --VBZ, verified with the restated check_pod5_files_equal)."""
import os

import numpy as np
import pytest

from rawnanoporesignalcompression_b200 import cli, pod5_container as pc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE = os.path.join(ROOT, "tests", "golden", "multi_fast5_zip_v3.pod5")
NAMES = {"VBZ": 0, "VBZ0": 1, "C1": 2, "C2": 3, "C3": 4, "C4": 5, "C5": 6}


class OracleCodec:
    """CPU stand-in with the Codec batch interface (tests only): exercises the CLI / container logic without a GPU."""

    def __init__(self, oracle):
        self.o = oracle

    def status_string(self, s):
        return f"status {s}"

    def decompress_batch(self, chunks, sample_counts, codec):
        out = [self.o.decode(NAMES[codec], bytes(c), int(n)) for c, n in zip(chunks, sample_counts)]
        return out, np.zeros(len(out), np.int32)

    def compress_batch(self, chunks, codec):
        out = [self.o.encode(NAMES[codec], c) for c in chunks]
        return out, np.zeros(len(out), np.int32)


def test_parse_args_grammar(tmp_path):
    cwd = str(tmp_path)
    assert cli.parse_args(["a.pod5", "b.pod5", "--VBZ"], cwd) == (os.path.join(cwd, "a.pod5"), os.path.join(cwd, "b.pod5"), "VBZ", "C5")
    assert cli.parse_args(["/x/a.pod5", "b.pod5", "--pgnano-anything"], cwd)[2] == "pgnano"         # prefix match (strncmp, copy.cpp:37-53)
    assert cli.parse_args(["a", "b", "--uncompressed", "ignored"], cwd)[2] == "uncompressed"
    assert cli.parse_args(["a", "b", "--pgnano", "--variant", "c3"], cwd)[3] == "C3"
    for bad in ([], ["a"], ["a", "b"], ["a", "b", "--zstd"], ["a", "b", "--VB"], ["a", "b", "--pgnano", "--variant", "C9"]):
        with pytest.raises(SystemExit) as e:
            cli.parse_args(bad, cwd)
        assert e.value.code == 1


def _round_trip(tmp_path, codec, variants):
    for variant in variants:
        a = str(tmp_path / f"a_{variant}.pod5"); b = str(tmp_path / f"b_{variant}.pod5"); u = str(tmp_path / f"u_{variant}.pod5")
        nbytes, nsamples = cli.convert(FIXTURE, a, "pgnano", variant, codec)
        assert nsamples == 1548931 and 0 < nbytes < 2 * nsamples
        t = pc.read_pod5(a)
        assert pc.signal_extension_name(t.signal) == pc.EXT_PGNANO and t.signal.num_rows == 22
        assert t.raw[pc.CONTENT_READS] == pc.read_pod5(FIXTURE).raw[pc.CONTENT_READS]            # metadata tables pass through
        cli.convert(a, b, "VBZ", variant, codec)                                                   # "decompress" = the same binary the other way
        ok, why = cli.files_equal(FIXTURE, b, variant, codec)
        assert ok, why
        # our VBZ bytes equal the reference's only up to the libzstd version; the samples must be identical
        cli.convert(a, u, "uncompressed", variant, codec)
        ok, why = cli.files_equal(FIXTURE, u, variant, codec)
        assert ok, why
    # existing output is replaced; a mismatch is reported
    ok, why = cli.files_equal(FIXTURE, str(tmp_path / f"a_{variants[0]}.pod5"), variants[0], codec)
    assert ok
    wrong = "C4" if variants[0] != "C4" else "C5"
    try:
        ok, _ = cli.files_equal(FIXTURE, str(tmp_path / f"a_{variants[0]}.pod5"), wrong, codec)
    except Exception:
        ok = False
    assert not ok                      # reading a pgnano file with the wrong variant cannot silently succeed


def test_cli_round_trip_cpu_stand_in(tmp_path, oracle):
    _round_trip(tmp_path, OracleCodec(oracle), ["C5"])


@pytest.mark.gpu
def test_cli_round_trip_gpu(tmp_path, gpu_codec):
    _round_trip(tmp_path, gpu_codec, ["C5", "VBZ0", "C1", "C2", "C3", "C4"])


@pytest.mark.gpu
def test_cli_main_exit_codes(tmp_path, gpu_codec, capsys):
    out = str(tmp_path / "o.pod5")
    assert cli.main([FIXTURE, out, "--pgnano"]) == 0
    bytes_, samples = capsys.readouterr().out.strip().split(";")
    assert int(samples) == 1548931 and int(bytes_) > 0
    assert cli.main([str(tmp_path / "missing.pod5"), out, "--VBZ"]) == 1


# ---- streaming repack (stream_convert.py): record batches through the codec in bounded memory -------------------
def _streaming_checks(tmp_path, codec, entropy_modes):
    from rawnanoporesignalcompression_b200 import stream_convert as sc
    a = str(tmp_path / "s_a.pod5"); b = str(tmp_path / "s_b.pod5"); u = str(tmp_path / "s_u.pod5"); ref = str(tmp_path / "s_ref.pod5")
    # tiny groups (40k samples): the 22 rows / 1 record batch of the fixture still go through the group machinery
    nbytes, nsamples = sc.convert_streaming(FIXTURE, a, "pgnano", "C5", codec, group_samples=40_000)
    assert nsamples == 1548931 and 0 < nbytes < 2 * nsamples
    cli.convert(FIXTURE, ref, "pgnano", "C5", codec)
    ta, tr = pc.read_pod5(a), pc.read_pod5(ref)
    assert ta.signal.equals(tr.signal)                                     # same bytes as the whole-file converter (host-libzstd entropy stage)
    assert ta.raw[pc.CONTENT_READS] == tr.raw[pc.CONTENT_READS] and ta.raw[pc.CONTENT_RUN_INFO] == tr.raw[pc.CONTENT_RUN_INFO]
    sc.convert_streaming(a, b, "VBZ", "C5", codec, group_samples=500_000)
    ok, why = cli.files_equal(FIXTURE, b, "C5", codec)
    assert ok, why
    sc.convert_streaming(a, u, "uncompressed", "C5", codec)
    ok, why = cli.files_equal(FIXTURE, u, "C5", codec)
    assert ok, why
    sc.convert_streaming(u, b, "pgnano", "C5", codec, group_samples=300_000)          # uncompressed input
    assert pc.read_pod5(b).signal.equals(tr.signal)
    for mode in entropy_modes:                                                         # device transcode: other bytes, same samples
        g = str(tmp_path / f"s_g_{mode}.pod5")
        sc.convert_streaming(FIXTURE, g, "pgnano", "C5", codec, group_samples=600_000, entropy=mode)
        ok, why = cli.files_equal(FIXTURE, g, "C5", codec)
        assert ok, why


def test_streaming_repack_on_emulated_kernels(tmp_path, emul_lib):
    from rawnanoporesignalcompression_b200.codec import Codec
    cd = Codec(0, lib_path=emul_lib)
    try:
        _streaming_checks(tmp_path, cd, [])
    finally:
        cd.close()


@pytest.mark.gpu
def test_streaming_repack_gpu(tmp_path, gpu_codec):
    _streaming_checks(tmp_path, gpu_codec, ["gpu"])
    # a multi-batch synthetic file written group by group, repacked by the device transcode and back
    from rawnanoporesignalcompression_b200 import stream_convert as sc, synth
    src = str(tmp_path / "syn.pod5"); p = str(tmp_path / "syn_p.pod5"); v = str(tmp_path / "syn_v.pod5")
    size, total = sc.write_synthetic_pod5(src, gpu_codec, 600, 20261018, synth.R10_4_1, "VBZ", template=FIXTURE, reads_per_group=128)
    assert pc.read_pod5(src).signal.num_rows > 600
    nb, ns = sc.convert_streaming(src, p, "pgnano", "C5", gpu_codec, group_samples=4_000_000, entropy="gpu")
    assert ns == total
    sc.convert_streaming(p, v, "VBZ", "C5", gpu_codec, group_samples=3_000_000)
    ok, why = cli.files_equal(src, v, "C5", gpu_codec)
    assert ok, why
