"""COMPRESSOR <input.pod5> <output.pod5> <--uncompressed | --VBZ | --pgnano> [--variant C5]

The reference's converter CLI (src/c++/copy.cpp:18-62,87-190; binaries utils/compressors/<variant>) on top of
librnsc: every signal chunk of the input is decoded on the GPU with the codec its signal column declares
(minknow.vbz / pgnano.signal / uncompressed, signal_table_schema.cpp:44-74), re-encoded with the requested one
and written into a new POD5 container; reads and run-info tables pass through unchanged. Same argument grammar
as the reference: the third argument is matched by PREFIX (strncmp), relative paths are resolved against the
working directory, an existing output file is replaced, exit status 0 / 1.

The pgnano variant, fixed at compile time in the reference (pgnano/pgnano.cpp:1), is `--variant` here (default
C5, the reference's checked-in build) and applies to both reading and writing pgnano.signal columns.

`files_equal` restates the reference's check script (pod5/python/pod5/src/pod5/test_utils/
check_pod5_files_equal.py:9-71): same reads in the same order, same per-read fields, equal signal.
"""
from __future__ import annotations

import os
import sys
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import pyarrow as pa

from . import pod5_container as pc

SIGNAL_ROWS_PER_BATCH = 100          # file_writer.h:22-24
UNCOMPRESSED, VBZ_COMPRESSION, PGNANO_COMPRESSION = "uncompressed", "VBZ", "pgnano"


class UsageError(SystemExit):
    def __init__(self, msg: str):
        print(msg, file=sys.stderr)
        super().__init__(1)


OPTIONS = {"entropy": "libzstd", "group_samples": 0, "whole_file": False}      # extensions picked up by main(): --entropy libzstd|gpu, --group-samples N, --whole-file


def parse_args(argv: Sequence[str], cwd: Optional[str] = None) -> Tuple[str, str, str, str]:
    """(input path, output path, compression, pgnano variant). copy.cpp:18-62 plus the --variant extension (and, for main(),
    --entropy libzstd|gpu: who writes the zstd frames; --group-samples N: samples per streaming group; --whole-file: the
    non-streaming converter)."""
    args = list(argv)
    variant = "C5"
    OPTIONS.update(entropy="libzstd", group_samples=0, whole_file=False)
    for flag, key, conv in (("--entropy", "entropy", str), ("--group-samples", "group_samples", int)):
        if flag in args:
            i = args.index(flag)
            if i + 1 >= len(args):
                raise UsageError(f"{flag} needs a value")
            try:
                OPTIONS[key] = conv(args[i + 1])
            except ValueError:
                raise UsageError(f"bad value for {flag}") from None
            del args[i : i + 2]
    if OPTIONS["entropy"] not in ("libzstd", "gpu"):
        raise UsageError("--entropy must be libzstd or gpu")
    if "--whole-file" in args:
        args.remove("--whole-file"); OPTIONS["whole_file"] = True
    if "--variant" in args:
        i = args.index("--variant")
        if i + 1 >= len(args):
            raise UsageError("--variant needs a value (VBZ0, C1..C5)")
        variant = args[i + 1].upper()
        del args[i : i + 2]
    if variant not in ("VBZ0", "C1", "C2", "C3", "C4", "C5"):
        raise UsageError("Incorrect pgnano variant")
    if len(args) > 3:
        print("Ignoring extra arguments (only first 3 considered)", file=sys.stderr)
    if len(args) == 0:
        raise UsageError("No input file specified")
    if len(args) == 1:
        raise UsageError("No output file specified")
    if len(args) == 2:
        raise UsageError("Incorrect compression method")
    flag = args[2]
    if flag[:5] == "--VBZ":
        comp = VBZ_COMPRESSION
    elif flag[:14] == "--uncompressed":
        comp = UNCOMPRESSED
    elif flag[:8] == "--pgnano":
        comp = PGNANO_COMPRESSION
    else:
        raise UsageError("Incorrect compression method")
    cwd = cwd or os.getcwd()
    paths = [p if os.path.isabs(p) else os.path.join(cwd, p) for p in args[:2]]
    return paths[0], paths[1], comp, variant


def _codec_of_column(ext_name: str, variant: str):
    if ext_name == pc.EXT_VBZ:
        return "VBZ"
    if ext_name == pc.EXT_PGNANO:
        return variant
    return None


def decode_signal_table(table: pa.Table, codec, variant: str) -> Tuple[List[np.ndarray], np.ndarray]:
    """All chunks of a signal table as int16 arrays (one batched GPU call) + the `samples` column."""
    samples = np.asarray(table.column("samples").to_numpy(), dtype=np.uint32)
    cname = _codec_of_column(pc.signal_extension_name(table), variant)
    col = table.column("signal")
    if cname is None:
        return [np.asarray(v, dtype=np.int16) for v in col.to_pylist()], samples
    chunks, status = codec.decompress_batch(col.to_pylist(), samples, cname)
    if status.any():
        bad = int(np.flatnonzero(status)[0])
        raise pc.Pod5Error(f"signal row {bad}: {codec.status_string(int(status[bad]))}")
    return chunks, samples


def encode_signal_column(chunks: List[np.ndarray], codec, comp: str, variant: str) -> Tuple[pa.Array, pa.Field]:
    if comp == UNCOMPRESSED:
        arr = pa.array([c.tolist() for c in chunks], type=pa.large_list(pa.int16()))
        return arr, pa.field("signal", arr.type)
    cname = "VBZ" if comp == VBZ_COMPRESSION else variant
    blobs, status = codec.compress_batch(chunks, cname)
    if status.any():
        bad = int(np.flatnonzero(status)[0])
        raise pc.Pod5Error(f"signal row {bad}: {codec.status_string(int(status[bad]))}")
    ext = pc.EXT_VBZ if comp == VBZ_COMPRESSION else pc.EXT_PGNANO
    fld = pa.field("signal", pa.large_binary(), metadata={"ARROW:extension:name": ext, "ARROW:extension:metadata": ""})
    return pa.array(blobs, type=pa.large_binary()), fld


def convert(in_path: str, out_path: str, comp: str, variant: str = "C5", codec=None) -> Tuple[int, int]:
    """The body of copy.cpp's main loop (:136-185). Returns (compressed signal bytes, samples)."""
    if codec is None:
        from .codec import default_codec
        codec = default_codec()
    src = pc.read_pod5(in_path)
    sig = src.signal
    chunks, samples = decode_signal_table(sig, codec, variant)
    column, fld = encode_signal_column(chunks, codec, comp, variant)
    schema = pa.schema([sig.schema.field("read_id"), fld, sig.schema.field("samples")], metadata=sig.schema.metadata)
    out_sig = pa.Table.from_arrays([sig.column("read_id"), column, sig.column("samples")], schema=schema)
    tables = dict(src.raw)
    tables[pc.CONTENT_SIGNAL] = pc.table_to_ipc_bytes(out_sig, SIGNAL_ROWS_PER_BATCH)
    if os.path.exists(out_path):
        os.remove(out_path)                       # copy.cpp:113-120
    pc.write_pod5(out_path, tables, src.footer.file_identifier, src.footer.software, src.footer.pod5_version, src.section_marker)
    nbytes = int(sum(len(b) for b in column.to_pylist())) if comp != UNCOMPRESSED else int(2 * samples.sum())
    return nbytes, int(samples.sum())


def _column_equal(x: pa.ChunkedArray, y: pa.ChunkedArray) -> bool:
    if x.type != y.type or len(x) != len(y):
        return False
    if pa.types.is_floating(x.type):                      # NaN == NaN here (median_before is NaN when unknown)
        return bool(np.array_equal(x.to_numpy(), y.to_numpy(), equal_nan=True))
    return x.to_pylist() == y.to_pylist()


def files_equal(path_a: str, path_b: str, variant: str = "C5", codec=None) -> Tuple[bool, str]:
    """check_pod5_files_equal.py:19-64 on top of the pyarrow container reader."""
    if codec is None:
        from .codec import default_codec
        codec = default_codec()
    a, b = pc.read_pod5(path_a), pc.read_pod5(path_b)
    ra, rb = a.reads, b.reads
    if ra.num_rows != rb.num_rows:
        return False, f"read count {ra.num_rows} != {rb.num_rows}"
    for name in ("read_id", "read_number", "start", "median_before", "channel", "well", "pore_type", "calibration_offset",
                 "calibration_scale", "end_reason", "end_reason_forced", "run_info"):
        if name in ra.schema.names or name in rb.schema.names:
            if name not in ra.schema.names or name not in rb.schema.names:
                return False, f"column {name} missing in one file"
            if not _column_equal(ra.column(name), rb.column(name)):
                return False, f"reads column {name} differs"
    sa, _ = decode_signal_table(a.signal, codec, variant)
    sb, _ = decode_signal_table(b.signal, codec, variant)
    if a.signal.column("read_id").to_pylist() != b.signal.column("read_id").to_pylist():
        return False, "signal rows belong to different reads"
    if len(sa) != len(sb) or any(x.shape != y.shape or (x != y).any() for x, y in zip(sa, sb)):
        return False, "signal differs"
    if ra.column("signal").to_pylist() != rb.column("signal").to_pylist():
        return False, "signal row indices differ"
    return True, "files are equal"


def main(argv: Optional[Sequence[str]] = None) -> int:
    in_path, out_path, comp, variant = parse_args(sys.argv[1:] if argv is None else argv)
    try:
        if OPTIONS["whole_file"]:
            nbytes, nsamples = convert(in_path, out_path, comp, variant)
        else:       # record batches stream through the codec in bounded memory (stream_convert.py)
            from .stream_convert import DEFAULT_GROUP_SAMPLES, convert_streaming
            nbytes, nsamples = convert_streaming(in_path, out_path, comp, variant, None, OPTIONS["group_samples"] or DEFAULT_GROUP_SAMPLES, OPTIONS["entropy"])
    except Exception as e:  # noqa: BLE001  (pod5_management_utils.h:32-37 prints the error and exits 1)
        print(f"Error: {e}", file=sys.stderr)
        return 1
    print(f"{nbytes};{nsamples}")                 # run_converters.py:24-35 expects `bytes;samples`
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
