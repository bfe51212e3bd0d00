"""Streaming POD5 repack: the caller-side batching that lets whole files flow through the GPU codec in bounded memory.

What the reference does per read (`copy.cpp:136-185`: `pod5_get_read_complete_signal` -> `pod5_add_reads_data`, one codec
call per 102,400-sample chunk) and what its repacker / loader do with batches (`pod5_format_pybind/repacker.h:232-245`:
signal rows moved between files as record batches; `async_signal_loader.cpp:112-174`: worker threads prefetch the next
batch while the consumer handles the current one; `file_writer.h:22-24`: 100 signal rows per record batch) is done here
at record-batch granularity:

  reader thread   record batches of the input's signal table are read one at a time from the embedded Arrow IPC file
                  (no `to_pylist`, no whole-file read) and gathered into GROUPS of about `group_samples` samples:
                  compressed bytes / offsets / sample counts are taken zero-copy from the Arrow buffers;
  codec           one batched call per group: `rnsc_decompress_batch` + `rnsc_compress_batch` (host-libzstd entropy stage:
                  the reference's bytes), or, with the GPU entropy stage, a TRANSCODE that never brings samples to the host:
                  compressed bytes H2D -> rnsc_dev_decode -> rnsc_dev_encode -> compressed bytes D2H;
  writer thread   the re-encoded chunks go out as Arrow record batches with the input's row grouping, built zero-copy over
                  the codec's output buffer, into an embedded IPC file of the output container.

Group k+1 is read while group k is in the codec and group k-1 is written (three stages, two in flight), so resident
memory is a few groups whatever the file size. Reads and run-info tables pass through byte for byte.
"""
from __future__ import annotations

import io
import os
import struct
import uuid
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
import pyarrow as pa
import pyarrow.ipc as ipc

from . import pod5_container as pc

DEFAULT_GROUP_SAMPLES = 32 << 20          # ~64 MiB of int16 per group


class _SubFile(io.RawIOBase):
    """Read-only window [base, base + length) of a file, as a seekable file object (for pyarrow.PythonFile)."""

    def __init__(self, f, base: int, length: int):
        super().__init__()
        self.f, self.base, self.length, self.pos = f, base, length, 0

    def readable(self):
        return True

    def seekable(self):
        return True

    def tell(self):
        return self.pos

    def seek(self, off, whence=0):
        self.pos = off if whence == 0 else (self.pos + off if whence == 1 else self.length + off)
        return self.pos

    def read(self, n=-1):
        if n is None or n < 0:
            n = self.length - self.pos
        n = max(0, min(n, self.length - self.pos))
        self.f.seek(self.base + self.pos)
        b = self.f.read(n)
        self.pos += len(b)
        return b

    def readinto(self, buf):
        b = self.read(len(buf))
        buf[: len(b)] = b
        return len(b)


class _SubWriter(io.RawIOBase):
    """Append-only window of a file whose tell() counts from `base`: Arrow's IPC file writer records block offsets from the
    sink's position, which must be relative to the start of the EMBEDDED file."""

    def __init__(self, f, base: int):
        super().__init__()
        self.f, self.base = f, base

    def writable(self):
        return True

    def tell(self):
        return self.f.tell() - self.base

    def write(self, b):
        return self.f.write(b)

    def flush(self):
        self.f.flush()


@dataclass
class Group:
    batches: List[pa.RecordBatch]
    rows: List[int]                         # rows per input record batch
    comp: Optional[np.ndarray]              # compressed input bytes of the group, chunks back to back (None: uncompressed input)
    comp_off: Optional[np.ndarray]          # uint64 [nchunks+1]
    samples_flat: Optional[np.ndarray]      # uncompressed input: int16 samples back to back
    n: np.ndarray                           # uint32 [nchunks]


def read_footer(f) -> Tuple[pc.Footer, bytes]:
    f.seek(0, os.SEEK_END)
    size = f.tell()
    f.seek(0)
    head = f.read(24)
    f.seek(size - 32)
    tail = f.read(32)
    if size < 64 or head[:8] != pc.SIGNATURE or tail[-8:] != pc.SIGNATURE:
        raise pc.Pod5Error("Invalid signature in file")
    flen = struct.unpack_from("<q", tail, 0)[0]
    end = size - 32
    if flen < 0 or flen > end:
        raise pc.Pod5Error("Invalid footer length")
    f.seek(end - flen)
    return pc.parse_footer(f.read(flen)), head[8:24]


def _binary_views(col: pa.Array) -> Tuple[np.ndarray, np.ndarray]:
    """(int64 offsets [len+1], uint8 data) of a large_binary array, zero-copy views of its Arrow buffers."""
    if isinstance(col, pa.ExtensionArray):
        col = col.storage
    bufs = col.buffers()
    offs = np.frombuffer(bufs[1], dtype=np.int64)[col.offset : col.offset + len(col) + 1]
    data = np.frombuffer(bufs[2], dtype=np.uint8) if bufs[2] is not None else np.zeros(0, np.uint8)
    return offs, data


def _gather_group(batches: List[pa.RecordBatch], ext: str) -> Group:
    n = np.concatenate([b.column("samples").to_numpy(zero_copy_only=False).astype(np.uint32, copy=False) for b in batches]) if batches else np.zeros(0, np.uint32)
    rows = [b.num_rows for b in batches]
    if ext in (pc.EXT_VBZ, pc.EXT_PGNANO):
        parts, sizes = [], []
        for b in batches:
            offs, data = _binary_views(b.column("signal"))
            parts.append(data[int(offs[0]) : int(offs[-1])])
            sizes.append(np.diff(offs))
        sizes = np.concatenate(sizes) if sizes else np.zeros(0, np.int64)
        coff = np.zeros(sizes.size + 1, np.uint64)
        np.cumsum(sizes, out=coff[1:])
        comp = np.concatenate(parts + [np.zeros(64, np.uint8)])          # one copy of the compressed bytes (+ the decoder's read slack)
        return Group(batches, rows, comp, coff, None, n)
    parts = []
    for b in batches:                                                   # large_list<int16>: values buffer, zero-copy
        col = b.column("signal")
        vals = col.values.to_numpy(zero_copy_only=False)
        o = col.offsets.to_numpy(zero_copy_only=False)
        parts.append(vals[int(o[0]) : int(o[-1])])
    flat = np.concatenate(parts).astype(np.int16, copy=False) if parts else np.zeros(0, np.int16)
    return Group(batches, rows, None, None, flat, n)


class StreamConverter:
    def __init__(self, codec, comp: str, variant: str = "C5", group_samples: int = DEFAULT_GROUP_SAMPLES, entropy: str = "libzstd"):
        from .cli import PGNANO_COMPRESSION, UNCOMPRESSED, VBZ_COMPRESSION
        self.codec, self.comp, self.variant, self.group_samples, self.entropy = codec, comp, variant, int(group_samples), entropy
        self.out_codec = None if comp == UNCOMPRESSED else ("VBZ" if comp == VBZ_COMPRESSION else variant)
        self.out_ext = "" if comp == UNCOMPRESSED else (pc.EXT_VBZ if comp == VBZ_COMPRESSION else pc.EXT_PGNANO)
        assert comp in (PGNANO_COMPRESSION, UNCOMPRESSED, VBZ_COMPRESSION)
        self.bytes_out = 0
        self.samples = 0
        self.groups = 0
        self._torch = None
        self.seconds = {"read": 0.0, "codec": 0.0, "write": 0.0}      # busy time per stage (the stages overlap)

    # ---- codec stage ----------------------------------------------------------------------------------------
    def _transcode_on_device(self, g: Group, in_codec: str):
        """compressed -> compressed without bringing samples to the host (GPU entropy stage on the way out)."""
        import torch
        if self._torch is None:
            self._torch = torch
            torch.cuda.set_device(0)
        dev = torch.device("cuda", torch.cuda.current_device())
        nch = g.n.size
        total = int(g.n.sum(dtype=np.int64))
        soff = np.zeros(nch, np.int64)
        if nch > 1:
            np.cumsum(g.n[:-1].astype(np.int64), out=soff[1:])
        d_comp = torch.from_numpy(g.comp).to(dev, non_blocking=False)
        d_coff = torch.from_numpy(g.comp_off.astype(np.int64)).to(dev)
        d_n = torch.from_numpy(g.n.astype(np.int32)).to(dev)
        d_soff = torch.from_numpy(soff).to(dev)
        d_sig = torch.empty(max(total, 1), dtype=torch.int16, device=dev)
        d_st = torch.empty(nch, dtype=torch.int32, device=dev)
        self.codec.dev_decode(in_codec, d_comp, d_coff, d_n, d_soff, int(g.n.max()) if nch else 0, d_sig, d_st)
        d_out = torch.empty(2 * total + total // 4 + 1104 * nch + 64, dtype=torch.uint8, device=dev)
        d_ooff = torch.empty(nch + 1, dtype=torch.int64, device=dev)
        d_st2 = torch.empty(nch, dtype=torch.int32, device=dev)
        self.codec.dev_encode(self.out_codec, d_sig, d_soff, d_n, total, d_out, d_ooff, d_st2)
        st = torch.maximum(d_st, d_st2).cpu().numpy()
        ooff = d_ooff.cpu().numpy()
        out = d_out[: int(ooff[-1])].cpu().numpy()
        return out, ooff, st

    def _codec_stage(self, g: Group, in_ext: str):
        """-> (kind, payload): ('binary', (bytes u8, offsets int64 [nchunks+1])) or ('samples', int16 flat)."""
        in_codec = "VBZ" if in_ext == pc.EXT_VBZ else (self.variant if in_ext == pc.EXT_PGNANO else None)
        nch = g.n.size
        if in_codec is not None and self.out_codec is not None and self.entropy == "gpu" and nch:
            out, ooff, st = self._transcode_on_device(g, in_codec)
        else:
            if in_codec is not None:
                flat, _, st = self.codec.decompress_flat(g.comp, g.comp_off, g.n, in_codec)
                if st.any():
                    bad = int(np.flatnonzero(st)[0])
                    raise pc.Pod5Error(f"signal row {bad} of a group: {self.codec.status_string(int(st[bad]))}")
            else:
                flat = g.samples_flat
            if self.out_codec is None:
                return "samples", flat
            soff = np.zeros(nch, np.uint64)
            if nch > 1:
                np.cumsum(g.n[:-1].astype(np.uint64), out=soff[1:])
            out, ooff, st = self.codec.compress_flat(flat, soff, g.n, self.out_codec)
            ooff = ooff.astype(np.int64)
        if st.any():
            bad = int(np.flatnonzero(st)[0])
            raise pc.Pod5Error(f"signal row {bad} of a group: {self.codec.status_string(int(st[bad]))}")
        return "binary", (np.ascontiguousarray(out), np.asarray(ooff, dtype=np.int64))

    # ---- output batches, zero-copy over the codec's output ---------------------------------------------------
    def _out_batches(self, g: Group, kind: str, payload, schema: pa.Schema) -> List[pa.RecordBatch]:
        res, row = [], 0
        for b, rows in zip(g.batches, g.rows):
            if kind == "binary":
                data, off = payload
                o = off[row : row + rows + 1]
                rel = np.ascontiguousarray(o - o[0])
                piece = data[int(o[0]) : int(o[-1])]
                col = pa.Array.from_buffers(pa.large_binary(), rows, [None, pa.py_buffer(rel), pa.py_buffer(piece)])
                self.bytes_out += int(o[-1] - o[0])
            else:
                flat = payload
                cnt = g.n[row : row + rows].astype(np.int64)
                start = int(g.n[:row].sum(dtype=np.int64))
                rel = np.zeros(rows + 1, np.int64)
                np.cumsum(cnt, out=rel[1:])
                vals = pa.Array.from_buffers(pa.int16(), int(rel[-1]), [None, pa.py_buffer(flat[start : start + int(rel[-1])])])
                col = pa.LargeListArray.from_arrays(pa.array(rel, type=pa.int64()), vals)
                self.bytes_out += 2 * int(rel[-1])
            res.append(pa.RecordBatch.from_arrays([b.column("read_id"), col, b.column("samples")], schema=schema))
            row += rows
        self.samples += int(g.n.sum(dtype=np.int64))
        self.groups += 1
        return res

    # ---- the whole file -------------------------------------------------------------------------------------
    def convert(self, in_path: str, out_path: str) -> Tuple[int, int]:
        with open(in_path, "rb") as fin:
            footer, marker = read_footer(fin)
            emb = {e.content_type: e for e in footer.contents}
            if pc.CONTENT_SIGNAL not in emb:
                raise pc.Pod5Error("no signal table in file")
            se = emb[pc.CONTENT_SIGNAL]
            reader = ipc.open_file(pa.PythonFile(_SubFile(fin, se.offset, se.length), mode="r"))
            in_schema = reader.schema
            in_ext = pc.signal_extension_name(in_schema)
            if self.out_codec is None:
                sig_field = pa.field("signal", pa.large_list(pa.int16()))
            else:
                sig_field = pa.field("signal", pa.large_binary(), metadata={"ARROW:extension:name": self.out_ext, "ARROW:extension:metadata": ""})
            out_schema = pa.schema([in_schema.field("read_id"), sig_field, in_schema.field("samples")], metadata=in_schema.metadata)
            nb = reader.num_record_batches

            import time as _t

            def read_group(first: int):
                t0 = _t.perf_counter()
                batches, samples, i = [], 0, first
                while i < nb and (not batches or samples < self.group_samples):
                    b = reader.get_batch(i)
                    batches.append(b)
                    samples += int(np.asarray(b.column("samples").to_numpy(zero_copy_only=False), dtype=np.int64).sum())
                    i += 1
                g = _gather_group(batches, in_ext)
                self.seconds["read"] += _t.perf_counter() - t0
                return g, i

            def write_group(bs):
                t0 = _t.perf_counter()
                for b in bs:
                    writer.write_batch(b)
                self.seconds["write"] += _t.perf_counter() - t0

            if os.path.exists(out_path):
                os.remove(out_path)                                       # copy.cpp:113-120
            contents = []
            with open(out_path, "wb") as fout, ThreadPoolExecutor(max_workers=2) as pool:
                fout.write(pc.SIGNATURE + marker)
                base = fout.tell()
                sink = pa.PythonFile(_SubWriter(fout, base), mode="w")
                writer = ipc.new_file(sink, out_schema)
                pending_write = None
                nxt = pool.submit(read_group, 0) if nb else None
                while nxt is not None:
                    g, i = nxt.result()
                    nxt = pool.submit(read_group, i) if i < nb else None      # group k+1 is read while group k is in the codec
                    t0 = _t.perf_counter()
                    kind, payload = self._codec_stage(g, in_ext)
                    out_b = self._out_batches(g, kind, payload, out_schema)
                    self.seconds["codec"] += _t.perf_counter() - t0
                    if pending_write is not None:
                        pending_write.result()                                 # group k-1 has left for the file
                    pending_write = pool.submit(write_group, out_b)
                if pending_write is not None:
                    pending_write.result()
                writer.close()
                sink.flush()
                end = fout.tell()
                contents.append(pc.EmbeddedFile(offset=base, length=end - base, format=0, content_type=pc.CONTENT_SIGNAL))
                fout.write(b"\0" * ((-end) % 8) + marker)
                for ct in (pc.CONTENT_RUN_INFO, pc.CONTENT_READS):             # the reference's order: signal, run info, reads (combined_file_utils.h:85-127)
                    if ct not in emb:
                        continue
                    e = emb[ct]
                    pos = fout.tell()
                    fin.seek(e.offset)
                    left = e.length
                    while left:
                        blk = fin.read(min(left, 64 << 20))
                        fout.write(blk)
                        left -= len(blk)
                    contents.append(pc.EmbeddedFile(offset=pos, length=e.length, format=0, content_type=ct))
                    fout.write(b"\0" * ((-fout.tell()) % 8) + marker)
                fout.write(pc.FOOTER_MAGIC)
                fb = pc.build_footer(pc.Footer(footer.file_identifier, footer.software, footer.pod5_version, contents))
                fout.write(fb)
                fout.write(b"\0" * ((-fout.tell()) % 8))
                fout.write(struct.pack("<q", len(fb)))
                fout.write(marker + pc.SIGNATURE)
        return self.bytes_out, self.samples


def convert_streaming(in_path: str, out_path: str, comp: str, variant: str = "C5", codec=None, group_samples: int = DEFAULT_GROUP_SAMPLES,
                      entropy: str = "libzstd") -> Tuple[int, int]:
    """Bounded-memory version of cli.convert: (compressed signal bytes written, samples)."""
    if codec is None:
        from .codec import default_codec
        codec = default_codec()
    sc = StreamConverter(codec, comp, variant, group_samples, entropy)
    res = sc.convert(in_path, out_path)
    if os.environ.get("RNSC_STREAM_STATS"):
        import sys
        print(f"[stream] groups {sc.groups}, busy seconds per stage (overlapped): " + ", ".join(f"{k} {v:.2f}" for k, v in sc.seconds.items()), file=sys.stderr)
    return res


def signal_checksum(path: str, variant: str = "C5", codec=None, group_samples: int = DEFAULT_GROUP_SAMPLES) -> Tuple[int, int, int]:
    """(crc32 of every decoded sample in file order, samples, signal rows): the signal of a file of any size, decoded group by
    group in bounded memory -- two files hold the same signal iff the triples agree (check_pod5_files_equal.py:60-64 at scale)."""
    import zlib
    if codec is None:
        from .codec import default_codec
        codec = default_codec()
    crc, total, rows = 0, 0, 0
    with open(path, "rb") as fin:
        footer, _ = read_footer(fin)
        se = {e.content_type: e for e in footer.contents}[pc.CONTENT_SIGNAL]
        reader = ipc.open_file(pa.PythonFile(_SubFile(fin, se.offset, se.length), mode="r"))
        ext = pc.signal_extension_name(reader.schema)
        in_codec = "VBZ" if ext == pc.EXT_VBZ else (variant if ext == pc.EXT_PGNANO else None)
        i, nb = 0, reader.num_record_batches
        while i < nb:
            batches, samples = [], 0
            while i < nb and (not batches or samples < group_samples):
                b = reader.get_batch(i); batches.append(b); i += 1
                samples += int(np.asarray(b.column("samples").to_numpy(zero_copy_only=False), dtype=np.int64).sum())
            g = _gather_group(batches, ext)
            if in_codec is not None:
                flat, _, st = codec.decompress_flat(g.comp, g.comp_off, g.n, in_codec)
                if st.any():
                    raise pc.Pod5Error("decode status in signal_checksum")
            else:
                flat = g.samples_flat
            crc = zlib.crc32(memoryview(np.ascontiguousarray(flat)).cast("B"), crc)
            total += int(flat.size); rows += int(g.n.size)
    return crc, total, rows


def write_synthetic_pod5(path: str, codec, n_reads: int, seed: int, model, codec_name: str = "VBZ", template: Optional[str] = None,
                         reads_per_group: int = 2048, device=None, entropy: str = "libzstd") -> Tuple[int, int]:
    """A POD5 file of synthetic reads (synth.generate_reads), written group by group: the test input of the streaming
    repack. The reads table cycles the per-read metadata of `template` (a real POD5 file) with fresh read ids and the right
    signal-row lists, so that the file is consistent. Returns (bytes of the file, samples)."""
    from . import synth
    tpl = pc.read_pod5(template) if template else None
    lengths = synth.read_lengths(n_reads, seed)
    ext = pc.EXT_VBZ if codec_name.upper() == "VBZ" else pc.EXT_PGNANO
    uu = pa.field("read_id", pa.binary(16), metadata={"ARROW:extension:name": pc.EXT_UUID, "ARROW:extension:metadata": ""})
    md = tpl.signal.schema.metadata if tpl else {b"MINKNOW:file_identifier": str(uuid.UUID(int=seed)).encode(), b"MINKNOW:software": b"rnsc-b200", b"MINKNOW:pod5_version": b"0.0.40"}
    schema = pa.schema([uu, pa.field("signal", pa.large_binary(), metadata={"ARROW:extension:name": ext, "ARROW:extension:metadata": ""}), pa.field("samples", pa.uint32())], metadata=md)
    marker = uuid.UUID(int=seed ^ 0x5A5A).bytes
    ids = [uuid.UUID(int=(seed << 64) | r).bytes for r in range(n_reads)]
    prev = codec.lib.rnsc_get_entropy_stage(codec.ctx)
    codec.set_entropy_stage(entropy)
    try:
        with open(path, "wb") as f:
            f.write(pc.SIGNATURE + marker)
            base = f.tell()
            sink = pa.PythonFile(_SubWriter(f, base), mode="w")
            w = ipc.new_file(sink, schema)
            row0, total, rows_of_read = 0, 0, []
            for a in range(0, n_reads, reads_per_group):
                b = min(n_reads, a + reads_per_group)
                sig = synth.generate_reads(lengths[a:b], seed, model, first_read=a, device=device).cpu().numpy()
                off, n, read_of = synth.chunk_table(lengths[a:b])
                out, coff, st = codec.compress_flat(sig, off.astype(np.uint64), n, codec_name)
                if st.any():
                    raise pc.Pod5Error("encoder status while writing the synthetic file")
                coff = coff.astype(np.int64)
                rid = pa.array([ids[a + int(r)] for r in read_of], type=pa.binary(16))
                for r0 in range(0, n.size, 100):
                    r1 = min(n.size, r0 + 100)
                    o = coff[r0 : r1 + 1]
                    col = pa.Array.from_buffers(pa.large_binary(), r1 - r0, [None, pa.py_buffer(np.ascontiguousarray(o - o[0])), pa.py_buffer(out[int(o[0]) : int(o[-1])])])
                    w.write_batch(pa.RecordBatch.from_arrays([rid.slice(r0, r1 - r0), col, pa.array(n[r0:r1], type=pa.uint32())], schema=schema))
                per = np.bincount(read_of, minlength=b - a)
                start = row0 + np.concatenate([[0], np.cumsum(per)[:-1]])
                rows_of_read += [list(range(int(s), int(s + p))) for s, p in zip(start, per)]
                row0 += n.size
                total += int(n.sum(dtype=np.int64))
            w.close()
            sink.flush()
            end = f.tell()
            contents = [pc.EmbeddedFile(offset=base, length=end - base, format=0, content_type=pc.CONTENT_SIGNAL)]
            f.write(b"\0" * ((-end) % 8) + marker)
            if tpl is not None:
                for ct in (pc.CONTENT_RUN_INFO, pc.CONTENT_READS):
                    if ct == pc.CONTENT_READS:
                        t = tpl.reads
                        idx = np.arange(n_reads) % t.num_rows
                        t = t.take(pa.array(idx))
                        cols = {name: t.column(name) for name in t.schema.names}
                        cols["read_id"] = pa.chunked_array([pa.array(ids, type=pa.binary(16))])
                        cols["signal"] = pa.chunked_array([pa.array(rows_of_read, type=pa.list_(pa.uint64()))])
                        if "num_samples" in cols:
                            cols["num_samples"] = pa.chunked_array([pa.array(lengths.astype(np.uint64))])
                        t = pa.Table.from_arrays([cols[nm].combine_chunks() for nm in t.schema.names], schema=t.schema)
                        blob = pc.table_to_ipc_bytes(t, 1000)
                    else:
                        blob = tpl.raw[ct]
                    pos = f.tell()
                    f.write(blob)
                    contents.append(pc.EmbeddedFile(offset=pos, length=len(blob), format=0, content_type=ct))
                    f.write(b"\0" * ((-f.tell()) % 8) + marker)
            f.write(pc.FOOTER_MAGIC)
            fb = pc.build_footer(pc.Footer(md[b"MINKNOW:file_identifier"].decode(), md[b"MINKNOW:software"].decode(), md[b"MINKNOW:pod5_version"].decode(), contents))
            f.write(fb)
            f.write(b"\0" * ((-f.tell()) % 8))
            f.write(struct.pack("<q", len(fb)))
            f.write(marker + pc.SIGNATURE)
            size = f.tell()
    finally:
        codec.set_entropy_stage(prev)
    return size, total
