"""Minimal POD5 container reader/writer on top of pyarrow (no `pod5` package needed).

Layout (reference: pod5/c++/pod5_format/internal/combined_file_utils.h:21-62,130-151,188-222 and
pod5/docs/SPECIFICATION.md "Combined file layout"):

    signature(8) | section marker(16) | [embedded Arrow IPC file | pad to 8 | marker(16)]* |
    "FOOTER\\0\\0" | flatbuffer footer | pad to 8 | i64 footer length | marker(16) | signature(8)

The flatbuffer footer (pod5_format/flatbuffers/footer.fbs) is parsed and emitted by hand: table Footer
{file_identifier:string; software:string; pod5_version:string; contents:[EmbeddedFile]} with
EmbeddedFile {offset:int64; length:int64; format:short; content_type:short}.
"""
from __future__ import annotations

import struct
import uuid
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import pyarrow as pa
import pyarrow.ipc as ipc

SIGNATURE = b"\x8bPOD\r\n\x1a\n"
FOOTER_MAGIC = b"FOOTER\x00\x00"

CONTENT_READS, CONTENT_SIGNAL, CONTENT_READ_ID_INDEX, CONTENT_OTHER_INDEX, CONTENT_RUN_INFO = range(5)

# Arrow extension names of the signal column (pod5_format/types.h:53-111, signal_table_schema.cpp:44-74)
EXT_VBZ = "minknow.vbz"
EXT_PGNANO = "pgnano.signal"
EXT_UUID = "minknow.uuid"


class Pod5Error(IOError):
    pass


# ----------------------------------------------------------------------------------------------
# flatbuffer helpers (read)
# ----------------------------------------------------------------------------------------------
def _u16(b, o): return struct.unpack_from("<H", b, o)[0]
def _i16(b, o): return struct.unpack_from("<h", b, o)[0]
def _u32(b, o): return struct.unpack_from("<I", b, o)[0]
def _i32(b, o): return struct.unpack_from("<i", b, o)[0]
def _i64(b, o): return struct.unpack_from("<q", b, o)[0]


def _field(b: bytes, table: int, idx: int) -> int:
    """Absolute position of field `idx` of the table at `table`, or 0 when absent."""
    vt = table - _i32(b, table)
    vt_size = _u16(b, vt)
    slot = 4 + 2 * idx
    if slot >= vt_size:
        return 0
    off = _u16(b, vt + slot)
    return table + off if off else 0


def _string(b: bytes, pos: int) -> Optional[str]:
    if not pos:
        return None
    s = pos + _u32(b, pos)
    n = _u32(b, s)
    return b[s + 4 : s + 4 + n].decode("utf-8")


@dataclass
class EmbeddedFile:
    offset: int
    length: int
    format: int
    content_type: int


@dataclass
class Footer:
    file_identifier: str
    software: str
    pod5_version: str
    contents: List[EmbeddedFile] = field(default_factory=list)


def parse_footer(b: bytes) -> Footer:
    root = _u32(b, 0)
    ft = Footer(
        file_identifier=_string(b, _field(b, root, 0)) or "",
        software=_string(b, _field(b, root, 1)) or "",
        pod5_version=_string(b, _field(b, root, 2)) or "",
    )
    vpos = _field(b, root, 3)
    if vpos:
        vec = vpos + _u32(b, vpos)
        n = _u32(b, vec)
        for i in range(n):
            e = vec + 4 + 4 * i
            t = e + _u32(b, e)
            fo, fl, ff, fc = (_field(b, t, k) for k in range(4))
            ft.contents.append(
                EmbeddedFile(
                    offset=_i64(b, fo) if fo else 0,
                    length=_i64(b, fl) if fl else 0,
                    format=_i16(b, ff) if ff else 0,
                    content_type=_i16(b, fc) if fc else 0,
                )
            )
    return ft


# ----------------------------------------------------------------------------------------------
# flatbuffer helpers (write) -- a straightforward back-to-front builder for this one schema
# ----------------------------------------------------------------------------------------------
def build_footer(ft: Footer) -> bytes:
    """Serialise `ft` as a flatbuffer that `parse_footer` (and flatbuffers' own verifier) accepts.

    Built front-to-back with explicit positions: all offsets in flatbuffers are relative and
    unsigned-forward (uoffset: target > location), vtable soffset = table - vtable (may be either sign).
    """
    out = bytearray()

    def align(n):
        while len(out) % n:
            out.append(0)

    out += b"\0\0\0\0"  # root uoffset placeholder
    # --- root table: vtable first, then table
    align(2)
    vt_root = len(out)
    out += struct.pack("<HHHHHH", 12, 20, 4, 8, 12, 16)  # vtable size, table size, 4 field offsets
    align(4)
    t_root = len(out)
    out += struct.pack("<i", t_root - vt_root)
    fld_pos = [len(out) + 4 * i for i in range(4)]
    out += b"\0" * 16
    struct.pack_into("<I", out, 0, t_root)

    def put_string(slot, s):
        align(4)
        pos = len(out)
        raw = s.encode("utf-8")
        out.extend(struct.pack("<I", len(raw)) + raw + b"\0")
        struct.pack_into("<I", out, slot, pos - slot)

    put_string(fld_pos[0], ft.file_identifier)
    put_string(fld_pos[1], ft.software)
    put_string(fld_pos[2], ft.pod5_version)
    align(4)
    vec = len(out)
    struct.pack_into("<I", out, fld_pos[3], vec - fld_pos[3])
    out += struct.pack("<I", len(ft.contents))
    elem_pos = [len(out) + 4 * i for i in range(len(ft.contents))]
    out += b"\0" * (4 * len(ft.contents))
    for ep, e in zip(elem_pos, ft.contents):
        align(2)
        vt = len(out)
        out += struct.pack("<HHHHHH", 12, 24, 4, 12, 20, 22)
        # table must be 8-aligned at (t+4) for the int64 fields: place table so that t+4 is 8-aligned
        while (len(out) + 4) % 8:
            out.append(0)
        t = len(out)
        out += struct.pack("<i", t - vt)
        out += struct.pack("<qqhh", e.offset, e.length, e.format, e.content_type)
        struct.pack_into("<I", out, ep, t - ep)
    align(8)
    return bytes(out)


# ----------------------------------------------------------------------------------------------
# container
# ----------------------------------------------------------------------------------------------
@dataclass
class Pod5Tables:
    footer: Footer
    section_marker: bytes
    raw: Dict[int, bytes]  # content_type -> embedded Arrow IPC file bytes

    def table(self, content_type: int) -> pa.Table:
        return ipc.open_file(pa.BufferReader(self.raw[content_type])).read_all()

    @property
    def signal(self) -> pa.Table:
        return self.table(CONTENT_SIGNAL)

    @property
    def reads(self) -> pa.Table:
        return self.table(CONTENT_READS)

    @property
    def run_info(self) -> pa.Table:
        return self.table(CONTENT_RUN_INFO)


def read_pod5(path: str) -> Pod5Tables:
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 64 or data[:8] != SIGNATURE or data[-8:] != SIGNATURE:
        raise Pod5Error("Invalid signature in file")
    marker = data[8:24]
    end = len(data) - 8 - 16
    flen = _i64(data, end - 8)
    if flen < 0 or flen > end - 8:
        raise Pod5Error("Invalid footer length")
    fb = data[end - 8 - flen : end - 8]
    footer = parse_footer(fb)
    raw = {}
    for e in footer.contents:
        raw[e.content_type] = data[e.offset : e.offset + e.length]
    return Pod5Tables(footer=footer, section_marker=marker, raw=raw)


def signal_extension_name(tbl_or_schema) -> str:
    """Extension name of the `signal` column: minknow.vbz | pgnano.signal | '' (uncompressed)."""
    schema = tbl_or_schema.schema if hasattr(tbl_or_schema, "schema") else tbl_or_schema
    fld = schema.field("signal")
    md = fld.metadata or {}
    return (md.get(b"ARROW:extension:name") or b"").decode()


def table_to_ipc_bytes(table: pa.Table, max_rows_per_batch: int) -> bytes:
    sink = pa.BufferOutputStream()
    with ipc.new_file(sink, table.schema) as w:
        for b in table.to_batches(max_chunksize=max_rows_per_batch):
            w.write_batch(b)
    return sink.getvalue().to_pybytes()


def write_pod5(path: str, tables: Dict[int, bytes], file_identifier: str, software: str,
               pod5_version: str, section_marker: Optional[bytes] = None) -> None:
    """Write embedded Arrow IPC files in the reference's order (signal, run info, reads:
    combined_file_utils.h:85-127) wrapped in the combined-file container."""
    marker = section_marker or uuid.uuid4().bytes
    out = bytearray()
    out += SIGNATURE + marker
    contents = []
    for ct in (CONTENT_SIGNAL, CONTENT_RUN_INFO, CONTENT_READS):
        if ct not in tables:
            continue
        blob = tables[ct]
        contents.append(EmbeddedFile(offset=len(out), length=len(blob), format=0, content_type=ct))
        out += blob
        while len(out) % 8:
            out.append(0)
        out += marker
    out += FOOTER_MAGIC
    fb = build_footer(Footer(file_identifier, software, pod5_version, contents))
    out += fb
    while len(out) % 8:
        out.append(0)
    out += struct.pack("<q", len(fb))
    out += marker + SIGNATURE
    with open(path, "wb") as f:
        f.write(bytes(out))
