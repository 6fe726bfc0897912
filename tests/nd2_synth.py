"""Writes a minimal ND2 v3 container (the chunk layout documented in nd2_raw.py) around given
frames, so the raw-frame reader and the device de-interleave path can be tested without the
reference's fixture files (which do not travel to the GPU box)."""

from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

_MAGIC = 0x0ABECEDA


def _chunk(name: bytes, payload: bytes) -> bytes:
    return struct.pack("<IIQ", _MAGIC, len(name), len(payload)) + name + payload


def _lv_u32(key: str, value: int) -> bytes:
    name = key.encode("utf-16le") + b"\x00\x00"
    return bytes([2, len(key) + 1]) + name + struct.pack("<I", value)


def write_nd2(path: Path, frames: np.ndarray, row_align: int = 1) -> None:
    """frames: (n_frames, C, Y, X) uint16.  row_align > 1 pads every row to a multiple of that many bytes
    (real ND2 writers use 4), recorded in uiWidthBytes."""
    n, c, h, w = frames.shape
    pitch = -(-(w * c * 2) // row_align) * row_align
    blob = bytearray()
    table: list[tuple[bytes, int, int]] = []

    def add(name: bytes, payload: bytes) -> None:
        table.append((name, len(blob), len(payload)))
        blob.extend(_chunk(name, payload))

    add(b"ND2 FILE SIGNATURE CHUNK NAME01!", b"Ver3.0" + b"\x00" * 58)
    attrs = b"".join(_lv_u32(k, v) for k, v in [("uiWidth", w), ("uiWidthBytes", pitch), ("uiHeight", h), ("uiComp", c),
                                                ("uiBpcInMemory", 16), ("uiBpcSignificant", 16), ("uiSequenceCount", n)])
    add(b"ImageAttributesLV!", attrs)
    for i in range(n):
        yxc = np.ascontiguousarray(frames[i].transpose(1, 2, 0)).astype("<u2")
        rows = np.zeros((h, pitch), dtype=np.uint8)
        rows[:, : w * c * 2] = yxc.reshape(h, w * c).view(np.uint8)
        add(f"ImageDataSeq|{i}!".encode(), struct.pack("<d", 0.25 * i) + rows.tobytes())
    map_offset = len(blob)
    payload = b"".join(name + struct.pack("<QQ", off, ln) for name, off, ln in table)
    payload += b"ND2 CHUNK MAP SIGNATURE 0000001!" + struct.pack("<Q", map_offset)
    blob.extend(_chunk(b"ND2 FILEMAP SIGNATURE NAME 0001!", payload))
    blob.extend(struct.pack("<Q", map_offset))  # the reader takes the map offset from the last 8 bytes
    Path(path).write_bytes(bytes(blob))
