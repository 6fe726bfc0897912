"""Minimal host-side reader for the raw frames of uncompressed Nikon ND2 files.

File decoding stays on the host (BASELINE.json north_star).  The reference decodes through
the third-party ``nd2`` package (``nikon.py:25-43`` -> ``nd2.ND2File.asarray``) and also parses
rich metadata; this reader only does what the B200 path needs: hand back the uint16 pixel
block in the ``(frames, C, Y, X)`` layout ``nd2.asarray()`` produces so it can be copied into
pinned staging.  Metadata parsing is out of scope.

Container layout (ND2 v3): a sequence of chunks
``[u32 magic 0x0ABECEDA][u32 name_len][u64 data_len][name (name_len bytes)][data]``; the last
8 bytes of the file hold the offset of the chunk-map chunk whose payload is a list of
``name!`` + ``u64 offset`` + ``u64 length`` records.  Frame ``i`` is chunk ``ImageDataSeq|i!``:
an 8-byte float64 timestamp followed by little-endian samples in (Y, X, C) interleaved order.
"""

from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

_MAGIC = 0x0ABECEDA


def _read_chunk(buf, offset: int):
    """Payload of the chunk at ``offset`` (a zero-copy slice of ``buf``: bytes or a memoryview).  Offsets and
    lengths come from the file: they are checked against its size before anything is sliced."""
    if not (0 <= offset <= len(buf) - 16):
        raise ValueError(f"ND2 chunk offset {offset} lies outside the file ({len(buf)} bytes)")
    magic, name_len, data_len = struct.unpack_from("<IIQ", buf, offset)
    if magic != _MAGIC:
        raise ValueError(f"not an ND2 chunk at offset {offset} (magic {magic:#x})")
    start = offset + 16 + name_len
    if start + data_len > len(buf):
        raise ValueError(f"ND2 chunk at offset {offset} claims {data_len} bytes beyond the end of the file")
    return buf[start : start + data_len]


def _chunk_map(buf) -> dict[bytes, tuple[int, int]]:
    if len(buf) < 40:
        raise ValueError("file too short to be an ND2 container")
    (map_offset,) = struct.unpack_from("<Q", buf, len(buf) - 8)
    payload = bytes(_read_chunk(buf, map_offset))
    out: dict[bytes, tuple[int, int]] = {}
    pos = 0
    while pos < len(payload):
        end = payload.find(b"!", pos)
        if end < 0:
            break
        name = payload[pos : end + 1]
        if name.startswith(b"ND2 CHUNK MAP SIGNATURE"):
            break
        if end + 17 > len(payload):
            raise ValueError("truncated ND2 chunk map")
        off, length = struct.unpack_from("<QQ", payload, end + 1)
        out[name] = (off, length)
        pos = end + 1 + 16
    return out


def _lite_variant_uint(payload, key: str) -> int:
    """Value of a 32-bit integer entry of a CLX-lite variant block: entries are
    ``[u8 type][u8 name_len][utf-16le name, NUL terminated][value]``."""
    payload = bytes(payload)
    needle = key.encode("utf-16le") + b"\x00\x00"
    pos = payload.find(needle)
    if pos < 0:
        raise KeyError(key)
    (val,) = struct.unpack_from("<I", payload, pos + len(needle))
    return int(val)


def _frame_geometry(attrs) -> tuple[int, int, int, int, int]:
    """(height, width, components, sequence count, row pitch in bytes).  ND2 writers pad rows to a multiple of four
    bytes (``uiWidthBytes``); files without the entry are tightly packed."""
    width = _lite_variant_uint(attrs, "uiWidth")
    height = _lite_variant_uint(attrs, "uiHeight")
    comps = _lite_variant_uint(attrs, "uiComp")
    bpc = _lite_variant_uint(attrs, "uiBpcInMemory")
    nseq = _lite_variant_uint(attrs, "uiSequenceCount")
    if bpc != 16:
        raise ValueError(f"only 16-bit ND2 frames are supported, got {bpc} bits")
    try:
        pitch = _lite_variant_uint(attrs, "uiWidthBytes")
    except KeyError:
        pitch = width * comps * 2
    if pitch < width * comps * 2 or pitch % 2:
        raise ValueError(f"ND2 row pitch {pitch} does not fit {width} x {comps} 16-bit samples")
    return height, width, comps, nseq, pitch


def read_nd2_frames(path: str | Path) -> np.ndarray:
    """Return the pixel data as ``(n_frames, C, Y, X)`` uint16, C-contiguous."""
    buf = Path(path).read_bytes()
    cmap = _chunk_map(buf)
    attrs = _read_chunk(buf, cmap[b"ImageAttributesLV!"][0])
    height, width, comps, nseq, pitch = _frame_geometry(attrs)
    frames = np.empty((nseq, comps, height, width), dtype=np.uint16)
    for i in range(nseq):
        off, _ = cmap[f"ImageDataSeq|{i}!".encode()]
        data = _read_chunk(buf, off)
        if len(data) < 8 + height * pitch:
            raise ValueError(f"ND2 frame {i} holds {len(data)} bytes, expected {8 + height * pitch}")
        rows = np.frombuffer(data, dtype="<u2", count=height * (pitch // 2), offset=8).reshape(height, pitch // 2)
        frames[i] = rows[:, : width * comps].reshape(height, width, comps).transpose(2, 0, 1)
    return frames


def nd2_frame_layout(path: str | Path) -> tuple[np.memmap, list[int], tuple[int, int, int]]:
    """Memory-map the file and locate the raw frames without touching the pixels.

    Returns ``(mmap, payload_offsets, (height, width, components))``: frame ``i``'s samples are the
    ``height*width*components`` little-endian uint16 values starting at byte ``payload_offsets[i]``,
    in (Y, X, C) order.  This is all the host does on the fast path: the payloads are memcpy'd
    into pinned staging and transposed on the device (``read_nd2_to_device``)."""
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    buf = memoryview(mm)  # struct.unpack_from and slicing without copying the file
    cmap = _chunk_map(buf)
    attrs = _read_chunk(buf, cmap[b"ImageAttributesLV!"][0])
    height, width, comps, nseq, pitch = _frame_geometry(attrs)
    if pitch != width * comps * 2:
        raise ValueError(f"ND2 rows are padded (pitch {pitch} bytes for {width * comps * 2}): use read_nd2_frames")
    offsets = []
    for i in range(nseq):
        off, _ = cmap[f"ImageDataSeq|{i}!".encode()]
        payload = _read_chunk(buf, off)  # bounds-checked
        if len(payload) < 8 + height * pitch:
            raise ValueError(f"ND2 frame {i} holds {len(payload)} bytes, expected {8 + height * pitch}")
        _, name_len, _ = struct.unpack_from("<IIQ", buf, off)
        offsets.append(off + 16 + name_len + 8)  # chunk header, name, float64 timestamp
    return mm, offsets, (height, width, comps)


def read_nd2_to_device(path: str | Path, device=None):
    """Fast path: raw frame payloads -> pinned staging (one memcpy per frame) -> one async H2D copy ->
    ``amt_deinterleave_u16``.  Returns a CUDA tensor ``(n_frames, C, Y, X)`` holding the uint16 bits
    (torch int16), identical to ``read_nd2_frames`` uploaded."""
    from . import _gpu, _lib

    torch = _gpu.torch_mod()
    dev = _gpu.require_cuda() if device is None else device
    try:
        mm, offsets, (height, width, comps) = nd2_frame_layout(path)
    except ValueError as err:
        if "padded" not in str(err):
            raise
        return _gpu.to_device(read_nd2_frames(path), dev)  # padded rows: the strided host reader, then one upload
    n_pix = height * width
    nbytes = n_pix * comps * 2
    staging = torch.empty((len(offsets), n_pix, comps), dtype=torch.int16, pin_memory=True)
    flat = staging.numpy().view(np.uint8).reshape(len(offsets), nbytes)
    for i, off in enumerate(offsets):
        flat[i] = mm[off : off + nbytes]
    raw = staging.to(dev, non_blocking=True)
    out = torch.empty((len(offsets), comps, height, width), dtype=torch.int16, device=dev)
    _lib.check(
        _lib.load().amt_deinterleave_u16(_gpu.ptr(raw), _gpu.ptr(out), len(offsets), n_pix, comps, _gpu.stream_ptr()),
        "amt_deinterleave_u16",
    )
    return out


def read_nd2(path: str | Path) -> np.ndarray:
    """Pixel block squeezed like ``nd2.ND2File.asarray()``: singleton frame/channel axes
    dropped, e.g. ``(C, Y, X)`` for a single multichannel frame, ``(T, Y, X)`` for a
    single-channel series."""
    frames = read_nd2_frames(path)
    return np.ascontiguousarray(np.squeeze(frames)) if frames.ndim > 2 else frames
