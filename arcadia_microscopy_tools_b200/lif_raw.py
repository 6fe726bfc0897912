"""Minimal host-side reader for the pixel blocks of Leica LIF files.

File decoding stays on the host (BASELINE.json north_star).  The reference decodes through the
third-party ``liffile`` package (``leica.py:52-80`` -> ``liffile.LifFile(...).images[name].asarray()``)
and also parses rich acquisition metadata; this reader only does what the B200 path needs: find the
named image, read its dimension / channel layout, and hand back the raw samples as an array whose axes
follow the strides stored in the file (e.g. ``(Z, C, Y, X)`` for a Stellaris z-stack), ready to be
copied into pinned staging.  Metadata parsing beyond sizes is out of scope.

Container layout (LIF versions 1 and 2), all little-endian:

* header block  ``[i32 0x70][i32 length][u8 0x2A][i32 n_chars][UTF-16LE XML, n_chars characters]``
* memory blocks ``[i32 0x70][i32 length][u8 0x2A][i32 (v1) or i64 (v2) memory_size][u8 0x2A]``
  ``[i32 n_chars][UTF-16LE block id, e.g. "MemBlock_221"][memory_size bytes of samples]``

The XML is a tree of ``<Element Name=...>``; an image element holds
``Data/Image/ImageDescription`` with ``Channels/ChannelDescription`` (``Resolution`` bits, ``BytesInc``
offset of the channel) and ``Dimensions/DimensionDescription`` (``DimID`` 1 = X, 2 = Y, 3 = Z, 4 = T, 10 = M
(mosaic tile), ``NumberOfElements``, ``BytesInc`` stride) and a ``Memory`` element naming its block.
The sample at (x, y, z, t, c) lives at ``sum(index * BytesInc) + channel BytesInc`` inside the block.

No LIF file ships with the reference, so this module is checked against files written by
``tests/lif_synth.py`` from the layout above (parity with ``liffile`` itself: unpinned).
"""

from __future__ import annotations

import struct
import xml.etree.ElementTree as ET
from dataclasses import dataclass
from pathlib import Path

import numpy as np

_MAGIC = 0x70
_TEST = 0x2A
_DIM_NAMES = {1: "X", 2: "Y", 3: "Z", 4: "T", 5: "L", 6: "R", 10: "M"}


@dataclass(frozen=True)
class LifImageInfo:
    """Layout of one image of a LIF file."""

    name: str
    block_id: str
    memory_size: int
    dtype: np.dtype
    sizes: dict[str, int]     # axis name -> length, in array order (slowest axis first); 'C' when > 1 channel
    strides: dict[str, int]   # axis name -> bytes between consecutive elements


def _expect(buf, offset: int, fmt: str, what: str):
    try:
        return struct.unpack_from(fmt, buf, offset)
    except struct.error as exc:
        raise ValueError(f"truncated LIF file while reading {what} at offset {offset}") from exc


def _read_header(buf) -> tuple[ET.Element, int]:
    """-> (XML root, offset of the first memory block)."""
    magic, _length, test, n_chars = _expect(buf, 0, "<iiBi", "the header block")
    if magic != _MAGIC or test != _TEST:
        raise ValueError(f"not a LIF file (magic {magic:#x}, marker {test:#x})")
    start = 13
    xml_bytes = bytes(buf[start : start + 2 * n_chars])
    if len(xml_bytes) != 2 * n_chars:
        raise ValueError("truncated LIF file: header XML is incomplete")
    return ET.fromstring(xml_bytes.decode("utf-16-le")), start + 2 * n_chars


def _memory_blocks(buf, offset: int, version: int) -> dict[str, tuple[int, int]]:
    """block id -> (data offset, size)."""
    blocks: dict[str, tuple[int, int]] = {}
    end = len(buf)
    while offset < end:
        magic, _length, test = _expect(buf, offset, "<iiB", "a memory block")
        if magic != _MAGIC or test != _TEST:
            raise ValueError(f"corrupt LIF file: bad memory block header at offset {offset}")
        offset += 9
        if version >= 2:
            (size,) = _expect(buf, offset, "<q", "a memory block size")
            offset += 8
        else:
            (size,) = _expect(buf, offset, "<i", "a memory block size")
            offset += 4
        test2, n_chars = _expect(buf, offset, "<Bi", "a memory block id")
        if test2 != _TEST:
            raise ValueError(f"corrupt LIF file: bad memory block marker at offset {offset}")
        offset += 5
        block_id = bytes(buf[offset : offset + 2 * n_chars]).decode("utf-16-le")
        offset += 2 * n_chars
        blocks[block_id] = (offset, size)
        offset += size
    return blocks


def _image_elements(root: ET.Element):
    """Every ``Element`` that describes an image, with its slash-joined path below the root element."""

    def walk(element: ET.Element, prefix: str):
        name = element.get("Name", "")
        path = f"{prefix}/{name}" if prefix else name
        if element.find("Data/Image/ImageDescription") is not None:
            yield path, element
        children = element.find("Children")
        if children is not None:
            for child in children.findall("Element"):
                yield from walk(child, path)

    top = root.find("Element")
    if top is None:
        return
    children = top.find("Children")
    for child in children.findall("Element") if children is not None else []:
        yield from walk(child, "")
    if top.find("Data/Image/ImageDescription") is not None:
        yield top.get("Name", ""), top


def _describe(path: str, element: ET.Element) -> LifImageInfo:
    description = element.find("Data/Image/ImageDescription")
    memory = element.find("Memory")
    if description is None or memory is None:
        raise ValueError(f"image '{path}' has no pixel block")
    channels = description.findall("Channels/ChannelDescription")
    if not channels:
        raise ValueError(f"image '{path}' lists no channels")
    bits = {int(c.get("Resolution", "8")) for c in channels}
    if len(bits) != 1:
        raise ValueError(f"image '{path}' mixes sample sizes {sorted(bits)}")
    resolution = bits.pop()
    if resolution <= 8:
        dtype = np.dtype(np.uint8)
    elif resolution <= 16:
        dtype = np.dtype("<u2")
    else:
        raise ValueError(f"image '{path}': {resolution}-bit samples are not supported")
    axes: list[tuple[int, str, int]] = []  # (stride, name, length)
    for dim in description.findall("Dimensions/DimensionDescription"):
        length = int(dim.get("NumberOfElements", "1"))
        dim_id = int(dim.get("DimID", "0"))
        if dim_id not in _DIM_NAMES:
            raise ValueError(f"image '{path}': unknown DimID {dim_id}")
        if length > 1 or dim_id in (1, 2):
            axes.append((int(dim.get("BytesInc", "0")), _DIM_NAMES[dim_id], length))
    if len(channels) > 1:
        offsets = [int(c.get("BytesInc", "0")) for c in channels]
        step = offsets[1] - offsets[0]
        if offsets[0] != 0 or step <= 0 or any(offsets[i] != i * step for i in range(len(offsets))):
            raise ValueError(f"image '{path}': irregular channel offsets {offsets}")
        axes.append((step, "C", len(channels)))
    # slowest axis first; at equal stride (length-1 axes) keep X last
    axes.sort(key=lambda a: (-a[0], a[1] == "X"))
    return LifImageInfo(
        name=path, block_id=memory.get("MemoryBlockID", ""), memory_size=int(memory.get("Size", "0")), dtype=dtype,
        sizes={name: length for _, name, length in axes}, strides={name: stride for stride, name, _ in axes},
    )


def _open(lif_path: Path):
    buf = np.memmap(lif_path, dtype=np.uint8, mode="r")
    root, offset = _read_header(buf)
    version = int(root.get("Version", "1"))
    images = {path: _describe(path, element) for path, element in _image_elements(root)}
    return buf, images, offset, version


def list_image_names(lif_path: Path) -> list[str]:
    """Names of the images in the file (ref: ``leica.py:39-49``); nested images are ``folder/name``."""
    _, images, _, _ = _open(Path(lif_path))
    return list(images)


def lif_image_info(lif_path: Path, image_name: str) -> LifImageInfo:
    _, images, _, _ = _open(Path(lif_path))
    if image_name not in images:
        raise ValueError(f"Image {image_name} not found in {lif_path}. Available images: {list(images)}")
    return images[image_name]


def read_lif_image(lif_path: Path, image_name: str) -> tuple[np.ndarray, dict[str, int]]:
    """(samples, sizes) of one image: a C-contiguous array whose axes are ``sizes``' keys in order
    (slowest stride first, e.g. ``Z, C, Y, X``), uint8 or uint16 as stored."""
    lif_path = Path(lif_path)
    buf, images, offset, version = _open(lif_path)
    if image_name not in images:
        raise ValueError(f"Image {image_name} not found in {lif_path}. Available images: {list(images)}")
    info = images[image_name]
    blocks = _memory_blocks(buf, offset, version)
    if info.block_id not in blocks:
        raise ValueError(f"memory block '{info.block_id}' of image '{image_name}' is missing from {lif_path}")
    start, size = blocks[info.block_id]
    if start + size > len(buf):
        raise ValueError(f"truncated LIF file: memory block '{info.block_id}' ends beyond the end of {lif_path}")
    shape = tuple(info.sizes.values())
    strides = tuple(info.strides[name] for name in info.sizes)
    last = sum((n - 1) * s for n, s in zip(shape, strides)) + info.dtype.itemsize
    if last > size:
        raise ValueError(f"image '{image_name}' needs {last} bytes but its memory block holds {size}")
    view = np.ndarray(shape, dtype=info.dtype, buffer=buf, offset=start, strides=strides)
    return np.ascontiguousarray(view), dict(info.sizes)
