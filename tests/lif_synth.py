"""Writer of small synthetic Leica LIF files for the raw-reader tests (layout: ``lif_raw.py`` docstring)."""

from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

_DIM_IDS = {"X": 1, "Y": 2, "Z": 3, "T": 4, "M": 10}


def _image_xml(name: str, array: np.ndarray, axes: str, block_id: str) -> str:
    """axes names the array's axes, slowest first; 'C' marks the channel axis (planar storage)."""
    item = array.dtype.itemsize
    strides = {ax: s for ax, s in zip(axes, array.strides)}
    n_channels = array.shape[axes.index("C")] if "C" in axes else 1
    channel_step = strides.get("C", 0)
    channels = "".join(
        f'<ChannelDescription DataType="0" ChannelTag="0" Resolution="{8 * item}" LUTName="Gray" BytesInc="{c * channel_step}" BitInc="0"/>'
        for c in range(n_channels))
    dims = "".join(
        f'<DimensionDescription DimID="{_DIM_IDS[ax]}" NumberOfElements="{array.shape[axes.index(ax)]}" Origin="0" Length="1" '
        f'Unit="m" BitInc="0" BytesInc="{strides[ax]}"/>' for ax in axes if ax != "C")
    return (f'<Element Name="{name}" Visibility="1"><Data><Image><ImageDescription><Channels>{channels}</Channels>'
            f'<Dimensions>{dims}</Dimensions></ImageDescription></Image></Data>'
            f'<Memory Size="{array.nbytes}" MemoryBlockID="{block_id}"/><Children/></Element>')


def write_lif(path: Path, images: list[tuple[str, np.ndarray, str]], version: int = 2, folder: str | None = None) -> None:
    """images: (name, C-contiguous uint8 / uint16 array, axes).  ``folder``: nest every image below one folder element."""
    blocks = []
    elements = []
    for i, (name, array, axes) in enumerate(images):
        array = np.ascontiguousarray(array)
        block_id = f"MemBlock_{100 + i}"
        elements.append(_image_xml(name, array, axes, block_id))
        blocks.append((block_id, array.astype(array.dtype.newbyteorder("<"), copy=False).tobytes()))
    body = "".join(elements)
    if folder is not None:
        body = f'<Element Name="{folder}"><Children>{body}</Children></Element>'
    xml = (f'<LMSDataContainerHeader Version="{version}"><Element Name="experiment.lif" Visibility="1">'
           f'<Memory Size="0" MemoryBlockID="MemBlock_0"/><Children>{body}</Children></Element></LMSDataContainerHeader>')
    encoded = xml.encode("utf-16-le")
    out = bytearray()
    out += struct.pack("<iiBi", 0x70, len(encoded) + 5, 0x2A, len(xml)) + encoded
    for block_id, payload in blocks:
        id_bytes = block_id.encode("utf-16-le")
        size = struct.pack("<q", len(payload)) if version >= 2 else struct.pack("<i", len(payload))
        description = bytes([0x2A]) + size + bytes([0x2A]) + struct.pack("<i", len(block_id)) + id_bytes
        out += struct.pack("<ii", 0x70, len(description)) + description + payload
    Path(path).write_bytes(bytes(out))
