"""3-D per-object quantification of label volumes (confocal z-stacks).

Extension beyond the reference, whose ``SegmentationMask`` rejects non-2-D masks
(``masks.py:171-172``): BASELINE.json's z-stack configuration asks for "per-slice filters + 3-D
per-label quantification".  The definitions are ``skimage.measure.regionprops_table``'s on a 3-D
label image; the arithmetic runs in ``csrc/regions3d.cu`` (one streaming
pass, exact integer accumulators).  Labels are made consecutive first with the same
``relabel_sequential`` kernel path ``SegmentationMask`` uses for integer masks.
"""

from __future__ import annotations

import numpy as np

from . import _gpu, _lib
from ._lib import check

VOLUME_COLUMNS = ["label", "area", "bbox-0", "bbox-1", "bbox-2", "bbox-3", "bbox-4", "bbox-5",
                  "centroid_z", "centroid_y", "centroid_x",
                  "inertia_tensor_eigvals-0", "inertia_tensor_eigvals-1", "inertia_tensor_eigvals-2",
                  "axis_major_length", "axis_minor_length"]
_PER_CHANNEL = ["intensity_sum", "intensity_mean", "intensity_max", "intensity_min", "intensity_std"]


def relabel_volume(label_volume, max_value: int | None = None):
    """``skimage.segmentation.relabel_sequential(label_volume)[0]`` for a (Z, Y, X) integer volume
    (host array or CUDA int32 tensor) -> (CUDA int32 volume with labels 1..K, K)."""
    torch = _gpu.torch_mod()
    if _gpu.is_device_array(label_volume):
        vol = label_volume
    else:
        a = np.asarray(label_volume)
        if a.ndim != 3:
            raise ValueError(f"label_volume must be 3-D, got {a.ndim}-D")
        if a.dtype.kind not in "iu":
            raise TypeError(f"label_volume must be an integer array, got {a.dtype}")
        if a.size and a.min() < 0:
            raise ValueError("label_volume must be non-negative")
        vol = _gpu.to_device(a.astype(np.int32))
    d, h, w = vol.shape
    if max_value is None:
        max_value = int(vol.max().item()) if vol.numel() else 0
    labels, counts = _gpu.label(vol.reshape(1, d * h, w), 2, False, None, max_value)
    return labels.reshape(d, h, w), int(counts.item())


def quantify_label_volume(label_volume, intensity_volume_dict: dict | None = None) -> dict[str, np.ndarray]:
    """Per-object table of a 3-D label volume: ``label, area, bbox-0..5, centroid_z/y/x,
    inertia_tensor_eigvals-0..2, axis_major_length, axis_minor_length`` and, per channel ``name``
    (lower-cased), ``intensity_sum/mean/max/min/std_<name>``; one row per label in ascending order.

    ``intensity_volume_dict`` maps channel names (or ``Channel`` objects) to uint16 (Z, Y, X)
    volumes.  Integer columns (label, bbox, intensity_sum) are exact.
    """
    torch = _gpu.torch_mod()
    lib = _lib.load()
    labels, k = relabel_volume(label_volume)
    d, h, w = labels.shape
    chans = intensity_volume_dict or {}
    names = [getattr(c, "name", c).lower() for c in chans]
    n_ch = len(names)
    if n_ch > 8:
        raise ValueError("at most 8 intensity channels")
    vols = []
    for c, v in chans.items():
        if _gpu.is_device_array(v):
            t = v
        else:
            v = np.asarray(v)
            if v.shape != (d, h, w):
                raise ValueError(f"intensity volume for {c} has shape {v.shape}, expected {(d, h, w)}")
            if v.dtype != np.uint16:
                raise TypeError(f"intensity volume for {c} must be uint16, got {v.dtype}")
            t = _gpu.to_device(v)
        vols.append(t.reshape(-1))
    channels = torch.stack(vols) if vols else None
    max_labels = max(k, 1)
    acc = torch.empty((_lib.acc3d_fields(n_ch), max_labels), dtype=torch.int64, device=labels.device)
    table = torch.empty((_lib.table3d_cols(n_ch), max_labels), dtype=torch.float64, device=labels.device)
    check(lib.amt_region_reduce3d(_gpu.ptr(labels), _gpu.ptr(channels), n_ch, d * h * w, d, h, w, max_labels,
                                  _gpu.ptr(acc), _gpu.stream_ptr()), "amt_region_reduce3d")
    check(lib.amt_region_finalize3d(_gpu.ptr(acc), k, n_ch, max_labels, _gpu.ptr(table), _gpu.stream_ptr()),
          "amt_region_finalize3d")
    host = _gpu.to_host(table)[:, :k]
    out: dict[str, np.ndarray] = {}
    cols = list(VOLUME_COLUMNS)
    for n in names:
        cols += [f"{p}_{n}" for p in _PER_CHANNEL]
    for i, name in enumerate(cols):
        col = np.ascontiguousarray(host[i])
        if name == "label" or name.startswith("bbox-"):
            col = col.astype(np.int64)
        elif name.startswith("intensity_sum_"):
            col = col.astype(np.uint64)
        out[name] = col
    return out
