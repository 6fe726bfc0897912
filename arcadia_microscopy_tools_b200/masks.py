"""Segmentation masks and per-cell quantification on the GPU.

Drop-in for the hot-path part of the reference's ``masks.py``: ``_process_mask`` (:38-65),
``SegmentationMask`` validation (:166-208), ``label_image`` / ``num_cells`` (:210-227),
``cell_properties`` (:247-328), ``centroids_yx`` (:330-353), ``filter`` (:355-418) and
``convert_properties_to_microns`` (:420-467) — same arguments, defaults, key order, dtypes and
error messages.  Labelling, border clearing, relabelling and every regionprops statistic are
CUDA kernels (``ccl.cu``, ``regions.cu``, ``shape.cu``); the reference's per-region Python loop
and its one-rescan-per-channel become a single streaming pass.  ``cell_outlines`` (:68-115,
:229-245): the pixel passes (marching-squares cases, OpenCV-style border following) are kernels of
``outlines.cu``; joining a cell's few dozen contour segments into an ordered polygon is host work.
"""

from __future__ import annotations

import warnings
from collections import deque
from collections.abc import Mapping
from functools import cached_property
from typing import ClassVar, Literal

import numpy as np

from . import _gpu, _lib
from .channels import Channel

DEFAULT_CELL_PROPERTY_NAMES = [
    "label",
    "centroid",
    "volume",
    "area",
    "area_convex",
    "perimeter",
    "eccentricity",
    "circularity",
    "solidity",
    "axis_major_length",
    "axis_minor_length",
    "orientation",
]

DEFAULT_INTENSITY_PROPERTY_NAMES = [
    "intensity_mean",
    "intensity_max",
    "intensity_min",
    "intensity_std",
]

# regionprops name -> table column(s) (include/amt_b200.h)
_MORPH_COLUMNS: dict[str, list[tuple[str, int]]] = {
    "label": [("label", _lib.COL_LABEL)],
    "area": [("area", _lib.COL_AREA)],
    "bbox": [(f"bbox-{i}", _lib.COL_BBOX0 + i) for i in range(4)],
    "centroid": [("centroid-0", _lib.COL_CENTROID0), ("centroid-1", _lib.COL_CENTROID0 + 1)],
    "inertia_tensor_eigvals": [("inertia_tensor_eigvals-0", _lib.COL_EIG0), ("inertia_tensor_eigvals-1", _lib.COL_EIG0 + 1)],
    "axis_major_length": [("axis_major_length", _lib.COL_AXIS_MAJOR)],
    "axis_minor_length": [("axis_minor_length", _lib.COL_AXIS_MINOR)],
    "eccentricity": [("eccentricity", _lib.COL_ECC)],
    "orientation": [("orientation", _lib.COL_ORIENT)],
    "perimeter": [("perimeter", _lib.COL_PERIMETER)],
    "area_convex": [("area_convex", _lib.COL_AREA_CONVEX)],
}
_SHAPE_PROPS = {"perimeter", "area_convex", "solidity"}
_INTENSITY_COLUMNS = {
    "intensity_sum": _lib.CH_SUM,  # extension: exact integer sum (north star), not a skimage name
    "intensity_mean": _lib.CH_MEAN,
    "intensity_max": _lib.CH_MAX,
    "intensity_min": _lib.CH_MIN,
    "intensity_std": _lib.CH_STD,
}
_INT_COLUMNS = {"label", "bbox-0", "bbox-1", "bbox-2", "bbox-3"}


def _label_on_device(mask_image: np.ndarray, remove_edge_cells: bool):
    """-> (labels int32 CUDA tensor (1, H, W), number of cells)."""
    mask = np.asarray(mask_image)
    if mask.dtype == np.bool_:
        labels, counts = _gpu.label(_gpu.to_device(mask)[None], 0, remove_edge_cells)
    else:
        max_value = int(mask.max())
        if max_value >= 2**31 - 1:
            raise MemoryError("label values beyond int32 are not supported on the B200 path")
        dev = _gpu.to_device(np.ascontiguousarray(mask, dtype=np.int32))[None]
        labels, counts = _gpu.label(dev, 2, remove_edge_cells, max_value=max_value)
    return labels, int(_gpu.to_host(counts)[0])


def _process_mask(mask_image: np.ndarray, remove_edge_cells: bool) -> np.ndarray:
    """Remove border-touching cells if asked, then hand back consecutive int64 labels
    (ref: ``masks.py:38-65``): bool masks are labelled by 8-connected components in raster
    order of first pixel; integer masks lose their border-touching connected fragments and are
    renumbered in ascending value order."""
    labels, count = _label_on_device(mask_image, remove_edge_cells)
    if remove_edge_cells and count == 0:
        raise ValueError("No cells remain after removing edge cells. Try setting remove_edge_cells=False.")
    return _gpu.to_host(labels[0]).astype(np.int64)



# ---- skimage.measure.find_contours, host part: segments of the marching-squares cases and their
# assembly into ordered contours (skimage ``_find_contours_cy._get_contour_segments`` with
# fully_connected='low', positive_orientation='low', and ``_find_contours._assemble_contours``) [3p].
# Points are kept as integer pairs (2*row, 2*col): every contour point of a binary image at level 0.5
# is an edge midpoint, so the doubled coordinates are exact.
def _case_segments(r2: int, c2: int, case: int):
    """(from, to) point pairs of one square whose upper-left pixel is (r2/2, c2/2)."""
    top, bottom = (r2, c2 + 1), (r2 + 2, c2 + 1)
    left, right = (r2 + 1, c2), (r2 + 1, c2 + 2)
    table = {
        1: ((top, left),), 2: ((right, top),), 3: ((right, left),), 4: ((left, bottom),),
        5: ((top, bottom),), 6: ((right, top), (left, bottom)), 7: ((right, bottom),),
        8: ((bottom, right),), 9: ((top, left), (bottom, right)), 10: ((bottom, top),),
        11: ((bottom, left),), 12: ((left, right),), 13: ((top, right),), 14: ((left, top),),
    }
    return table[case]


def _join_segments(segments) -> list[list[tuple[int, int]]]:
    """Chains directed segments head to tail in the order given.  When a segment links two chains the
    one created first survives (the other is appended / prepended to it), which fixes both the order
    of the returned contours and the point each closed contour starts from."""
    chains: dict[int, deque] = {}
    by_start: dict[tuple[int, int], int] = {}
    by_end: dict[tuple[int, int], int] = {}
    serial = 0
    for src, dst in segments:
        after = by_start.pop(dst, None)   # chain that begins where this segment ends
        before = by_end.pop(src, None)    # chain that ends where this segment begins
        if after is None and before is None:
            chains[serial] = deque((src, dst))
            by_start[src] = serial
            by_end[dst] = serial
            serial += 1
        elif before is None:
            chains[after].appendleft(src)
            by_start[src] = after
        elif after is None:
            chains[before].append(dst)
            by_end[dst] = before
        elif after == before:  # the chain closes on itself
            chains[before].append(dst)
        elif after > before:  # the later chain is appended to the earlier one
            tail = chains.pop(after)
            chains[before].extend(tail)
            by_end[chains[before][-1]] = before
        else:  # the later chain (ending at src) goes in front of the earlier one
            head = chains.pop(before)
            by_start.pop(head[0], None)
            chains[after].extendleft(reversed(head))
            by_start[chains[after][0]] = after
    return [list(chains[k]) for k in sorted(chains)]


def _contours_from_square_keys(keys: np.ndarray, num_cells: int) -> list[np.ndarray]:
    """keys: sorted ``label<<34 | r0<<19 | c0<<4 | case`` (``amt_outline_squares``) -> per label the
    longest contour (first of equals) as float64 (y, x) points in image coordinates."""
    labels = (keys >> np.uint64(34)).astype(np.int64)
    rows2 = (2 * ((keys >> np.uint64(19)) & np.uint64(0x7FFF))).astype(np.int64).tolist()
    cols2 = (2 * ((keys >> np.uint64(4)) & np.uint64(0x7FFF))).astype(np.int64).tolist()
    cases = (keys & np.uint64(15)).astype(np.int64).tolist()
    bounds = np.searchsorted(labels, np.arange(1, num_cells + 2))
    outlines = []
    for k in range(num_cells):
        lo, hi = int(bounds[k]), int(bounds[k + 1])
        segments = [seg for i in range(lo, hi) for seg in _case_segments(rows2[i], cols2[i], cases[i])]
        contours = _join_segments(segments)
        if contours:
            outlines.append(np.array(max(contours, key=len), dtype=np.float64) / 2.0)
        else:
            outlines.append(np.array([]).reshape(0, 2))
    return outlines



def _outlines_cellpose_device(labels, num_cells: int) -> list[np.ndarray]:
    points, offsets = _gpu.outline_borders(labels, num_cells)
    points = points.astype(np.int64)
    return [points[offsets[k] : offsets[k + 1]] if offsets[k + 1] > offsets[k] else np.zeros((0, 2)) for k in range(num_cells)]


def _outlines_skimage_device(labels, num_cells: int) -> list[np.ndarray]:
    return _contours_from_square_keys(_gpu.outline_square_keys(labels), num_cells)


def _labels_for_outlines(label_image: np.ndarray):
    """A host label image as the reference's outline helpers take it (labels 1..K, every one present)."""
    label_image = np.asarray(label_image)
    if label_image.ndim != 2:
        raise ValueError("label_image must be a 2D array")
    if label_image.size and (label_image.min() < 0 or label_image.max() >= 2**30):
        raise ValueError("labels must lie in [0, 2**30)")
    return _gpu.to_device(np.ascontiguousarray(label_image, dtype=np.int32)), int(label_image.max()) if label_image.size else 0


def _extract_outlines_cellpose(label_image: np.ndarray) -> list[np.ndarray]:
    """Same name, argument and result as the reference's helper (ref: ``masks.py:68-79``): one (y, x) outline per
    label 1..K from OpenCV-order border following on the GPU."""
    labels, num_cells = _labels_for_outlines(label_image)
    return _outlines_cellpose_device(labels, num_cells) if num_cells else []


def _extract_outlines_skimage(label_image: np.ndarray) -> list[np.ndarray]:
    """Same name, argument and result as the reference's helper (ref: ``masks.py:82-115``; its tests call it
    directly, ``tests/test_masks.py:86-149``): marching-squares cases on the GPU, contours joined on the host."""
    labels, num_cells = _labels_for_outlines(label_image)
    return _outlines_skimage_device(labels, num_cells) if num_cells else []


class SegmentationMask:
    """A label (or boolean) mask with optional per-channel intensity images.

    Args mirror the reference dataclass (ref: ``masks.py:118-143``): ``mask_image``,
    ``intensity_image_dict``, ``remove_edge_cells=True``, ``outline_extractor="cellpose"``,
    ``property_names=None``, ``intensity_property_names=None``.
    """

    _IMMUTABLE_FIELDS: ClassVar[frozenset[str]] = frozenset(
        {
            "mask_image",
            "intensity_image_dict",
            "remove_edge_cells",
            "outline_extractor",
            "property_names",
            "intensity_property_names",
        }
    )

    def __init__(
        self,
        mask_image: np.ndarray,
        intensity_image_dict: Mapping[Channel, np.ndarray] | None = None,
        remove_edge_cells: bool = True,
        outline_extractor: Literal["cellpose", "skimage"] = "cellpose",
        property_names: list[str] | None = None,
        intensity_property_names: list[str] | None = None,
    ) -> None:
        self.mask_image = mask_image
        self.intensity_image_dict = intensity_image_dict
        self.remove_edge_cells = remove_edge_cells
        self.outline_extractor = outline_extractor
        self.property_names = property_names
        self.intensity_property_names = intensity_property_names
        self._validate()
        object.__setattr__(self, "_initialized", True)

    def __setattr__(self, name: str, value: object) -> None:
        if getattr(self, "_initialized", False) and name in self._IMMUTABLE_FIELDS:
            raise AttributeError(
                f"Cannot modify '{name}' after SegmentationMask is initialized. Create a new instance instead."
            )
        super().__setattr__(name, value)

    def __repr__(self) -> str:
        shape = getattr(self.mask_image, "shape", None)
        return f"SegmentationMask(shape={shape}, remove_edge_cells={self.remove_edge_cells})"

    def _validate(self) -> None:
        if not isinstance(self.mask_image, np.ndarray):
            raise TypeError("mask_image must be a numpy array")
        if self.mask_image.ndim != 2:
            raise ValueError("mask_image must be a 2D array")
        if np.any(self.mask_image < 0):
            raise ValueError("mask_image must have non-negative values")
        if self.mask_image.max() == 0:
            raise ValueError("mask_image contains no cells (all values are 0)")
        if self.intensity_image_dict is not None:
            if not isinstance(self.intensity_image_dict, Mapping):
                raise TypeError("intensity_image_dict must be a Mapping of channels to 2D arrays")
            for channel, intensities in self.intensity_image_dict.items():
                if not isinstance(intensities, np.ndarray):
                    raise TypeError(f"Intensity image for '{channel.name}' must be a numpy array")
                if intensities.ndim != 2:
                    raise ValueError(f"Intensity image for '{channel.name}' must be 2D")
                if intensities.shape != self.mask_image.shape:
                    raise ValueError(f"Intensity image for '{channel.name}' must have same shape as mask_image")
            # shallow copy: key changes in one instance must not leak into another
            self.intensity_image_dict = dict(self.intensity_image_dict)
        if self.property_names is None:
            self.property_names = DEFAULT_CELL_PROPERTY_NAMES.copy()
        if self.intensity_property_names is None:
            self.intensity_property_names = (
                DEFAULT_INTENSITY_PROPERTY_NAMES.copy() if self.intensity_image_dict else []
            )

    # ------------------------------------------------------------------ labels
    @cached_property
    def _labels_device(self):
        labels, count = _label_on_device(self.mask_image, self.remove_edge_cells)
        if self.remove_edge_cells and count == 0:
            raise ValueError("No cells remain after removing edge cells. Try setting remove_edge_cells=False.")
        return labels, count

    @cached_property
    def label_image(self) -> np.ndarray:
        """Consecutive int64 labels starting at 1, background 0 (ref: ``masks.py:210-218``)."""
        return _gpu.to_host(self._labels_device[0][0]).astype(np.int64)

    @cached_property
    def num_cells(self) -> int:
        return int(self._labels_device[1])

    @cached_property
    def cell_outlines(self) -> list[np.ndarray]:
        """One (y, x) outline per cell, index i = label i + 1, an empty (0, 2) array where no contour
        exists (ref: ``masks.py:229-245``).  ``outline_extractor="cellpose"``: OpenCV's outer border of
        the label (``cellpose.utils.outlines_list``: longest contour, more than 4 points), traced on the
        GPU; ``"skimage"``: ``find_contours`` at level 0.5 on the 1-px padded crop (``masks.py:82-115``),
        marching-squares cases from the GPU, joined into ordered contours here."""
        labels = self._labels_device[0][0]
        if self.outline_extractor == "cellpose":
            return _outlines_cellpose_device(labels, self.num_cells)
        return _outlines_skimage_device(labels, self.num_cells)

    # ------------------------------------------------------------------ properties
    @cached_property
    def cell_properties(self) -> dict[str, np.ndarray]:
        """Per-cell morphology and per-channel intensity statistics (ref: ``masks.py:247-328``).

        Keys, key order and dtypes are those of the reference's two ``regionprops_table`` calls
        plus the derived ``circularity`` / ``volume`` and the ``centroid_y/x`` rename;
        ``intensity_sum`` is accepted as an extra intensity property (exact integer sum).
        """
        assert self.property_names is not None and self.intensity_property_names is not None
        needs_circularity = "circularity" in self.property_names
        needs_volume = "volume" in self.property_names
        region_props = [p for p in self.property_names if p not in ("circularity", "volume")]
        added: set[str] = set()
        for dep in (["area", "perimeter"] if needs_circularity else []) + (
            ["axis_major_length", "axis_minor_length"] if needs_volume else []
        ):
            if dep not in region_props:
                region_props.append(dep)
                added.add(dep)
        for p in region_props:
            if p not in _MORPH_COLUMNS and p != "solidity":
                raise NotImplementedError(f"property '{p}' is not available on the B200 path")
        for p in self.intensity_property_names:
            if p not in _INTENSITY_COLUMNS:
                raise NotImplementedError(f"intensity property '{p}' is not available on the B200 path")

        labels, count = self._labels_device
        channels = list(self.intensity_image_dict.items()) if self.intensity_image_dict else []
        use_channels = bool(channels and self.intensity_property_names)
        stack = None
        if use_channels:
            imgs = []
            for channel, img in channels:
                if img.dtype not in (np.uint16, np.uint8):
                    raise TypeError(
                        f"Intensity image for '{channel.name}' must be uint16 on the B200 path, got {img.dtype}"
                    )
                imgs.append(np.ascontiguousarray(img, dtype=np.uint16))
            stack = _gpu.to_device(np.stack(imgs))[None]
        with_shape = any(p in _SHAPE_PROPS for p in region_props)
        counts = _gpu.torch_mod().tensor([count], dtype=_gpu.torch_mod().int32, device=labels.device)
        table, _ = _gpu.region_table(labels, counts, stack, max(count, 1), with_shape=with_shape)
        tab = _gpu.to_host(table[0])[:, :count]

        def column(name: str, idx: int) -> np.ndarray:
            col = np.ascontiguousarray(tab[idx])
            return col.astype(np.int64) if name in _INT_COLUMNS else col

        properties: dict[str, np.ndarray] = {}
        for p in region_props:
            if p == "solidity":
                properties[p] = tab[_lib.COL_AREA] / tab[_lib.COL_AREA_CONVEX]
                continue
            for key, idx in _MORPH_COLUMNS[p]:
                properties[key] = column(key, idx)
        if needs_circularity:
            area, perimeter = properties["area"], properties["perimeter"]
            with np.errstate(divide="ignore", invalid="ignore"):
                properties["circularity"] = np.where(perimeter > 0, (4.0 * np.pi * area) / (perimeter**2), 0.0)
        if needs_volume:
            a = properties["axis_major_length"] / 2.0
            b = properties["axis_minor_length"] / 2.0
            properties["volume"] = np.where((a > 0) & (b > 0), (4.0 / 3.0) * np.pi * a * b * b, 0.0)
        for p in added:
            properties.pop(p, None)
        if "centroid-0" in properties:
            properties["centroid_y"] = properties.pop("centroid-0")
        if "centroid-1" in properties:
            properties["centroid_x"] = properties.pop("centroid-1")
        if use_channels:
            for c, (channel, _) in enumerate(channels):
                base = _lib.AMT_TABLE_BASE + c * _lib.AMT_TABLE_PER_CHANNEL
                for p in self.intensity_property_names:
                    col = np.ascontiguousarray(tab[base + _INTENSITY_COLUMNS[p]])
                    if p == "intensity_sum":
                        col = col.astype(np.uint64)
                    properties[f"{p}_{channel.name.lower()}"] = col
        return properties

    @cached_property
    def centroids_yx(self) -> np.ndarray:
        """(num_cells, 2) array of [y, x] centroids (ref: ``masks.py:330-353``)."""
        if self.property_names is None:
            raise ValueError("property_names cannot be None.")
        if "centroid" not in self.property_names:
            warnings.warn(
                "Centroid property not available. Include 'centroid' in property_names "
                "to get centroid coordinates. Returning empty array.",
                UserWarning,
                stacklevel=2,
            )
            return np.array([]).reshape(0, 2)
        return np.array([self.cell_properties["centroid_y"], self.cell_properties["centroid_x"]], dtype=float).T

    # ------------------------------------------------------------------ filter / units
    def filter(self, property_name: str, min_value: float | None = None, max_value: float | None = None) -> "SegmentationMask":
        """New mask keeping the cells whose property lies in the inclusive range
        (ref: ``masks.py:355-418``)."""
        assert self.property_names is not None and self.intensity_property_names is not None
        if min_value is None and max_value is None:
            raise ValueError("At least one of min_value or max_value must be provided.")
        if property_name not in self.cell_properties:
            raise ValueError(
                f"Property '{property_name}' not found. Available properties: {list(self.cell_properties.keys())}"
            )
        values = self.cell_properties[property_name]
        keep = np.ones(self.num_cells, dtype=bool)
        if min_value is not None:
            keep &= values >= min_value
        if max_value is not None:
            keep &= values <= max_value
        lut = np.zeros(self.num_cells + 1, dtype=np.int64)  # labels are consecutive 1..num_cells
        lut[1:][keep] = np.arange(1, self.num_cells + 1)[keep]
        new_label_image = lut[self.label_image]
        if new_label_image.max() == 0:
            raise ValueError(
                f"No cells remain after filtering '{property_name}' with min={min_value}, max={max_value}."
            )
        return SegmentationMask(
            mask_image=new_label_image,
            intensity_image_dict=self.intensity_image_dict,
            remove_edge_cells=False,
            outline_extractor=self.outline_extractor,
            property_names=list(self.property_names),
            intensity_property_names=list(self.intensity_property_names),
        )

    def convert_properties_to_microns(self, pixel_size_um: float) -> dict[str, np.ndarray]:
        """Scale lengths / areas / volumes to microns and suffix their keys with
        ``_um`` / ``_um2`` / ``_um3`` (ref: ``masks.py:420-467``)."""
        power = {"perimeter": 1, "axis_major_length": 1, "axis_minor_length": 1, "area": 2, "area_convex": 2,
                 "volume": 3, "inertia_tensor": 2, "inertia_tensor_eigvals": 2}
        suffix = {1: "_um", 2: "_um2", 3: "_um3"}
        converted: dict[str, np.ndarray] = {}
        for name, values in self.cell_properties.items():
            k = power.get(name)
            if k is None:
                converted[name] = values
            else:
                converted[f"{name}{suffix[k]}"] = values * (pixel_size_um**k)
        return converted
