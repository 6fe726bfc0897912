"""Oracle: the reference's own call sequences on top of the restated third-party functions.
TEST INFRASTRUCTURE ONLY.

Each function mirrors one reference function (same argument meaning, same exceptions):
``operations.py:10-54, 57-97, 135-216`` and ``masks.py:38-65, 247-328``.
"""

from __future__ import annotations

import numpy as np

from . import exposure, filters, labeling, regionprops, threshold

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
DEFAULT_INTENSITY_PROPERTY_NAMES = ["intensity_mean", "intensity_max", "intensity_min", "intensity_std"]


def rescale_by_percentile(intensities, percentile_range=(0, 100), out_range=(0, 1)):
    """ref: operations.py:10-54."""
    if not (0 <= percentile_range[0] < percentile_range[1] <= 100):
        raise ValueError(
            f"Invalid percentile range: {percentile_range}. "
            f"Values must be in ascending order between 0 and 100."
        )
    if intensities.size == 0:
        return np.zeros_like(intensities, dtype=float)
    if intensities.min() == intensities.max():
        return np.full_like(intensities, out_range[0], dtype=float)
    p1, p2 = np.percentile(intensities, percentile_range)
    return exposure.rescale_intensity(intensities, in_range=(p1, p2), out_range=out_range)


def subtract_background_dog(intensities, low_sigma=0.6, high_sigma=16.0, percentile=0):
    """ref: operations.py:57-97."""
    if not (0 <= percentile <= 100):
        raise ValueError(f"Percentile must be between 0 and 100, got {percentile}")
    if low_sigma >= high_sigma:
        raise ValueError(f"low_sigma ({low_sigma}) must be smaller than high_sigma ({high_sigma})")
    dog = filters.difference_of_gaussians(intensities, low_sigma, high_sigma)
    background_level = np.percentile(dog, percentile)
    return np.clip(dog - background_level, 0, None)


def crop_to_center(intensities, output_shape):
    """ref: operations.py:100-132: centred slice of the last two axes, clamped to the image size (a view)."""
    height, width = intensities.shape[-2:]
    crop_height = min(height, output_shape[0])
    crop_width = min(width, output_shape[1])
    top = (height - crop_height) // 2
    left = (width - crop_width) // 2
    return intensities[..., top : top + crop_height, left : left + crop_width]


def apply_threshold(intensities, method="otsu", **kwargs):
    """ref: operations.py:135-216 (all ten methods)."""
    if intensities.size == 0:
        return np.zeros_like(intensities, dtype=bool)
    if intensities.min() == intensities.max():
        return np.zeros_like(intensities, dtype=bool)
    funcs = {"otsu": threshold.threshold_otsu, "isodata": threshold.threshold_isodata,
             "yen": threshold.threshold_yen, "mean": threshold.threshold_mean, "li": threshold.threshold_li,
             "minimum": threshold.threshold_minimum, "triangle": threshold.threshold_triangle,
             "local": threshold.threshold_local, "niblack": threshold.threshold_niblack,
             "sauvola": threshold.threshold_sauvola}
    if method.lower() not in funcs:
        raise ValueError(f"Unsupported thresholding method: '{method}'.")
    return intensities > funcs[method.lower()](intensities, **kwargs)


def process_mask(mask_image, remove_edge_cells):
    """ref: masks.py:38-65 (``_process_mask``)."""
    label_image = mask_image
    if remove_edge_cells:
        label_image = labeling.clear_border(label_image)
        if label_image.max() == 0:
            raise ValueError(
                "No cells remain after removing edge cells. Try setting remove_edge_cells=False."
            )
    if label_image.dtype == bool:
        return labeling.label(label_image).astype(np.int64)
    return labeling.relabel_sequential(label_image).astype(np.int64)


def cell_properties(label_image, intensity_image_dict=None, property_names=None, intensity_property_names=None):
    """ref: masks.py:247-328 (``SegmentationMask.cell_properties``).  ``intensity_image_dict``
    maps a lower-cased channel-name suffix (or an object with ``.name``) to a 2-D image."""
    if property_names is None:
        property_names = list(DEFAULT_CELL_PROPERTY_NAMES)
    if intensity_property_names is None:
        intensity_property_names = list(DEFAULT_INTENSITY_PROPERTY_NAMES) if intensity_image_dict else []
    needs_circularity = "circularity" in property_names
    needs_volume = "volume" in property_names
    skimage_props = [p for p in property_names if p not in ("circularity", "volume")]
    added_props: set[str] = set()
    for dep in ["area", "perimeter"] if needs_circularity else []:
        if dep not in skimage_props:
            skimage_props.append(dep)
            added_props.add(dep)
    for dep in ["axis_major_length", "axis_minor_length"] if needs_volume else []:
        if dep not in skimage_props:
            skimage_props.append(dep)
            added_props.add(dep)
    properties = regionprops.regionprops_table(label_image, properties=skimage_props)
    if needs_circularity:
        area = properties["area"]
        perimeter = properties["perimeter"]
        with np.errstate(divide="ignore", invalid="ignore"):
            properties["circularity"] = np.where(perimeter > 0, (4.0 * np.pi * area) / (perimeter**2), 0.0)
    if needs_volume:
        a = properties["axis_major_length"] / 2.0
        b = properties["axis_minor_length"] / 2.0
        properties["volume"] = np.where((a > 0) & (b > 0), (4.0 / 3.0) * np.pi * a * b * b, 0.0)
    for prop in added_props:
        properties.pop(prop, None)
    if "centroid-0" in properties:
        properties["centroid_y"] = properties.pop("centroid-0")
    if "centroid-1" in properties:
        properties["centroid_x"] = properties.pop("centroid-1")
    if intensity_image_dict and intensity_property_names:
        for channel, intensities in intensity_image_dict.items():
            name = channel if isinstance(channel, str) else channel.name
            channel_props = regionprops.regionprops_table(
                label_image, intensity_image=intensities, properties=intensity_property_names
            )
            for prop_name, prop_values in channel_props.items():
                properties[f"{prop_name}_{name.lower()}"] = prop_values
    return properties
