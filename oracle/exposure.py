"""Oracle: ``skimage.exposure.rescale_intensity``.  TEST INFRASTRUCTURE ONLY.

Called by the reference at ``operations.py:50-54`` with explicit ``in_range=(p1, p2)`` and
``out_range=(o1, o2)`` tuples.
"""

from __future__ import annotations

import numpy as np


def _supported_float_type(dtype: np.dtype) -> np.dtype:
    dtype = np.dtype(dtype)
    if dtype == np.float32 or dtype == np.float16:
        return np.dtype(np.float32)
    return np.dtype(np.float64)


def rescale_intensity(image: np.ndarray, in_range, out_range) -> np.ndarray:
    """Tuple ``out_range`` => output dtype is the supported float type of the input
    (float64 for uint16/float64).  ``imin/imax/omin/omax`` are Python floats (weak scalars
    under NEP 50), so ``np.clip(uint16_image, imin, imax)`` is computed in float64."""
    image = np.asarray(image)
    out_dtype = _supported_float_type(image.dtype)
    imin, imax = map(float, in_range)
    omin, omax = map(float, out_range)
    image = np.clip(image, imin, imax)
    if imin != imax:
        image = (image - imin) / (imax - imin)
        return (image * (omax - omin) + omin).astype(out_dtype)
    return np.clip(image, omin, omax).astype(out_dtype)
