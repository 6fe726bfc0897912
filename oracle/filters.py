"""Oracle: img_as_float, Gaussian, difference of Gaussians.  TEST INFRASTRUCTURE ONLY.

Restates ``skimage.filters.difference_of_gaussians`` as called by the reference at
``src/arcadia_microscopy_tools/operations.py:91`` (skimage 0.25.2 -> ``skimage.filters.gaussian``
-> ``scipy.ndimage.gaussian_filter(mode='nearest', truncate=4.0)``), SURVEY.md 8a items 1-2.
"""

from __future__ import annotations

import numpy as np
from scipy import ndimage as ndi

_UINT_MAX = {np.dtype(np.uint8): 255.0, np.dtype(np.uint16): 65535.0, np.dtype(np.uint32): 4294967295.0}


def img_as_float(image: np.ndarray) -> np.ndarray:
    """``skimage.util.img_as_float`` for the dtypes on the hot path.

    Unsigned ints are multiplied by the reciprocal of the dtype maximum (``np.multiply(image,
    1.0 / imax, dtype=float64)``), not divided; float64 passes through unscaled; bool -> 0/1.
    """
    image = np.asarray(image)
    if image.dtype == np.float64:
        return image
    if image.dtype == np.float32 or image.dtype == np.float16:
        # skimage keeps float32 (float16 -> float32); not on the B200 path.
        return image.astype(np.float32, copy=False)
    if image.dtype == np.bool_:
        return image.astype(np.float64)
    if image.dtype in _UINT_MAX:
        return np.multiply(image, 1.0 / _UINT_MAX[image.dtype], dtype=np.float64)
    if image.dtype.kind == "u":  # skimage.util.dtype._convert, unsigned -> float: multiply by 1 / imax
        return np.multiply(image, 1.0 / np.iinfo(image.dtype).max, dtype=np.float64)
    if image.dtype.kind == "i":  # signed -> float: (x + 0.5) * 2 / (imax - imin), in float64, in this order
        info = np.iinfo(image.dtype)
        out = np.add(image, 0.5, dtype=np.float64)
        out *= 2
        out /= int(info.max) - int(info.min)
        return out
    raise TypeError(f"oracle.img_as_float: dtype {image.dtype} is outside the hot path")


def gaussian_kernel1d(sigma: float, truncate: float = 4.0) -> np.ndarray:
    """Weights exactly as ``scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius)``."""
    sd = float(sigma)
    radius = int(truncate * sd + 0.5)
    sigma2 = sd * sd
    x = np.arange(-radius, radius + 1)
    phi_x = np.exp(-0.5 / sigma2 * x**2)
    phi_x = phi_x / phi_x.sum()
    return phi_x


def correlate1d_symmetric_nearest(x: np.ndarray, weights: np.ndarray, axis: int) -> np.ndarray:
    """Restatement of scipy's ``correlate1d`` symmetric-kernel loop, edge-clamped.

    Per output sample: ``acc = x[0]*w[c]`` then for ``j = r .. 1``:
    ``acc += (x[-j] + x[+j]) * w[c-j]`` in float64 with separately rounded add/mul (numpy
    elementwise ops never contract to FMA).  Bit-identical to
    ``scipy.ndimage.correlate1d(x, weights, axis, mode='nearest')`` for symmetric weights
    (checked in tests/test_oracle.py).
    """
    x = np.asarray(x, dtype=np.float64)
    r = (len(weights) - 1) // 2
    n = x.shape[axis]
    xm = np.moveaxis(x, axis, 0)
    idx = np.clip(np.arange(-r, n + r), 0, n - 1)
    xp = xm[idx]  # edge-clamped ("nearest") extension along the filter axis
    acc = xp[r : r + n] * weights[r]
    for j in range(r, 0, -1):
        acc = acc + (xp[r - j : r - j + n] + xp[r + j : r + j + n]) * weights[r - j]
    return np.moveaxis(acc, 0, axis)


def gaussian_restated(image: np.ndarray, sigma: float, truncate: float = 4.0) -> np.ndarray:
    """N-D Gaussian as scipy does it: axis 0, 1, ... in order, skipping sigma <= 1e-15."""
    out = np.asarray(image, dtype=np.float64)
    if sigma > 1e-15:
        w = gaussian_kernel1d(sigma, truncate)
        for axis in range(out.ndim):
            out = correlate1d_symmetric_nearest(out, w, axis)
    return out


def gaussian(image: np.ndarray, sigma: float, truncate: float = 4.0) -> np.ndarray:
    """``skimage.filters.gaussian(image, sigma, mode='nearest', truncate=4.0)`` (preserve_range
    False => ``img_as_float`` first) through the real scipy routine it dispatches to."""
    f = img_as_float(image)
    if f.dtype != np.float64:
        raise TypeError("oracle.gaussian: only float64 compute is on the hot path")
    return ndi.gaussian_filter(f, sigma, mode="nearest", cval=0, truncate=truncate)


def difference_of_gaussians(image: np.ndarray, low_sigma: float, high_sigma: float) -> np.ndarray:
    """``skimage.filters.difference_of_gaussians(image, low, high)`` with its defaults
    (mode='nearest', cval=0, truncate=4.0, channel_axis=None => every axis is spatial)."""
    if high_sigma < low_sigma:
        raise ValueError("high_sigma must be equal to or larger than low_sigma for all axes")
    f = img_as_float(image)
    im1 = ndi.gaussian_filter(f, low_sigma, mode="nearest", cval=0, truncate=4.0)
    im2 = ndi.gaussian_filter(f, high_sigma, mode="nearest", cval=0, truncate=4.0)
    return im1 - im2


def white_tophat(image: np.ndarray, size: int) -> np.ndarray:
    """Extension (no reference call site): the real ``scipy.ndimage.white_tophat`` with a flat
    ``size``-wide box, mode='reflect' -- what ``operations.subtract_background_tophat`` must equal."""
    return ndi.white_tophat(image, size=size)
