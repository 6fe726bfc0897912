"""Image operations of the preprocessing / thresholding path, on the GPU.

Signatures, defaults, validation order and error messages follow the reference's
``operations.py`` (``rescale_by_percentile`` :10-54, ``subtract_background_dog`` :57-97,
``crop_to_center`` :100-132, ``apply_threshold`` :135-216).  Arithmetic follows the scikit-image
/ SciPy / NumPy routines those lines dispatch to and runs in hand-written CUDA
kernels behind ``libamt_b200.so``; there is no CPU path.

Inputs may be NumPy arrays (result: NumPy array, as in the reference) or CUDA tensors produced
by another operation of this module (result stays on the device; ``Pipeline`` uses this to keep
a whole chain resident).  ``_batched=True`` (used by ``Pipeline(parallel=True)``) treats the
first axis as independent slices, exactly like the reference's per-slice thread map.
"""

from __future__ import annotations

from typing import Literal

import numpy as np

from . import _gpu, _lib

_THRESHOLD_METHODS = ("otsu", "li", "yen", "isodata", "mean", "minimum", "triangle", "local", "niblack", "sauvola")


_LOCAL_THRESHOLD_METHODS = ("local", "niblack", "sauvola")


def _device_op(func):
    func.__amt_device_op__ = True
    return func


def _img_as_float_wide(a: np.ndarray) -> np.ndarray:
    """``skimage.util.img_as_float`` for the integer dtypes that have no kernel of their own (int8, int16, int32,
    int64, uint32, uint64), as scikit-image's ``_convert`` does it: unsigned ``x * (1 / max)``; signed
    ``(x + 0.5) * 2 / (max - min)``.  A dtype conversion at the host boundary (the array has to become float64
    to be uploaded at all); uint8 / uint16 never come here, their scale is applied inside the kernels."""
    info = np.iinfo(a.dtype)
    if a.dtype.kind == "u":
        return np.multiply(a, 1.0 / info.max, dtype=np.float64)
    out = np.add(a, 0.5, dtype=np.float64)
    out *= 2
    out /= float(info.max) - float(info.min)
    return out


def _prepare(intensities, *, allow_bool: bool = False, as_float: bool = False):
    """-> (device tensor, numpy dtype of the logical input, was_numpy).  The tensor is C-contiguous (the kernels
    take a base pointer and plane strides: a cropped or sliced view is packed first).  Integer dtypes other than
    uint8/uint16 are promoted to float64 — with ``img_as_float``'s scale and offset when ``as_float`` (the
    Gaussian / DoG ops, which scikit-image feeds through ``img_as_float``), unscaled otherwise (exact while
    |v| < 2**53)."""
    if _gpu.is_device_array(intensities):
        torch = _gpu.torch_mod()
        if intensities.dtype in (torch.int16, torch.uint16):
            return intensities.contiguous(), np.dtype(np.uint16), False
        if intensities.dtype == torch.float64:
            return intensities.contiguous(), np.dtype(np.float64), False
        if intensities.dtype == torch.bool and allow_bool:
            return intensities.to(torch.float64).contiguous(), np.dtype(np.bool_), False
        raise TypeError(f"unsupported device dtype {intensities.dtype}")
    a = np.asarray(intensities)
    if a.dtype == np.uint16 or a.dtype == np.float64:
        return _gpu.to_device(a), a.dtype, True
    if as_float and a.dtype.kind in "iu" and a.dtype != np.uint8:
        return _gpu.to_device(_img_as_float_wide(a)), np.dtype(np.float64), True
    if a.dtype == np.uint8:
        return _gpu.to_device(a.astype(np.uint16)), a.dtype, True
    if a.dtype == np.bool_ and allow_bool:
        return _gpu.to_device(a.astype(np.float64)), a.dtype, True
    if a.dtype.kind in "iu":
        if a.size and max(abs(int(a.min())), abs(int(a.max()))) >= 2**53:
            raise TypeError("integer values beyond 2**53 are not representable on the float64 path")
        return _gpu.to_device(a.astype(np.float64)), a.dtype, True
    raise TypeError(
        f"dtype {a.dtype} is outside the B200 hot path (supported: uint8, uint16, integer, float64)"
    )


def _finish(out, was_numpy: bool):
    # inside a device-resident Pipeline chain the result stays on the device whatever the first input was
    if getattr(_gpu._tls, "keep_on_device", False):
        return out
    return _gpu.to_host(out) if was_numpy else out


def _slices(t, batched: bool):
    """View as (n_img, n) planes: one plane for the whole array unless batched."""
    if batched:
        return t.reshape(t.shape[0], -1)
    return t.reshape(1, -1)


# ---------------------------------------------------------------------------------------------
@_device_op
def rescale_by_percentile(
    intensities,
    percentile_range: tuple[float, float] = (0, 100),
    out_range: tuple[float, float] = (0, 1),
    *,
    _batched: bool = False,
):
    """Percentile-based contrast stretching (ref: ``operations.py:10-54``).

    ``p1, p2 = np.percentile(x, percentile_range)`` then
    ``skimage.exposure.rescale_intensity(x, in_range=(p1, p2), out_range=out_range)``; empty
    input -> zeros, constant input -> ``out_range[0]`` everywhere.  Output float64.
    """
    if not (0 <= percentile_range[0] < percentile_range[1] <= 100):
        raise ValueError(
            f"Invalid percentile range: {percentile_range}. "
            f"Values must be in ascending order between 0 and 100."
        )
    if not _gpu.is_device_array(intensities) and np.asarray(intensities).size == 0:
        return np.zeros_like(np.asarray(intensities), dtype=float)
    t, _, was_numpy = _prepare(intensities)
    planes = _slices(t, _batched)
    is_f64 = _gpu.dtype_code(planes) == _lib.AMT_F64
    mm = _gpu.minmax_keys(planes)
    mnmx = _gpu.minmax_values(mm, is_f64)
    pcts = _gpu.percentiles(planes, [percentile_range[0], percentile_range[1]], mm if is_f64 else None)
    o1, o2 = float(out_range[0]), float(out_range[1])
    params = []
    for i in range(planes.shape[0]):
        p = _lib.MapParams()
        p.o1, p.o2 = o1, o2
        if mnmx[i, 0] == mnmx[i, 1]:
            p.flags = _lib.AMT_MAP_FILL
        else:
            p.flags = _lib.AMT_MAP_RESCALE
            p.p1, p.p2 = float(pcts[i, 0]), float(pcts[i, 1])
        params.append(p)
    out = _gpu.apply_map(planes, params).reshape(t.shape)
    return _finish(out, was_numpy)


@_device_op
def subtract_background_dog(
    intensities,
    low_sigma: float = 0.6,
    high_sigma: float = 16.0,
    percentile: float = 0,
    *,
    _batched: bool = False,
):
    """Difference-of-Gaussians background subtraction (ref: ``operations.py:57-97``).

    ``dog = skimage.filters.difference_of_gaussians(x, low, high)`` (float64, every axis
    filtered, edge-clamped, truncate 4), ``level = np.percentile(dog, percentile)``,
    ``np.clip(dog - level, 0, None)``.  Bit-identical to the scipy-backed reference.
    """
    if not (0 <= percentile <= 100):
        raise ValueError(f"Percentile must be between 0 and 100, got {percentile}")
    if low_sigma >= high_sigma:
        raise ValueError(f"low_sigma ({low_sigma}) must be smaller than high_sigma ({high_sigma})")
    if not _gpu.is_device_array(intensities) and np.asarray(intensities).size == 0:
        return np.zeros_like(np.asarray(intensities), dtype=float)
    t, np_dtype, was_numpy = _prepare(intensities, allow_bool=True, as_float=True)
    scale = _gpu.input_scale(np_dtype)
    slice_ndim = t.ndim - 1 if _batched else t.ndim
    if slice_ndim == 2:
        stack = t if _batched else t.reshape(1, *t.shape)
        dog, mm = _gpu.dog2d(stack, scale, low_sigma, high_sigma)
        planes = dog.reshape(dog.shape[0], -1)
    else:
        torch = _gpu.torch_mod()
        parts = [t[i] for i in range(t.shape[0])] if _batched else [t]
        dogs = [_gpu.sub_f64(_gpu.gaussian_nd(p, scale, low_sigma), _gpu.gaussian_nd(p, scale, high_sigma)) for p in parts]
        planes = torch.stack([d.reshape(-1) for d in dogs])
        mm = _gpu.minmax_keys(planes)
    levels = _gpu.percentiles(planes, [percentile], mm)
    params = []
    for i in range(planes.shape[0]):
        p = _lib.MapParams()
        p.flags = _lib.AMT_MAP_SUBCLIP
        p.lvl = float(levels[i, 0])
        params.append(p)
    out = _gpu.apply_map(planes, params).reshape(t.shape)
    return _finish(out, was_numpy)


@_device_op
def gaussian_smooth(intensities, sigma: float = 1.0, *, _batched: bool = False):
    """Stand-alone Gaussian smoothing = ``skimage.filters.gaussian(x, sigma)`` (the
    reference has no such op, users wrap skimage's in an ``ImageOperation``): ``img_as_float``
    scaling for integer input, every axis filtered, mode='nearest', truncate 4, float64,
    bit-identical to ``scipy.ndimage.gaussian_filter``."""
    if sigma < 0:
        raise ValueError(f"sigma must be non-negative, got {sigma}")
    if not _gpu.is_device_array(intensities) and np.asarray(intensities).size == 0:
        return np.zeros_like(np.asarray(intensities), dtype=float)
    t, np_dtype, was_numpy = _prepare(intensities, allow_bool=True, as_float=True)
    scale = _gpu.input_scale(np_dtype)
    torch = _gpu.torch_mod()
    parts = [t[i] for i in range(t.shape[0])] if _batched else [t]
    outs = [_gpu.gaussian_nd(p, scale, sigma) for p in parts]
    out = torch.stack(outs) if _batched else outs[0]
    return _finish(out, was_numpy)


@_device_op
def subtract_background_tophat(intensities, size: int = 50, *, _batched: bool = False):
    """White top-hat (rolling-background) subtraction: ``x - opening(x)`` with a flat ``size``-wide
    box, = ``scipy.ndimage.white_tophat(x, size=size)`` bit for bit (mode='reflect'; the same as
    ``skimage.morphology.white_tophat`` with a square footprint away from the border).  Keeps the
    input dtype (uint16 stays uint16).  Extension named by the north star; the reference ships
    only the DoG (``operations.py:57-97``)."""
    size = int(size)
    if not (1 <= size <= 256):
        raise ValueError(f"size must be between 1 and 256, got {size}")
    if not _gpu.is_device_array(intensities) and np.asarray(intensities).size == 0:
        return np.array(intensities, copy=True)
    if not _gpu.is_device_array(intensities) and np.asarray(intensities).dtype not in (np.uint16, np.float64):
        raise TypeError("subtract_background_tophat supports uint16 and float64 images")
    t, np_dtype, was_numpy = _prepare(intensities)
    torch = _gpu.torch_mod()
    parts = [t[i] for i in range(t.shape[0])] if _batched else [t]
    outs = []
    for p in parts:
        eroded = _gpu.minmax_filter_nd(p.contiguous(), size, False)
        outs.append(_gpu.minmax_filter_nd(eroded, size, True, minuend=p.contiguous()))
    out = torch.stack(outs) if _batched else outs[0]
    if was_numpy:
        host = _gpu.to_host(out)
        return host.view(np.uint16) if np_dtype == np.uint16 else host
    return out


def crop_to_center(intensities, output_shape: tuple[int, int], *, _batched: bool = False):
    """Centred crop of the last two axes, clamped to the image size; returns a view
    (ref: ``operations.py:100-132``).  Pure indexing, works on NumPy arrays and CUDA tensors."""
    height, width = intensities.shape[-2:]
    crop_height = min(height, output_shape[0])
    crop_width = min(width, output_shape[1])
    top = (height - crop_height) // 2
    left = (width - crop_width) // 2
    return intensities[..., top : top + crop_height, left : left + crop_width]


crop_to_center.__amt_device_op__ = True  # type: ignore[attr-defined]


def _isodata_from_histogram(counts: np.ndarray, centers: np.ndarray):
    """skimage ``threshold_isodata`` scan (float32 counts, float64 means), first threshold."""
    counts = counts.astype("float32", copy=False)
    csuml = np.cumsum(counts)
    csumh = csuml[-1] - csuml
    csum_intensity = np.cumsum(counts * centers)
    lower = csum_intensity[:-1] / csuml[:-1]
    higher = (csum_intensity[-1] - csum_intensity[:-1]) / csumh[:-1]
    all_mean = (lower + higher) / 2.0
    distances = all_mean - centers[:-1]
    return centers[:-1][(distances >= 0) & (distances < centers[1] - centers[0])][0]


def _yen_from_histogram(counts: np.ndarray, centers: np.ndarray):
    """skimage ``threshold_yen`` scan over the normalised float32 histogram."""
    pmf = counts.astype("float32", copy=False) / counts.sum()
    p1 = np.cumsum(pmf)
    p1_sq = np.cumsum(pmf**2)
    p2_sq = np.cumsum(pmf[::-1] ** 2)[::-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        crit = np.log(((p1_sq[:-1] * p2_sq[1:]) ** -1) * (p1[:-1] * (1.0 - p1[:-1])) ** 2)
    return centers[crit.argmax()]


def _triangle_from_histogram(counts: np.ndarray, centers: np.ndarray):
    """skimage ``threshold_triangle`` on the image histogram: the bin farthest from the line that joins
    the histogram peak with the far end of its longer tail (the histogram is flipped when the left
    tail is the shorter one)."""
    hist = np.asarray(counts)
    nbins = len(hist)
    arg_peak = int(np.argmax(hist))
    peak_height = hist[arg_peak]
    nonzero = np.flatnonzero(hist)
    arg_low, arg_high = int(nonzero[0]), int(nonzero[-1])
    flip = arg_peak - arg_low < arg_high - arg_peak
    if flip:
        hist = hist[::-1]
        arg_low = nbins - arg_high - 1
        arg_peak = nbins - arg_peak - 1
    width = arg_peak - arg_low
    x1 = np.arange(width)
    y1 = hist[x1 + arg_low]
    norm = np.sqrt(peak_height**2 + width**2)
    peak_height = peak_height / norm
    width = width / norm
    length = peak_height * x1 - width * y1
    arg_level = int(np.argmax(length)) + arg_low
    if flip:
        arg_level = nbins - arg_level - 1
    return centers[arg_level]


def _uniform3_float32(x: np.ndarray) -> np.ndarray:
    """``scipy.ndimage.uniform_filter1d(x, 3)`` for a float32 line (mode='reflect'): scipy keeps ONE
    running float64 window sum, ``s0 = x[-1]+x[0]+x[1]`` and ``s[k] = s[k-1] + (x[k+1] - x[k-2])``,
    and writes ``s[k] / 3`` rounded to float32; ``np.cumsum`` performs the same additions in the same
    order (checked bit for bit against scipy in ``tests/test_host_api.py``)."""
    line = np.concatenate((x[:1], x, x[-1:])).astype(np.float64)
    first = ((0.0 + line[0]) + line[1]) + line[2]
    sums = np.cumsum(np.concatenate(([first], line[3:] - line[:-3])))
    return (sums / 3.0).astype(np.float32)


def _local_maxima(hist: np.ndarray) -> np.ndarray:
    """skimage's ``find_local_maxima_idx`` (plateau-aware scan of ``threshold_minimum``), vectorised:
    index i is a maximum when the first change before it (or the start) is a rise and hist[i+1] < hist[i]."""
    d = np.sign(np.diff(hist.astype(np.float64)))
    moving = np.flatnonzero(d)
    if moving.size == 0:
        return moving
    s = d[moving]
    prev = np.concatenate(([1.0], s[:-1]))  # the scan starts in the "rising" state
    return moving[(s < 0) & (prev > 0)]


def _minimum_from_histogram(counts: np.ndarray, centers: np.ndarray, max_num_iter: int = 10000):
    """skimage ``threshold_minimum``: smooth the float32 histogram with a 3-bin mean until only two
    maxima remain; the threshold is the lowest bin between them."""
    smooth = np.asarray(counts).astype(np.float32, copy=False)
    maxima = np.empty(0, dtype=np.intp)
    counter = -1
    for counter in range(max_num_iter):
        smooth = _uniform3_float32(smooth)
        maxima = _local_maxima(smooth)
        if len(maxima) < 3:
            break
    if len(maxima) != 2:
        raise RuntimeError("Unable to find two maxima in histogram")
    if counter == max_num_iter - 1:
        raise RuntimeError("Maximum iteration reached for histogram smoothing")
    lowest = int(np.argmin(smooth[maxima[0] : maxima[1] + 1]))
    return centers[maxima[0] + lowest]


def _li_from_histogram(counts: np.ndarray, centers: np.ndarray, tolerance=None, initial_guess=None):
    """skimage ``threshold_li`` for an integer image (minimum cross-entropy iteration on the exact
    histogram of ``image - image.min()``, float32 weights as in skimage)."""
    image_min = centers[0]
    shifted = np.arange(len(centers))  # bin centres of the shifted image
    tolerance = tolerance or 0.5
    n = int(np.asarray(counts).sum())
    if initial_guess is None:
        # np.mean of the shifted integer image: float64 accumulation of integers is exact below 2**53
        t_next = np.float64(int((np.asarray(counts, dtype=np.int64) * shifted).sum())) / np.float64(n)
    elif callable(initial_guess):
        raise NotImplementedError("a callable initial_guess would need the image on the host")
    elif np.isscalar(initial_guess):
        t_next = initial_guess - float(image_min)
        image_max = shifted[-1] + image_min
        if not 0 < t_next < shifted[-1]:
            raise ValueError(
                f"The initial guess for threshold_li must be within the range of the image. Got {initial_guess} "
                f"for image min {image_min} and max {image_max}."
            )
    else:
        raise TypeError("Incorrect type for `initial_guess`; should be a floating point value, or a function "
                        "mapping an array to a floating point value.")
    t_curr = -2 * tolerance
    hist = np.asarray(counts).astype(np.float32, copy=False)
    while abs(t_next - t_curr) > tolerance:
        t_curr = t_next
        foreground = shifted > t_curr
        background = ~foreground
        mean_fore = np.average(shifted[foreground], weights=hist[foreground])
        mean_back = np.average(shifted[background], weights=hist[background])
        if mean_back == 0:
            break
        t_next = (mean_back - mean_fore) / (np.log(mean_back) - np.log(mean_fore))
    return t_next + image_min


_THRESHOLD_KWARGS = {
    "otsu": ("nbins",), "yen": ("nbins",), "isodata": ("nbins",), "triangle": ("nbins",), "mean": (),
    "minimum": ("nbins", "max_num_iter"), "li": ("tolerance", "initial_guess"),
    "niblack": ("window_size", "k"), "sauvola": ("window_size", "k", "r"),
    # threshold_local's own `method` can never arrive: apply_threshold's `method` parameter takes the name
    "local": ("block_size", "offset", "mode", "param", "cval"),
}


def _check_threshold_kwargs(method: str, kwargs: dict):
    """The reference forwards ``**kwargs`` to the scikit-image function (``operations.py:214``): accept the
    ones this path implements, refuse the rest loudly (an unknown name is a TypeError there as well)."""
    for name in kwargs:
        if name not in _THRESHOLD_KWARGS[method]:
            raise TypeError(f"threshold_{method}() got an unexpected keyword argument '{name}'")
    nbins = kwargs.get("nbins", 256)
    if not (isinstance(nbins, (int, np.integer)) and 2 <= nbins <= 1 << 20):
        raise ValueError(f"nbins must be an integer between 2 and 2**20, got {nbins!r}")


def _otsu_from_histogram(counts: np.ndarray, centers: np.ndarray):
    """skimage ``threshold_otsu`` scan (float32 class weights, float64 class means, first maximum); the
    host twin of ``amt_otsu``, used for integer images that reach the GPU shifted by their minimum."""
    counts = np.asarray(counts).astype(np.float32, copy=False)
    weight1 = np.cumsum(counts)
    weight2 = np.cumsum(counts[::-1])[::-1]
    mean1 = np.cumsum(counts * centers) / weight1
    mean2 = (np.cumsum((counts * centers)[::-1]) / weight2[::-1])[::-1]
    variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
    return centers[int(np.argmax(variance12))]


def _shift_wide_integers(intensities):
    """NumPy integer images other than uint8 / uint16 (scikit-image histograms them value by value,
    whatever the dtype): travel as ``image - min`` in uint16 when the value range allows it; the
    histogram scans then see the true bin centres again.  -> (array for the GPU, offset)."""
    if _gpu.is_device_array(intensities):
        return intensities, 0
    a = np.asarray(intensities)
    if a.dtype.kind not in "iu" or a.dtype in (np.uint8, np.uint16):
        return a, 0
    lo, hi = int(a.min()), int(a.max())
    if hi - lo > 65535:
        raise NotImplementedError(
            f"{a.dtype} image spanning {hi - lo + 1} values: the exact per-value histogram of the B200 path holds 65536 bins")
    shifted = a - np.uint64(lo) if a.dtype == np.uint64 else a.astype(np.int64) - lo
    return shifted.astype(np.uint16), lo


def _apply_histogram_threshold(intensities, method: str, batched: bool, kwargs: dict, offset: int = 0):
    """li / isodata / yen / mean / minimum / triangle: the per-pixel passes (min/max, histogram,
    comparison) run on the GPU; the scan over the <= 65536 histogram bins is skimage's own NumPy
    arithmetic on the host (ref: ``operations.py:185-196`` -> ``ski.filters.threshold_*`` [3p])."""
    t, np_dtype, was_numpy = _prepare(intensities)
    planes = _slices(t, batched)
    integer_image = np_dtype.kind in "iu" and planes.dtype != _gpu.torch_mod().float64
    thr = np.empty(planes.shape[0], dtype=np.float64)
    if method == "li" and not integer_image:
        # float image: scikit-image iterates over means of the thresholded PIXELS and takes its tolerance from the
        # sorted unique values; both are device passes (csrc/li.cu), the scalar recurrence is NumPy's
        limits = _gpu.minmax_values(_gpu.minmax_keys(planes), True)
        if not np.all(np.isfinite(limits)):
            raise NotImplementedError("method 'li' on a float image holding NaN or infinite values is not built")
        for i in range(planes.shape[0]):
            if limits[i, 0] == limits[i, 1]:
                thr[i] = limits[i, 0]
            else:
                thr[i] = _gpu.li_threshold_f64(planes[i], limits[i, 0], kwargs.get("tolerance"), kwargs.get("initial_guess"))
        d_thr = _gpu.torch_mod().from_numpy(thr).to(planes.device)
        mask = _gpu.threshold_gt(planes, d_thr).reshape(t.shape).view(_gpu.torch_mod().bool)
        return _finish(mask, was_numpy)
    if method == "mean" and not integer_image:
        # np.mean(image): NumPy's pairwise float64 sum, reproduced bit for bit on the device, / n
        limits = _gpu.minmax_values(_gpu.minmax_keys(planes), True)
        sums = _gpu.plane_sums_f64(planes)
        for i in range(planes.shape[0]):
            thr[i] = limits[i, 0] if limits[i, 0] == limits[i, 1] else float(np.float64(sums[i]) / np.float64(planes.shape[1]))
        d_thr = _gpu.torch_mod().from_numpy(thr).to(planes.device)
        mask = _gpu.threshold_gt(planes, d_thr).reshape(t.shape).view(_gpu.torch_mod().bool)
        return _finish(mask, was_numpy)
    hists = _gpu.plane_histograms(planes, nbins=int(kwargs.get("nbins", 256)))
    for i, (counts, centers) in enumerate(hists):
        centers = centers + offset if offset else centers
        if len(centers) == 1 or counts.sum() == counts.max():  # constant plane: nothing is above it
            thr[i] = float(centers[int(np.argmax(counts))])
        elif method == "otsu":
            thr[i] = float(_otsu_from_histogram(counts, centers))
        elif method == "isodata":
            thr[i] = float(_isodata_from_histogram(counts, centers))
        elif method == "yen":
            thr[i] = float(_yen_from_histogram(counts, centers))
        elif method == "triangle":
            thr[i] = float(_triangle_from_histogram(counts, centers))
        elif method == "minimum":
            thr[i] = float(_minimum_from_histogram(counts, centers, kwargs.get("max_num_iter", 10000)))
        elif method == "li":
            thr[i] = float(_li_from_histogram(counts, centers, kwargs.get("tolerance"), kwargs.get("initial_guess")))
        else:
            # np.mean of an integer image: exact integer sum (float64 accumulation is exact below 2**53)
            thr[i] = float(np.float64(int((counts * centers).sum())) / np.float64(planes.shape[1]))
    if offset:  # integer pixels: x > t  <=>  x > floor(t)  <=>  x - offset > floor(t) - offset, all exact
        thr = np.floor(thr) - offset
    d_thr = _gpu.torch_mod().from_numpy(thr).to(planes.device)
    mask = _gpu.threshold_gt(planes, d_thr).reshape(t.shape).view(_gpu.torch_mod().bool)
    return _finish(mask, was_numpy)


def _window_per_axis(size, ndim: int, what: str) -> tuple[int, ...]:
    sizes = (size,) * ndim if np.isscalar(size) else tuple(size)
    if len(sizes) != ndim:
        raise ValueError(f"{what} must be a scalar or have one entry per image axis ({ndim}), got {sizes}")
    return tuple(int(v) for v in sizes)


def _apply_local_threshold(intensities, method: str, batched: bool, kwargs: dict):
    """local / niblack / sauvola: per-pixel thresholds from a window around every pixel
    (ref: ``operations.py:193-195`` -> ``ski.filters.threshold_local / _niblack / _sauvola`` [3p]).

    niblack, sauvola: box mean m and standard deviation s over ``window_size`` (odd), ``m - k*s`` and
    ``m*(1 + k*(s/r - 1))``.  uint8 / uint16 images: exact integer window sums (scikit-image's float64 integral
    images are exact integer arithmetic there, so the result is bit-identical); float64 images: the float64
    integral images themselves, summed in ``np.cumsum``'s order (``amt_window_threshold_f64``).  local: Gaussian-weighted mean (threshold_local's default method, the only one reachable
    through the reference's signature, whose own ``method`` parameter takes the name; sigma =
    (block_size-1)/6 unless ``param`` is given, ``mode`` 'reflect' or 'nearest') minus ``offset``."""
    torch = _gpu.torch_mod()
    t, np_dtype, was_numpy = _prepare(intensities)
    planes = t if batched else t.reshape((1,) + tuple(t.shape))
    if planes.ndim != 3:
        raise NotImplementedError(f"method '{method}' is built for 2-D images (got {planes.ndim - 1} axes per image)")
    n_img, h, w = planes.shape
    integer_image = np_dtype.kind in "iu" and planes.dtype != torch.float64
    flat = planes.reshape(n_img, -1)
    mm = _gpu.minmax_keys(flat)
    if method == "local":
        block = _window_per_axis(kwargs.get("block_size", 3), 2, "block_size")
        if any(b % 2 == 0 for b in block):
            raise ValueError(f"block_size must be odd! Given block_size {block} contains even values.")
        modes = {"reflect": _lib.AMT_EXTEND_REFLECT, "nearest": _lib.AMT_EXTEND_NEAREST}
        mode = kwargs.get("mode", "reflect")
        if mode not in modes:
            raise NotImplementedError(f"threshold_local: mode '{mode}' is not built (available: reflect, nearest)")
        param = kwargs.get("param")
        sigma = tuple((b - 1) / 6.0 for b in block) if param is None else param
        masks = []
        for i in range(n_img):  # the image as float64 without rescaling (astype), every axis filtered
            smooth = _gpu.gaussian_nd(planes[i], 1.0, sigma, modes[mode])
            masks.append(_gpu.threshold_gt_image(planes[i], smooth, float(kwargs.get("offset", 0))))
        mask = torch.stack(masks)
    else:
        window = _window_per_axis(kwargs.get("window_size", 15), 2, "window_size")
        if any(v % 2 == 0 for v in window):
            raise ValueError(
                "Window size for `threshold_sauvola` or `threshold_niblack` must not be even on any dimension. "
                f"Got {window}")
        if not integer_image:
            # float image: scikit-image's float64 integral images, summed in np.cumsum's order on the device
            if window[0] // 2 + 1 >= h or window[1] // 2 + 1 >= w:
                raise NotImplementedError(f"window_size {window} on a {h}x{w} image: windows smaller than twice the image are built")
            k = kwargs.get("k", 0.2)
            r = kwargs.get("r")
            if r is None:  # dtype_limits(float image, clip_negative=False) = (-1, 1)
                r = 1.0
            mask, _ = _gpu.window_threshold_f64(planes, window, 1 if method == "sauvola" else 0, k, r)
            limits = _gpu.minmax_values(mm, True)
            for i in np.flatnonzero(limits[:, 0] == limits[:, 1]):
                mask[int(i)].zero_()
            mask = mask.reshape(t.shape).view(torch.bool)
            return _finish(mask, was_numpy)
        if max(window) > 127 or window[0] // 2 >= max(h, 2) or window[1] // 2 >= max(w, 2):
            raise NotImplementedError(f"window_size {window} on a {h}x{w} image: windows up to 127 and smaller than "
                                      "twice the image are built")
        # every float64 sum scikit-image forms (integral image of the squared, padded image) stays an exact integer
        for counts, values in _gpu.plane_histograms(flat, mm):
            if 4 * int((counts * values.astype(np.int64) ** 2).sum()) >= 2**53:
                raise NotImplementedError("sum of squared intensities beyond 2**53 / 4: scikit-image's float64 integral "
                                          "image rounds there, which the exact integer path does not reproduce")
        k = kwargs.get("k", 0.2)
        r = kwargs.get("r")
        if r is None:  # dtype_limits(image, clip_negative=False): half the dtype's range
            r = 0.5 * (float(np.iinfo(np_dtype).max) - float(np.iinfo(np_dtype).min))
        mask, _ = _gpu.window_threshold_u16(planes, window, 1 if method == "sauvola" else 0, k, r)
    # a constant image is all False before the method is looked at (operations.py:201-202)
    limits = _gpu.minmax_values(mm, planes.dtype == torch.float64)
    for i in np.flatnonzero(limits[:, 0] == limits[:, 1]):
        mask[int(i)].zero_()
    mask = mask.reshape(t.shape).view(torch.bool)
    return _finish(mask, was_numpy)


@_device_op
def apply_threshold(
    intensities,
    method: Literal["otsu", "li", "yen", "isodata", "mean", "minimum", "triangle", "local", "niblack", "sauvola"] = "otsu",
    *,
    _batched: bool = False,
    **kwargs,
):
    """Binary image ``intensities > threshold`` (ref: ``operations.py:135-216``).

    Empty or constant input -> all False (checked before the method name, as in the
    reference).  All ten methods of the reference run on the B200 path: ``otsu``, ``li``, ``yen``,
    ``isodata``, ``mean``, ``minimum``, ``triangle`` from skimage's histogram (exact per-value counts for
    integer images, ``nbins`` uniform bins for float images; ``li`` on a float image from the pixels themselves), and
    the local-window methods ``local`` (Gaussian), ``niblack``, ``sauvola``.
    """
    if not _gpu.is_device_array(intensities) and np.asarray(intensities).size == 0:
        return np.zeros_like(np.asarray(intensities), dtype=bool)
    method_lower = method.lower()
    if method_lower not in _THRESHOLD_METHODS:
        # the reference returns all-False for a constant image before it looks at the method
        host = _gpu.to_host(intensities) if _gpu.is_device_array(intensities) else np.asarray(intensities)
        if host.min() == host.max():
            return np.zeros_like(host, dtype=bool)
        raise ValueError(
            f"Unsupported thresholding method: '{method}'. "
            f"Supported methods: {', '.join(_THRESHOLD_METHODS)}"
        )
    _check_threshold_kwargs(method_lower, kwargs)
    if method_lower in _LOCAL_THRESHOLD_METHODS:
        return _apply_local_threshold(intensities, method_lower, _batched, kwargs)
    intensities, offset = _shift_wide_integers(intensities)
    float_nbins = kwargs.get("nbins", 256) != 256 and not (
        _gpu.is_device_array(intensities) and _gpu.dtype_code(intensities) == _lib.AMT_U16
        or not _gpu.is_device_array(intensities) and np.asarray(intensities).dtype.kind in "iub")
    if method_lower != "otsu" or offset or float_nbins:
        return _apply_histogram_threshold(intensities, method_lower, _batched, kwargs, offset)
    t, _, was_numpy = _prepare(intensities)
    planes = _slices(t, _batched)
    # a constant plane gets threshold == its value, so nothing is above it (all False)
    thr, _ = _gpu.otsu_threshold(planes)
    mask = _gpu.threshold_gt(planes, thr).reshape(t.shape).view(_gpu.torch_mod().bool)
    return _finish(mask, was_numpy)
