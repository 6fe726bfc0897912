"""Oracle: image histogram + Otsu threshold.  TEST INFRASTRUCTURE ONLY.

Restates ``skimage.filters.threshold_otsu(image)`` (nbins=256) as called by the reference at
``operations.py:186`` / ``:214``.
"""

from __future__ import annotations

import numpy as np


def histogram_float_256(image: np.ndarray, nbins: int = 256):
    """``skimage.exposure.histogram`` for float images: ``np.histogram(image, bins=nbins,
    range=None)`` and bin centres ``(edges[:-1] + edges[1:]) / 2``."""
    hist, edges = np.histogram(np.asarray(image).reshape(-1), bins=nbins, range=None)
    centers = (edges[:-1] + edges[1:]) / 2.0
    return hist, centers


def linspace_edges(first: float, last: float, nbins: int = 256) -> np.ndarray:
    """The bin edges numpy builds for a uniform histogram (``np.linspace(first, last,
    nbins+1)``); if ``first == last`` numpy widens the range by +-0.5 first."""
    first = float(first)
    last = float(last)
    if first == last:
        first, last = first - 0.5, last + 0.5
    return np.linspace(first, last, nbins + 1, endpoint=True, dtype=np.float64)


def histogram_float_restated(image: np.ndarray, nbins: int = 256):
    """Candidate-then-correct binning of numpy's uniform-bin fast path
    (``numpy/lib/_histograms_impl.py``): ``i = int(((x-first)/(last-first))*nbins)``,
    ``nbins -> nbins-1``, then ``if x < edges[i]: i -= 1`` and
    ``if x >= edges[i+1] and i != nbins-1: i += 1``.  This is what the CUDA kernel does."""
    x = np.asarray(image, dtype=np.float64).reshape(-1)
    first, last = float(x.min()), float(x.max())
    edges = linspace_edges(first, last, nbins)
    first, last = float(edges[0]), float(edges[-1])
    f = ((x - first) / (last - first)) * nbins
    idx = f.astype(np.intp)
    idx[idx == nbins] -= 1
    dec = x < edges[idx]
    idx[dec] -= 1
    inc = (x >= edges[idx + 1]) & (idx != nbins - 1)
    idx[inc] += 1
    hist = np.bincount(idx, minlength=nbins).astype(np.int64)
    centers = (edges[:-1] + edges[1:]) / 2.0
    return hist, centers


def histogram_int(image: np.ndarray):
    """``skimage.exposure.histogram`` for integer images: exact per-value ``bincount`` over
    ``[min, max]`` with ``bin_centers = arange(min, max+1)`` (nbins ignored)."""
    flat = np.asarray(image).reshape(-1)
    imin, imax = int(flat.min()), int(flat.max())
    hist = np.bincount((flat.astype(np.int64) - imin), minlength=imax - imin + 1)
    centers = np.arange(imin, imax + 1)
    return hist, centers


def otsu_from_histogram(hist: np.ndarray, centers: np.ndarray):
    """The Otsu scan with numpy's mixed precision: float32 counts and class weights,
    float64 class means, float32 product ``w1*w2`` promoted on the last multiply; first
    maximum wins."""
    counts = np.asarray(hist).astype("float32", copy=False)
    weight1 = np.cumsum(counts)
    weight2 = np.cumsum(counts[::-1])[::-1]
    mean1 = np.cumsum(counts * centers) / weight1
    mean2 = (np.cumsum((counts * centers)[::-1]) / weight2[::-1])[::-1]
    variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
    idx = int(np.argmax(variance12))
    return centers[idx]


def threshold_otsu(image: np.ndarray, nbins: int = 256):
    image = np.asarray(image)
    first_pixel = image.reshape(-1)[0]
    if np.all(image == first_pixel):
        return first_pixel
    if np.issubdtype(image.dtype, np.integer):
        hist, centers = histogram_int(image)
    else:
        hist, centers = histogram_float_256(image, nbins)
    return otsu_from_histogram(hist, centers)


# ---- further histogram-based methods of the reference's ``apply_threshold`` (operations.py:185-196).
# scikit-image is not installable here: these restate skimage 0.25.2 ``filters/thresholding.py`` from
# the published source (parity unpinned beyond this restatement).
def _image_histogram(image: np.ndarray, nbins: int = 256):
    image = np.asarray(image)
    if np.issubdtype(image.dtype, np.integer):
        return histogram_int(image)
    return histogram_float_256(image, nbins)


def isodata_from_histogram(hist: np.ndarray, centers: np.ndarray):
    """``threshold_isodata``: the first bin centre t with 0 <= (mean(<=t) + mean(>t))/2 - t < bin width."""
    if len(centers) == 1:
        return centers[0]
    counts = np.asarray(hist).astype("float32", copy=False)
    csuml = np.cumsum(counts)
    csumh = csuml[-1] - csuml
    intensity_sum = counts * centers
    csum_intensity = np.cumsum(intensity_sum)
    lower = csum_intensity[:-1] / csuml[:-1]
    higher = (csum_intensity[-1] - csum_intensity[:-1]) / csumh[:-1]
    all_mean = (lower + higher) / 2.0
    bin_width = centers[1] - centers[0]
    distances = all_mean - centers[:-1]
    thresholds = centers[:-1][(distances >= 0) & (distances < bin_width)]
    return thresholds[0]


def yen_from_histogram(hist: np.ndarray, centers: np.ndarray):
    """``threshold_yen``: maximum of Yen's criterion over the normalised float32 histogram."""
    if len(centers) == 1:
        return centers[0]
    counts = np.asarray(hist)
    pmf = counts.astype("float32", copy=False) / counts.sum()
    P1 = np.cumsum(pmf)
    P1_sq = np.cumsum(pmf**2)
    P2_sq = np.cumsum(pmf[::-1] ** 2)[::-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        crit = np.log(((P1_sq[:-1] * P2_sq[1:]) ** -1) * (P1[:-1] * (1.0 - P1[:-1])) ** 2)
    return centers[crit.argmax()]


def threshold_isodata(image: np.ndarray, nbins: int = 256):
    return isodata_from_histogram(*_image_histogram(image, nbins))


def threshold_yen(image: np.ndarray, nbins: int = 256):
    return yen_from_histogram(*_image_histogram(image, nbins))


def threshold_mean(image: np.ndarray):
    return np.mean(image)


def threshold_triangle(image: np.ndarray, nbins: int = 256):
    """``threshold_triangle`` (skimage 0.25.2 ``filters/thresholding.py``): histogram peak, the longer
    tail (the histogram is flipped when the left tail is the shorter one), and the level with the largest distance to the peak-to-tail-end line."""
    image = np.asarray(image)
    hist, bin_centers = _image_histogram(image.reshape(-1), nbins)
    nbins = len(hist)
    arg_peak_height = np.argmax(hist)
    peak_height = hist[arg_peak_height]
    arg_low_level, arg_high_level = np.flatnonzero(hist)[[0, -1]]
    if arg_low_level == arg_high_level:
        return image.ravel()[0]
    flip = arg_peak_height - arg_low_level < arg_high_level - arg_peak_height
    if flip:
        hist = hist[::-1]
        arg_low_level = nbins - arg_high_level - 1
        arg_peak_height = nbins - arg_peak_height - 1
    del arg_high_level
    width = arg_peak_height - arg_low_level
    x1 = np.arange(width)
    y1 = hist[x1 + arg_low_level]
    norm = np.sqrt(peak_height**2 + width**2)
    peak_height = peak_height / norm
    width = width / norm
    length = peak_height * x1 - width * y1
    arg_level = np.argmax(length) + arg_low_level
    if flip:
        arg_level = nbins - arg_level - 1
    return bin_centers[arg_level]


def threshold_minimum(image: np.ndarray, nbins: int = 256, max_num_iter: int = 10000):
    """``threshold_minimum``: the histogram is smoothed with the REAL ``scipy.ndimage.uniform_filter1d``
    (size 3, float32) until two maxima remain; the maxima scan is skimage's plateau-aware loop."""
    from scipy import ndimage as ndi

    def find_local_maxima_idx(hist):
        maximum_idxs = []
        direction = 1
        for i in range(hist.shape[0] - 1):
            if direction > 0:
                if hist[i + 1] < hist[i]:
                    direction = -1
                    maximum_idxs.append(i)
            else:
                if hist[i + 1] > hist[i]:
                    direction = 1
        return maximum_idxs

    counts, bin_centers = _image_histogram(np.asarray(image).reshape(-1), nbins)
    smooth_hist = counts.astype("float32", copy=False)
    maximum_idxs = []
    counter = -1
    for counter in range(max_num_iter):
        smooth_hist = ndi.uniform_filter1d(smooth_hist, 3)
        maximum_idxs = find_local_maxima_idx(smooth_hist)
        if len(maximum_idxs) < 3:
            break
    if len(maximum_idxs) != 2:
        raise RuntimeError("Unable to find two maxima in histogram")
    elif counter == max_num_iter - 1:
        raise RuntimeError("Maximum iteration reached for histogram smoothing")
    minimum_idx = np.argmin(smooth_hist[maximum_idxs[0] : maximum_idxs[1] + 1])
    return bin_centers[maximum_idxs[0] + minimum_idx]


def threshold_li(image: np.ndarray, *, tolerance=None, initial_guess=None):
    """``threshold_li`` (minimum cross-entropy): iterate t <- (mean_back - mean_fore) / (log mean_back -
    log mean_fore) on ``image - image.min()``; integer images iterate on their exact histogram with
    float32 weights (``np.average``), float images on the pixels themselves."""
    image = np.asarray(image)
    image = image[~np.isnan(image)]
    if image.size == 0:
        return np.nan
    if np.all(image == image.flat[0]):
        return image.flat[0]
    image = image[np.isfinite(image)]
    if image.size == 0:
        return 0.0
    image_min = np.min(image)
    image -= image_min
    if image.dtype.kind in "iu":
        tolerance = tolerance or 0.5
    else:
        tolerance = tolerance or np.min(np.diff(np.unique(image))) / 2
    if initial_guess is None:
        t_next = np.mean(image)
    elif callable(initial_guess):
        t_next = initial_guess(image)
    elif np.isscalar(initial_guess):
        t_next = initial_guess - float(image_min)
        if not 0 < t_next < np.max(image):
            raise ValueError("The initial guess for threshold_li must be within the range of the image.")
    else:
        raise TypeError("Incorrect type for `initial_guess`")
    t_curr = -2 * tolerance
    if image.dtype.kind in "iu":
        hist, bin_centers = histogram_int(image.reshape(-1))
        hist = hist.astype("float32", copy=False)
        while abs(t_next - t_curr) > tolerance:
            t_curr = t_next
            foreground = bin_centers > t_curr
            background = ~foreground
            mean_fore = np.average(bin_centers[foreground], weights=hist[foreground])
            mean_back = np.average(bin_centers[background], weights=hist[background])
            if mean_back == 0:
                break
            t_next = (mean_back - mean_fore) / (np.log(mean_back) - np.log(mean_fore))
    else:
        while abs(t_next - t_curr) > tolerance:
            t_curr = t_next
            foreground = image > t_curr
            mean_fore = np.mean(image[foreground])
            mean_back = np.mean(image[~foreground])
            if mean_back == 0.0:
                break
            t_next = (mean_back - mean_fore) / (np.log(mean_back) - np.log(mean_fore))
    return t_next + image_min


# ---- local-window methods (operations.py:193-195).  Restated from skimage 0.25.2 ``filters/thresholding.py``
# (``_mean_std``, ``threshold_niblack``, ``threshold_sauvola``, ``threshold_local``) on the real numpy / scipy
# routines it calls (``np.pad``, ``cumsum``, ``scipy.ndimage.gaussian_filter``); parity unpinned beyond that.
def _integral_image(image: np.ndarray) -> np.ndarray:
    """``skimage.transform.integral_image(image, dtype=float64)``: cumulative sums along every axis in turn."""
    S = image
    for axis in range(image.ndim):
        S = S.cumsum(axis=axis, dtype=np.float64)
    return S


def _correlate_sparse(image: np.ndarray, kernel_shape, kernel_indices, kernel_values) -> np.ndarray:
    """``skimage.filters._sparse._correlate_sparse`` (mode 'valid'): the first corner is copied, the others
    are added in the order given."""
    out_shape = tuple(s - k + 1 for s, k in zip(image.shape, kernel_shape))
    first_idx, first_val = kernel_indices[0], kernel_values[0]
    assert tuple(first_idx) == (0,) * image.ndim
    out = image[tuple(slice(0, s) for s in out_shape)].copy()
    if first_val != 1:
        out *= first_val
    for idx, val in zip(kernel_indices[1:], kernel_values[1:]):
        out += image[tuple(slice(i, i + s) for i, s in zip(idx, out_shape))] * val
    return out


def _mean_std(image: np.ndarray, w):
    import itertools
    import math

    image = np.asarray(image)
    if np.isscalar(w):
        w = (w,) * image.ndim
    for axis_size in w:
        if axis_size % 2 == 0:
            raise ValueError(
                f"Window size for `threshold_sauvola` or `threshold_niblack` must not be even on any dimension. Got {w}")
    float_dtype = np.float64  # _supported_float_type of integer / float64 images
    pad_width = tuple((k // 2 + 1, k // 2) for k in w)
    padded = np.pad(image.astype(float_dtype, copy=False), pad_width, mode="reflect")
    integral = _integral_image(padded)
    padded *= padded
    integral_sq = _integral_image(padded)
    kernel_indices = list(itertools.product(*tuple([(0, _w) for _w in w])))
    kernel_values = [(-1) ** (image.ndim % 2 != np.sum(indices) % 2) for indices in kernel_indices]
    total_window_size = math.prod(w)
    kernel_shape = tuple(_w + 1 for _w in w)
    m = _correlate_sparse(integral, kernel_shape, kernel_indices, kernel_values)
    m = m.astype(float_dtype, copy=False)
    m /= total_window_size
    g2 = _correlate_sparse(integral_sq, kernel_shape, kernel_indices, kernel_values)
    g2 = g2.astype(float_dtype, copy=False)
    g2 /= total_window_size
    s = np.sqrt(np.clip(g2 - m * m, 0, None))
    return m, s


def threshold_niblack(image: np.ndarray, window_size=15, k=0.2):
    m, s = _mean_std(image, window_size)
    return m - k * s


def threshold_sauvola(image: np.ndarray, window_size=15, k=0.2, r=None):
    image = np.asarray(image)
    if r is None:
        # dtype_limits(image, clip_negative=False)
        if image.dtype.kind in "iu":
            imin, imax = np.iinfo(image.dtype).min, np.iinfo(image.dtype).max
        else:
            imin, imax = -1.0, 1.0
        r = 0.5 * (imax - imin)
    m, s = _mean_std(image, window_size)
    return m * (1 + k * ((s / r) - 1))


def threshold_local(image: np.ndarray, block_size=3, method="gaussian", offset=0, mode="reflect", param=None, cval=0):
    from scipy import ndimage as ndi

    image = np.asarray(image)
    if np.isscalar(block_size):
        block_size = (block_size,) * image.ndim
    if any(b % 2 == 0 for b in block_size):
        raise ValueError(f"block_size must be odd! Given block_size {block_size} contains even values.")
    if method != "gaussian":
        raise NotImplementedError("oracle restates method='gaussian' only")
    image = image.astype(np.float64, copy=False)
    thresh_image = np.zeros(image.shape, dtype=np.float64)
    sigma = tuple((b - 1) / 6.0 for b in block_size) if param is None else param
    ndi.gaussian_filter(image, sigma, output=thresh_image, mode=mode, cval=cval, truncate=4.0)
    return thresh_image - offset
