"""Device-side building blocks: torch tensors own the memory, libamt_b200 does the work.

PyTorch is plumbing only (allocation, pinned staging, streams); every pixel is touched by a
hand-written sm_100a kernel behind the C ABI.  All functions here take/return CUDA tensors
and are batched over a leading plane axis, so a ``Pipeline(parallel=True)`` stack or a chunk
of fields of view is one launch per stage.
"""

from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib
from ._lib import AMT_F64, AMT_U16, check

_tls = threading.local()


def torch_mod():
    import torch

    return torch


def require_cuda(device: int | None = None):
    """Return the torch device to use; fail loudly when there is none (no CPU fallback)."""
    torch = torch_mod()
    if not torch.cuda.is_available():
        raise _lib.AmtLibraryError(
            "arcadia_microscopy_tools_b200 needs an NVIDIA B200 (sm_100a) CUDA device; "
            "there is no CPU fallback"
        )
    _lib.load()
    idx = torch.cuda.current_device() if device is None else int(device)
    return torch.device("cuda", idx)


def stream_ptr() -> int:
    return int(torch_mod().cuda.current_stream().cuda_stream)


def ptr(t) -> int:
    return 0 if t is None else int(t.data_ptr())


def is_device_array(x) -> bool:
    torch = torch_mod()
    return isinstance(x, torch.Tensor) and x.is_cuda


# ------------------------------------------------------------------ host <-> device
def to_device(a: np.ndarray, device=None):
    """Upload a NumPy array (uint16 travels as int16 bits: torch's uint16 support is partial)."""
    torch = torch_mod()
    dev = require_cuda() if device is None else device
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint16:
        return torch.from_numpy(a.view(np.int16)).to(dev, non_blocking=False)
    if a.dtype == np.bool_:
        return torch.from_numpy(a.view(np.uint8)).to(dev)
    return torch.from_numpy(a).to(dev)


def to_host(t) -> np.ndarray:
    return t.detach().cpu().numpy()


def dtype_code(t) -> int:
    torch = torch_mod()
    if t.dtype in (torch.int16, torch.uint16):
        return AMT_U16
    if t.dtype == torch.float64:
        return AMT_F64
    raise TypeError(f"unsupported device dtype {t.dtype}")


# ------------------------------------------------------------------ Gaussian weights (host, NumPy)
def gaussian_half_weights(sigma: float, truncate: float = 4.0) -> np.ndarray:
    """``weights[c - j]`` for j = 0..r of scipy's ``_gaussian_kernel1d(sigma, 0, r)``, computed
    with the same NumPy expression scipy uses so the weights are bit-identical."""
    sd = float(sigma)
    if sd <= 1e-15:  # scipy skips the axis: identity
        return np.ones(1, dtype=np.float64)
    radius = int(truncate * sd + 0.5)
    sigma2 = sd * sd
    x = np.arange(-radius, radius + 1)
    phi_x = np.exp(-0.5 / sigma2 * x**2)
    phi_x = phi_x / phi_x.sum()
    return np.ascontiguousarray(phi_x[radius::-1])


def rank_pair(n: int, q: float) -> tuple[int, int, float]:
    """Floor / ceil ranks and lerp fraction of ``np.percentile(a, q)`` over n samples."""
    quant = np.true_divide(np.float64(q), 100)
    v = (n - 1) * quant
    lo = int(np.floor(v))
    lo = min(max(lo, 0), n - 1)
    hi = min(lo + 1, n - 1)
    return lo, hi, float(v - lo)


def np_lerp(a: float, b: float, t: float) -> float:
    a, b, t = np.float64(a), np.float64(b), np.float64(t)
    diff = b - a
    return float(b - diff * (1 - t)) if t >= 0.5 else float(a + diff * t)


# ------------------------------------------------------------------ kernels
def input_scale(np_dtype) -> float:
    """img_as_float's multiplier for the integer dtypes on the path."""
    if np_dtype == np.uint16:
        return 1.0 / 65535.0
    if np_dtype == np.uint8:
        return 1.0 / 255.0
    return 1.0


def dog2d(x, scale: float, low_sigma: float, high_sigma: float):
    """x: (n_img, H, W) int16-as-uint16 or float64 -> (dog float64, minmax keys)."""
    torch = torch_mod()
    lib = _lib.load()
    n_img, h, w = x.shape
    hw_lo = gaussian_half_weights(low_sigma)
    hw_hi = gaussian_half_weights(high_sigma)
    d_lo = torch.from_numpy(hw_lo).to(x.device)
    d_hi = torch.from_numpy(hw_hi).to(x.device)
    out = torch.empty((n_img, h, w), dtype=torch.float64, device=x.device)
    tmp_lo = torch.empty_like(out)
    tmp_hi = torch.empty_like(out)
    mm = torch.empty((n_img, 2), dtype=torch.int64, device=x.device)
    check(
        lib.amt_dog2d(ptr(x), dtype_code(x), scale, ptr(out), n_img, h, w, ptr(d_lo), len(hw_lo) - 1, ptr(d_hi),
                      len(hw_hi) - 1, ptr(tmp_lo), ptr(tmp_hi), ptr(mm), stream_ptr()),
        "amt_dog2d",
    )
    return out, mm


class TensorCoreGaussian:
    """Handle of the tcgen05 / TMA Gaussian (``amt_tcg_*``): quantised weights + band tiles on the device."""

    def __init__(self, sigma: float, device=None):
        torch = torch_mod()
        self.dev = require_cuda() if device is None else device
        self.lib = _lib.load()
        self.hw = gaussian_half_weights(sigma)
        self.radius = len(self.hw) - 1
        self.handle = C.c_void_p()
        check(self.lib.amt_tcg_create(self.hw.ctypes.data, self.radius, self.dev.index or 0, C.byref(self.handle)),
              "amt_tcg_create")
        w = np.zeros(self.radius + 1, dtype=np.uint64)
        s = C.c_int()
        check(self.lib.amt_tcg_weights(self.handle, w.ctypes.data, C.byref(s)), "amt_tcg_weights")
        self.int_weights, self.scale_bits = w, int(s.value)
        self._torch = torch

    def close(self):
        if self.handle:
            self.lib.amt_tcg_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def supported(self, h: int, w: int) -> bool:
        return bool(self.lib.amt_tcg_supported(h, w, self.radius))

    def axis0(self, x, skip_every: int = 0, skip_offset: int = 0):
        """x: (n_img, H, W) uint16 bits -> (n_img, 5, H, W) uint8 digit planes of the axis-0 pass."""
        torch = self._torch
        n_img, h, w = x.shape
        digits = torch.zeros((n_img, 5, h, w), dtype=torch.uint8, device=x.device)
        check(self.lib.amt_tcg_axis0(self.handle, ptr(x), n_img, h, w, ptr(digits), skip_every, skip_offset, stream_ptr()),
              "amt_tcg_axis0")
        return digits

    def axis1(self, digits, lo=None, scale: float = 1.0 / 65535.0, want_buckets: bool = False, skip_every: int = 0,
              skip_offset: int = 0):
        """digits -> (lo - G_hi float64 planes, min/max keys, bucket codes or None)."""
        torch = self._torch
        n_img, _, h, w = digits.shape
        out = torch.zeros((n_img, h, w), dtype=torch.float64, device=digits.device)
        mm = torch.empty((n_img, 2), dtype=torch.int64, device=digits.device)
        buckets = torch.zeros((n_img, h, w), dtype=torch.int16, device=digits.device) if want_buckets else None
        check(self.lib.amt_tcg_axis1(self.handle, ptr(digits), ptr(lo), scale, ptr(out), n_img, h, w, ptr(buckets), ptr(mm),
                                     skip_every, skip_offset, stream_ptr()),
              "amt_tcg_axis1")
        return out, mm, buckets

    def axis1_dog(self, digits, raw, sigma_lo: float, scale: float = 1.0 / 65535.0, want_buckets: bool = False,
                  skip_every: int = 0, skip_offset: int = 0):
        """digits + the raw uint16 planes -> (G_lo(raw) - G_hi float64 planes, min/max keys, bucket codes or None): the
        narrow Gaussian computed inside the tensor-core kernel (``amt_tcg_axis1_dog``, the executor's path)."""
        torch = self._torch
        n_img, _, h, w = digits.shape
        hw = gaussian_half_weights(sigma_lo)
        d_hw = torch.from_numpy(hw).to(digits.device)
        out = torch.zeros((n_img, h, w), dtype=torch.float64, device=digits.device)
        mm = torch.empty((n_img, 2), dtype=torch.int64, device=digits.device)
        buckets = torch.zeros((n_img, h, w), dtype=torch.int16, device=digits.device) if want_buckets else None
        check(self.lib.amt_tcg_axis1_dog(self.handle, ptr(digits), ptr(raw), ptr(d_hw), len(hw) - 1, scale, ptr(out), n_img, h,
                                         w, ptr(buckets), ptr(mm), skip_every, skip_offset, stream_ptr()),
              "amt_tcg_axis1_dog")
        return out, mm, buckets


def gauss_lo2d(x, scale: float, sigma: float, skip_every: int = 0, skip_offset: int = 0):
    """The narrow Gaussian (radius <= 4) of (n_img, H, W) uint16 planes, float64, scipy's order."""
    torch = torch_mod()
    n_img, h, w = x.shape
    hw = gaussian_half_weights(sigma)
    d_hw = torch.from_numpy(hw).to(x.device)
    out = torch.zeros((n_img, h, w), dtype=torch.float64, device=x.device)
    check(_lib.load().amt_gauss_lo2d(ptr(x), scale, ptr(out), n_img, h, w, ptr(d_hw), len(hw) - 1, skip_every, skip_offset,
                                     stream_ptr()),
          "amt_gauss_lo2d")
    return out


def gaussian_nd(x, scale: float, sigma, mode: int = 0):
    """All-axes Gaussian of one N-D array (scipy's axis order 0, 1, ...).  sigma: one value or one per
    axis; mode: 0 = scipy 'nearest', 1 = scipy 'reflect' (``AMT_EXTEND_*``)."""
    torch = torch_mod()
    lib = _lib.load()
    shape = tuple(x.shape)
    sigmas = [float(sigma)] * len(shape) if np.isscalar(sigma) else [float(s) for s in sigma]
    if len(sigmas) != len(shape):
        raise ValueError("sigma must be a scalar or have one entry per axis")
    cur = x
    code = dtype_code(x)
    for axis in range(len(shape)):
        hw = gaussian_half_weights(sigmas[axis])
        d_hw = torch.from_numpy(hw).to(x.device)
        outer = int(np.prod(shape[:axis], dtype=np.int64))
        n = shape[axis]
        inner = int(np.prod(shape[axis + 1:], dtype=np.int64))
        out = torch.empty(shape, dtype=torch.float64, device=x.device)
        check(
            lib.amt_gaussian_axis_mode(ptr(cur), code, scale, ptr(out), outer, n, inner, ptr(d_hw), len(hw) - 1, mode,
                                       stream_ptr()),
            "amt_gaussian_axis_mode",
        )
        cur, code = out, AMT_F64
    return cur


def minmax_filter_nd(x, size: int, is_max: bool, minuend=None):
    """Flat erosion / dilation with a size**ndim box over every axis of one N-D array (uint16 bits
    or float64), scipy's window alignment (grey_erosion / grey_dilation with ``size``),
    mode='reflect'.  ``minuend``: the last pass writes ``minuend - result``."""
    torch = torch_mod()
    lib = _lib.load()
    shape = tuple(x.shape)
    code = dtype_code(x)
    left = size // 2 - (1 if (is_max and size % 2 == 0) else 0)
    cur = x
    for axis in range(len(shape)):
        outer = int(np.prod(shape[:axis], dtype=np.int64))
        n = shape[axis]
        inner = int(np.prod(shape[axis + 1:], dtype=np.int64))
        out = torch.empty(shape, dtype=x.dtype, device=x.device)
        last = axis == len(shape) - 1
        check(
            lib.amt_minmax_filter_axis(ptr(cur), code, ptr(out), ptr(minuend) if (last and minuend is not None) else None,
                                       outer, n, inner, size, left, 1 if is_max else 0, stream_ptr()),
            "amt_minmax_filter_axis",
        )
        cur = out
    return cur


def sub_f64(a, b):
    torch = torch_mod()
    out = torch.empty_like(a)
    check(_lib.load().amt_sub_f64(ptr(a), ptr(b), ptr(out), a.numel(), stream_ptr()), "amt_sub_f64")
    return out


def minmax_keys(x2d):
    """x2d: (n_img, n) -> int64 keys (n_img, 2)."""
    torch = torch_mod()
    lib = _lib.load()
    n_img, n = x2d.shape
    mm = torch.empty((n_img, 2), dtype=torch.int64, device=x2d.device)
    fn = lib.amt_minmax_f64 if dtype_code(x2d) == AMT_F64 else lib.amt_minmax_u16
    check(fn(ptr(x2d), n_img, n, ptr(mm), stream_ptr()), "amt_minmax")
    return mm


def minmax_values(mm, is_f64: bool) -> np.ndarray:
    torch = torch_mod()
    out = torch.empty((mm.shape[0], 2), dtype=torch.float64, device=mm.device)
    check(_lib.load().amt_minmax_decode(ptr(mm), 1 if is_f64 else 0, mm.shape[0], ptr(out), stream_ptr()), "amt_minmax_decode")
    return to_host(out)


def order_statistics(x2d, ranks: list[int], mm=None) -> np.ndarray:
    """Exact order statistics of every plane: (n_img, len(ranks)) float64 on the host."""
    torch = torch_mod()
    lib = _lib.load()
    n_img, n = x2d.shape
    r = (C.c_int64 * len(ranks))(*ranks)
    out = torch.empty((n_img, len(ranks)), dtype=torch.float64, device=x2d.device)
    if dtype_code(x2d) == AMT_F64:
        if mm is None:
            mm = minmax_keys(x2d)
        nbytes = lib.amt_select_f64_scratch_bytes(n_img, n)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=x2d.device)
        check(lib.amt_select_f64(ptr(x2d), n_img, n, r, len(ranks), ptr(mm), ptr(out), ptr(scratch), nbytes, stream_ptr()),
              "amt_select_f64")
    else:
        nbytes = lib.amt_select_u16_scratch_bytes(n_img)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=x2d.device)
        check(lib.amt_select_u16(ptr(x2d), n_img, n, r, len(ranks), ptr(out), ptr(scratch), nbytes, stream_ptr()),
              "amt_select_u16")
    return to_host(out)


def percentiles(x2d, qs: list[float], mm=None) -> np.ndarray:
    """``np.percentile(plane, qs)`` for every plane -> (n_img, len(qs)) float64."""
    n = x2d.shape[1]
    ranks: list[int] = []
    gammas: list[float] = []
    for q in qs:
        lo, hi, g = rank_pair(n, q)
        ranks += [lo, hi]
        gammas.append(g)
    vals = order_statistics(x2d, ranks, mm)
    out = np.empty((x2d.shape[0], len(qs)), dtype=np.float64)
    for i in range(x2d.shape[0]):
        for j, g in enumerate(gammas):
            out[i, j] = np_lerp(vals[i, 2 * j], vals[i, 2 * j + 1], g)
    return out


def apply_map(x2d, params_host: list[_lib.MapParams]):
    """Elementwise map with per-plane parameters -> float64 (n_img, n)."""
    torch = torch_mod()
    n_img, n = x2d.shape
    arr = (_lib.MapParams * n_img)(*params_host)
    raw = np.frombuffer(arr, dtype=np.uint8).copy()
    d_params = torch.from_numpy(raw).to(x2d.device)
    out = torch.empty((n_img, n), dtype=torch.float64, device=x2d.device)
    check(_lib.load().amt_map(ptr(x2d), dtype_code(x2d), ptr(out), n_img, n, ptr(d_params), None, stream_ptr()), "amt_map")
    return out


def otsu_threshold(x2d, mm=None):
    """Per-plane Otsu threshold (device float64 tensor of n_img values) and the min/max keys."""
    torch = torch_mod()
    lib = _lib.load()
    n_img, n = x2d.shape
    if mm is None:
        mm = minmax_keys(x2d)
    thr = torch.empty(n_img, dtype=torch.float64, device=x2d.device)
    if dtype_code(x2d) == AMT_F64:
        hist = torch.empty((n_img, 256), dtype=torch.int32, device=x2d.device)
        check(lib.amt_hist256_f64(ptr(x2d), n_img, n, ptr(mm), ptr(hist), stream_ptr()), "amt_hist256_f64")
        check(lib.amt_otsu(ptr(hist), 1, None, ptr(mm), n_img, ptr(thr), None, 0, stream_ptr()), "amt_otsu")
    else:
        hist = torch.empty((n_img, 65536), dtype=torch.int32, device=x2d.device)
        check(lib.amt_hist_u16(ptr(x2d), n_img, n, ptr(hist), stream_ptr()), "amt_hist_u16")
        nbytes = lib.amt_otsu_scratch_bytes(2, n_img)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=x2d.device)
        check(lib.amt_otsu(ptr(hist), 2, None, ptr(mm), n_img, ptr(thr), ptr(scratch), nbytes, stream_ptr()), "amt_otsu")
    return thr, mm


def plane_histograms(x2d, mm=None, nbins: int = 256):
    """The histogram skimage's histogram-based thresholds start from, per plane, on the device:
    float64 planes -> ``nbins`` uniform bins over [min, max] (np.histogram semantics: ``amt_hist256_f64`` for the
    default 256, ``amt_hist_f64`` otherwise), uint16 planes -> exact per-value counts (``amt_hist_u16``; skimage
    ignores nbins for integer images).  Returns ``[(counts int64, centers)]`` on the host: the scalar scans over
    these bins are the host-side plan step (NumPy, in skimage's own dtypes); every per-pixel pass stays on the GPU."""
    torch = torch_mod()
    lib = _lib.load()
    n_img, n = x2d.shape
    if mm is None:
        mm = minmax_keys(x2d)
    out = []
    if dtype_code(x2d) == AMT_F64:
        mnmx = minmax_values(mm, True)
        all_edges = []
        for i in range(n_img):
            first, last = float(mnmx[i, 0]), float(mnmx[i, 1])
            if first == last:
                first, last = first - 0.5, last + 0.5
            all_edges.append(np.linspace(first, last, nbins + 1, endpoint=True, dtype=np.float64))
        hist = torch.empty((n_img, nbins), dtype=torch.int32, device=x2d.device)
        if nbins == 256:
            check(lib.amt_hist256_f64(ptr(x2d), n_img, n, ptr(mm), ptr(hist), stream_ptr()), "amt_hist256_f64")
        else:
            d_edges = torch.from_numpy(np.stack(all_edges)).to(x2d.device)
            check(lib.amt_hist_f64(ptr(x2d), n_img, n, ptr(d_edges), nbins, ptr(hist), stream_ptr()), "amt_hist_f64")
        h = to_host(hist).astype(np.int64)
        for i in range(n_img):
            edges = all_edges[i]
            out.append((h[i], (edges[:-1] + edges[1:]) / 2.0))
    else:
        hist = torch.empty((n_img, 65536), dtype=torch.int32, device=x2d.device)
        check(lib.amt_hist_u16(ptr(x2d), n_img, n, ptr(hist), stream_ptr()), "amt_hist_u16")
        h = to_host(hist).astype(np.int64)
        mnmx = minmax_values(mm, False)
        for i in range(n_img):
            lo, hi = int(mnmx[i, 0]), int(mnmx[i, 1])
            out.append((h[i, lo : hi + 1], np.arange(lo, hi + 1)))
    return out


def plane_sums_f64(x2d) -> np.ndarray:
    """``np.sum`` of every contiguous float64 plane in NumPy's pairwise order, bit for bit (``amt_pairwise_sum_f64``)."""
    torch = torch_mod()
    lib = _lib.load()
    n_img, n = x2d.shape
    x2d = x2d.contiguous()
    nbytes = lib.amt_pairwise_sum_scratch_bytes(n_img, n)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=x2d.device)
    out = torch.empty(n_img, dtype=torch.float64, device=x2d.device)
    check(lib.amt_pairwise_sum_f64(ptr(x2d), n_img, n, ptr(out), ptr(scratch), nbytes, stream_ptr()), "amt_pairwise_sum_f64")
    return to_host(out)


def li_threshold_f64(plane, lo: float, tolerance=None, initial_guess=None) -> float:
    """``ski.filters.threshold_li`` of one finite, non-constant float64 plane (flattened, on the device): scikit-image's
    iteration on ``image - image.min()`` with both class means taken over the PIXELS in NumPy's pairwise order
    (``amt_li_shift_f64`` / ``amt_li_min_gap_f64`` / ``amt_li_split_f64`` + ``amt_pairwise_sum_f64``); the scalar
    recurrence runs here in NumPy float64 scalars, as it does in scikit-image."""
    torch = torch_mod()
    lib = _lib.load()
    n = int(plane.numel())
    plane = plane.contiguous().reshape(-1)
    shifted = torch.empty(n, dtype=torch.float64, device=plane.device)
    check(lib.amt_li_shift_f64(ptr(plane), n, float(lo), ptr(shifted), stream_ptr()), "amt_li_shift_f64")
    if not tolerance:
        nbytes = lib.amt_li_min_gap_scratch_bytes(n)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=plane.device)
        gap = torch.empty(1, dtype=torch.float64, device=plane.device)
        check(lib.amt_li_min_gap_f64(ptr(shifted), n, ptr(gap), ptr(scratch), nbytes, stream_ptr()), "amt_li_min_gap_f64")
        tolerance = np.float64(to_host(gap)[0]) / 2
        del scratch
    if initial_guess is None:
        t_next = np.float64(plane_sums_f64(shifted.reshape(1, n))[0]) / n  # np.mean
    elif np.isscalar(initial_guess):
        t_next = initial_guess - float(lo)
        top = float(minmax_values(minmax_keys(shifted.reshape(1, n)), True)[0, 1])
        if not 0 < t_next < top:
            raise ValueError("The initial guess for threshold_li must be within the range of the image.")
    elif callable(initial_guess):
        raise NotImplementedError("threshold_li on a float image: a callable initial_guess would run on the host copy of the "
                                  "image; pass its value instead")
    else:
        raise TypeError("Incorrect type for `initial_guess`")
    above = torch.empty(n, dtype=torch.float64, device=plane.device)
    rest = torch.empty(n, dtype=torch.float64, device=plane.device)
    totals = torch.empty(2, dtype=torch.int64, device=plane.device)
    nbytes = lib.amt_li_split_scratch_bytes(n)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=plane.device)

    def mean_of(part, count):  # np.mean of the compacted array: pairwise sum / count (nan for an empty one, as NumPy)
        if count == 0:
            return np.float64(np.nan)
        return np.float64(plane_sums_f64(part[:count].reshape(1, count))[0]) / count

    t_curr = -2 * tolerance
    with np.errstate(all="ignore"):
        while abs(t_next - t_curr) > tolerance:
            t_curr = t_next
            check(lib.amt_li_split_f64(ptr(shifted), n, float(t_curr), ptr(above), ptr(rest), ptr(totals), ptr(scratch), nbytes,
                                       stream_ptr()), "amt_li_split_f64")
            n_above, n_rest = (int(v) for v in to_host(totals))
            mean_fore = mean_of(above, n_above)
            mean_back = mean_of(rest, n_rest)
            if mean_back == 0.0:
                break
            t_next = (mean_back - mean_fore) / (np.log(mean_back) - np.log(mean_fore))
    return float(t_next + np.float64(lo))


def plane_sums_u16(x2d) -> np.ndarray:
    """Exact integer sum of every uint16 plane (from the device histogram) -> int64 per plane."""
    return np.array([int((c * v).sum()) for c, v in plane_histograms(x2d)], dtype=np.int64)


def threshold_gt(x2d, thr):
    torch = torch_mod()
    n_img, n = x2d.shape
    mask = torch.empty((n_img, n), dtype=torch.uint8, device=x2d.device)
    check(_lib.load().amt_threshold_gt(ptr(x2d), dtype_code(x2d), n_img, n, ptr(thr), ptr(mask), stream_ptr()),
          "amt_threshold_gt")
    return mask


def label(x, kind: int, clear_border: bool, thresholds=None, max_value: int = 0):
    """x: (n_img, H, W) uint8 mask (kind 0), float64 plane (kind 1) or int32 labels (kind 2)
    -> (labels int32 (n_img, H, W), counts int32 (n_img,))."""
    torch = torch_mod()
    lib = _lib.load()
    n_img, h, w = x.shape
    labels = torch.empty((n_img, h, w), dtype=torch.int32, device=x.device)
    counts = torch.empty(n_img, dtype=torch.int32, device=x.device)
    nbytes = lib.amt_label_scratch_bytes(n_img, h, w, max_value)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    check(
        lib.amt_label(ptr(x), kind, ptr(thresholds), max_value, n_img, h, w, 1 if clear_border else 0, ptr(labels),
                      ptr(counts), ptr(scratch), nbytes, stream_ptr()),
        "amt_label",
    )
    return labels, counts


def region_table(labels, counts, channels, max_labels: int, with_shape: bool = False):
    """labels (n_img, H, W) int32, channels (n_img, C, H, W) uint16-as-int16 or None ->
    float64 table (n_img, cols, max_labels) on the device."""
    torch = torch_mod()
    lib = _lib.load()
    n_img, h, w = labels.shape
    n_ch = 0 if channels is None else channels.shape[1]
    acc = torch.empty((n_img, _lib.acc_fields(n_ch), max_labels), dtype=torch.int64, device=labels.device)
    table = torch.empty((n_img, _lib.table_cols(n_ch), max_labels), dtype=torch.float64, device=labels.device)
    check(
        lib.amt_region_reduce(ptr(labels), ptr(channels), n_ch, n_ch * h * w, h * w, n_img, h, w, max_labels, ptr(acc),
                              stream_ptr()),
        "amt_region_reduce",
    )
    check(lib.amt_region_finalize(ptr(acc), ptr(counts), n_ch, n_img, max_labels, ptr(table), stream_ptr()),
          "amt_region_finalize")
    if with_shape:
        nbytes = lib.amt_region_shape_scratch_bytes(n_img, h, w, max_labels)
        scratch = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=labels.device)
        check(
            lib.amt_region_shape(ptr(labels), ptr(acc), n_ch, ptr(counts), n_img, h, w, max_labels, ptr(table),
                                 ptr(scratch), nbytes, stream_ptr()),
            "amt_region_shape",
        )
    return table, acc


def outline_square_keys(labels2d) -> np.ndarray:
    """Sorted marching-squares keys of one (H, W) int32 device label image (``amt_outline_squares``):
    ``label<<34 | r0<<19 | c0<<4 | case``, i.e. per label the non-trivial squares in raster order."""
    torch = torch_mod()
    lib = _lib.load()
    h, w = labels2d.shape
    count = torch.zeros(1, dtype=torch.int64, device=labels2d.device)
    capacity = max(4096, (h * w) // 8)
    while True:
        keys = torch.empty(capacity, dtype=torch.int64, device=labels2d.device)
        check(lib.amt_outline_squares(ptr(labels2d), h, w, ptr(keys), capacity, ptr(count), stream_ptr()),
              "amt_outline_squares")
        n = int(count.item())
        if n <= capacity:
            break
        capacity = n
    return np.sort(to_host(keys[:n]).view(np.uint64))


def outline_borders(labels2d, n_labels: int, min_points: int = 5):
    """OpenCV-ordered outer border of every label 1..n_labels of one (H, W) int32 device label image
    (``amt_outline_trace_find`` / ``_write``): -> (points int32 (N, 2) as (y, x), offsets int64
    (n_labels + 1)); labels whose longest border has fewer than ``min_points`` points get an empty range."""
    torch = torch_mod()
    lib = _lib.load()
    h, w = labels2d.shape
    best = torch.empty(max(n_labels, 1), dtype=torch.int64, device=labels2d.device)
    check(lib.amt_outline_trace_find(ptr(labels2d), h, w, max(n_labels, 1), ptr(best), stream_ptr()),
          "amt_outline_trace_find")
    lengths = (to_host(best).view(np.uint64) >> np.uint64(32)).astype(np.int64)[:n_labels]
    lengths[lengths < min_points] = 0
    offsets = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)
    total = int(offsets[-1])
    points = torch.empty((max(total, 1), 2), dtype=torch.int32, device=labels2d.device)
    if n_labels > 0 and total > 0:
        d_off = torch.from_numpy(offsets).to(labels2d.device)
        check(lib.amt_outline_trace_write(ptr(labels2d), h, w, n_labels, ptr(best), ptr(d_off), ptr(points), stream_ptr()),
              "amt_outline_trace_write")
    return to_host(points)[:total], offsets


def window_threshold_u16(x3d, window: tuple[int, int], kind: int, k: float, r: float, want_thresholds: bool = False):
    """niblack (kind 0) / sauvola (kind 1) of (n_img, H, W) uint16 planes -> uint8 mask (and the float64
    threshold image when asked for)."""
    torch = torch_mod()
    n_img, h, w = x3d.shape
    mask = torch.empty((n_img, h, w), dtype=torch.uint8, device=x3d.device)
    thr = torch.empty((n_img, h, w), dtype=torch.float64, device=x3d.device) if want_thresholds else None
    check(
        _lib.load().amt_window_threshold_u16(ptr(x3d), n_img, h, w, int(window[0]), int(window[1]), kind, float(k),
                                             float(r), ptr(mask), ptr(thr), stream_ptr()),
        "amt_window_threshold_u16",
    )
    return mask, thr


def window_threshold_f64(x3d, window: tuple[int, int], kind: int, k: float, r: float, want_thresholds: bool = False):
    """niblack (kind 0) / sauvola (kind 1) of (n_img, H, W) float64 planes, scikit-image's float route operation by
    operation -> uint8 mask (and the float64 threshold image when asked for)."""
    torch = torch_mod()
    lib = _lib.load()
    n_img, h, w = x3d.shape
    x3d = x3d.contiguous()
    mask = torch.empty((n_img, h, w), dtype=torch.uint8, device=x3d.device)
    thr = torch.empty((n_img, h, w), dtype=torch.float64, device=x3d.device) if want_thresholds else None
    nbytes = lib.amt_window_threshold_f64_scratch_bytes(n_img, h, w, int(window[0]), int(window[1]))
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=x3d.device)
    check(
        lib.amt_window_threshold_f64(ptr(x3d), n_img, h, w, int(window[0]), int(window[1]), kind, float(k), float(r),
                                     ptr(mask), ptr(thr), ptr(scratch), nbytes, stream_ptr()),
        "amt_window_threshold_f64",
    )
    return mask, thr


def threshold_gt_image(x, thr, offset: float):
    """mask = x > thr - offset, elementwise (x uint16 or float64, thr float64, same shape)."""
    torch = torch_mod()
    mask = torch.empty(tuple(x.shape), dtype=torch.uint8, device=x.device)
    check(_lib.load().amt_threshold_gt_image(ptr(x), dtype_code(x), x.numel(), ptr(thr), float(offset), ptr(mask),
                                             stream_ptr()), "amt_threshold_gt_image")
    return mask


class PinnedBuffer:
    """A page-locked host array from ``amt_host_alloc`` (no torch involved): ``.array`` is a NumPy view of the
    pinned pages.  ``write_combined=True`` suits buffers the host only fills (decoded frames on their way to the
    device); reading them back on the CPU is slow.  Freed by ``close()`` / the context manager."""

    def __init__(self, shape, dtype, write_combined: bool = False) -> None:
        import ctypes as C

        self._lib = _lib.load()
        dt = np.dtype(dtype)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
        self._ptr = C.c_void_p()
        check(self._lib.amt_host_alloc(nbytes, 1 if write_combined else 0, C.byref(self._ptr)), "amt_host_alloc")
        raw = (C.c_uint8 * nbytes).from_address(self._ptr.value)
        self.array = np.frombuffer(raw, dtype=dt).reshape(shape)

    def close(self) -> None:
        if self._ptr is not None and self._ptr.value:
            self.array = None
            check(self._lib.amt_host_free(self._ptr), "amt_host_free")
            self._ptr = None

    def __enter__(self) -> "PinnedBuffer":
        return self

    def __exit__(self, *exc) -> None:
        self.close()
