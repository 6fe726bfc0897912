"""Numerics probe for the tensor-core route named in DESIGN.md ("the way past the exact order"): the 1-D
Gaussian as a sum of small-integer products, the form tcgen05 `kind::i8` MMAs accumulate EXACTLY in int32.

  pass 1 (uint16 samples):  x = 256*xh + xl, two uint8 slices, exact
  weights:                  w_j ~ sum_s ws[s][j] * 2^-(E + 7*(s+1)),  ws in [-64, 64] (signed base-128 digits)
  pass 2 (float64 input G): G ~ sum_t g[t] * 2^-(8*(t+1)),            g in [0, 255]   (G in [0, 1))

Every inner sum  sum_j ws[s][j] * slice[j]  over the 129 taps is an integer below 2^23 in magnitude, so an
int32 accumulator holds it exactly; the only error is the truncation of the weight and G expansions.  This
script (CPU, NumPy, no GPU) measures that error against the exact-order oracle on the reference fixture for
several slice counts, and counts the int8 products per sample.  It is an experiment for the next round's
kernel, not part of the product or the tests.

    python scripts/sliced_toeplitz_probe.py
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402
from oracle import filters  # noqa: E402


def half_weights(sigma: float) -> np.ndarray:
    r = int(4.0 * sigma + 0.5)
    x = np.arange(-r, r + 1)
    w = np.exp(-0.5 / (sigma * sigma) * x**2)
    return w / w.sum()


def weight_slices(w: np.ndarray, n_slices: int):
    """Signed base-128 digits of round(w * 2^(E + 7*n_slices)), most significant first."""
    e = int(np.floor(-np.log2(w.max()))) - 1  # w.max() * 2^e in [0.25, 0.5): the leading balanced digit stays below 64
    scale_bits = e + 7 * n_slices
    big = np.array([int(round(float(v) * 2.0**scale_bits)) for v in w], dtype=object)
    digits = []
    for _ in range(n_slices):  # least significant first, balanced digits in [-64, 63]
        d = np.array([((int(b) + 64) % 128) - 64 for b in big], dtype=np.int64)
        digits.append(d)
        big = np.array([(int(b) - int(dd)) // 128 for b, dd in zip(big, d)], dtype=object)
    assert all(int(b) == 0 for b in big), "leading digit overflow"
    digits.reverse()
    shifts = [e + 7 * (s + 1) for s in range(n_slices)]
    return digits, shifts


def correlate_int(slice_u8: np.ndarray, digits: np.ndarray, axis: int) -> np.ndarray:
    """sum_j digits[j] * slice[i + j - r] along `axis`, edge-clamped, exact in int64 (checked to fit int32)."""
    r = (len(digits) - 1) // 2
    pad = [(0, 0)] * slice_u8.ndim
    pad[axis] = (r, r)
    p = np.pad(slice_u8.astype(np.int64), pad, mode="edge")
    out = np.zeros(slice_u8.shape, dtype=np.int64)
    n = slice_u8.shape[axis]
    for j, d in enumerate(digits):
        if d:
            out += int(d) * np.take(p, range(j, j + n), axis=axis)
    assert np.abs(out).max() < 2**31
    return out


def gaussian_pass1(u16: np.ndarray, w: np.ndarray, n_w: int, axis: int) -> tuple[np.ndarray, int]:
    digits, shifts = weight_slices(w, n_w)
    xh, xl = (u16 >> 8).astype(np.uint8), (u16 & 255).astype(np.uint8)
    acc = np.zeros(u16.shape, dtype=np.float64)
    products = 0
    for d, sh in zip(digits, shifts):
        part = 256 * correlate_int(xh, d, axis) + correlate_int(xl, d, axis)  # exact integers
        acc += part.astype(np.float64) * 2.0**-sh
        products += 2
    return acc * (1.0 / 65535.0), products


def gaussian_pass2(g: np.ndarray, w: np.ndarray, n_w: int, n_g: int, axis: int, keep_bits: int) -> tuple[np.ndarray, int]:
    digits, shifts = weight_slices(w, n_w)
    fixed = np.floor(g * 2.0 ** (8 * n_g)).astype(np.int64)  # G in [0, 1): n_g uint8 slices, truncated
    slices = [((fixed >> (8 * (n_g - 1 - t))) & 255).astype(np.uint8) for t in range(n_g)]
    acc = np.zeros(g.shape, dtype=np.float64)
    products = 0
    for t, sl in enumerate(slices):
        for d, sh in zip(digits, shifts):
            bits = 8 * (t + 1) + sh
            if bits - 15 > keep_bits:  # the whole term is below 2^-keep_bits
                continue
            acc += correlate_int(sl, d, axis).astype(np.float64) * 2.0**-bits
            products += 1
    return acc, products


def main() -> None:
    with np.load(ROOT / "tests" / "golden" / "config1_multichannel.npz") as z:
        fov = z["fov"]
    x = fov[2]  # FITC: a channel that is not thresholded in workload W
    lo_w, hi_w = half_weights(0.6), half_weights(16.0)
    exact = filters.difference_of_gaussians(x, 0.6, 16.0)
    p_exact = oracle.rescale_by_percentile(oracle.subtract_background_dog(x, 0.6, 16.0, percentile=0), (1, 99))
    print("slices (weights, G) | int8 products / sample (pass1 + pass2, hi filter) | max |dDoG| | max |dP|")
    for n_w, n_g in ((3, 3), (4, 4), (5, 5), (6, 6)):
        keep = 7 * n_w + 3
        g_hi, p1 = gaussian_pass1(x, hi_w, n_w, axis=0)
        g_hi2, p2 = gaussian_pass2(g_hi, hi_w, n_w, n_g, axis=1, keep_bits=keep)
        g_lo, _ = gaussian_pass1(x, lo_w, n_w, axis=0)
        g_lo2, _ = gaussian_pass2(g_lo, lo_w, n_w, n_g, axis=1, keep_bits=keep)
        dog = g_lo2 - g_hi2
        level = np.percentile(dog, 0)
        p = oracle.rescale_by_percentile(np.clip(dog - level, 0, None), (1, 99))
        taps = len(hi_w)
        print(f"  ({n_w}, {n_g})           | {p1 * taps} + {p2 * taps} = {(p1 + p2) * taps:6d}"
              f"                         | {np.abs(dog - exact).max():.2e} | {np.abs(p - p_exact).max():.2e}")


if __name__ == "__main__":
    main()
