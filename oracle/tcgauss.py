"""CPU restatement of the integer pipeline of the tensor-core Gaussian (test infrastructure only).

The product (`csrc/tcgauss.cu`) evaluates the sigma_high Gaussian of
``ski.filters.difference_of_gaussians`` (ref: src/arcadia_microscopy_tools/operations.py:91) for the
channels nothing discrete is derived from with integer weights ``W[t] ~ w[t] * 2**S`` on the tensor
cores.  Everything in that path is exact integer arithmetic plus ONE conversion to float64, so this
file restates it bit for bit from the weights the library reports (``amt_tcg_weights``); the tests
then compare (a) the product with this restatement exactly and (b) this restatement with the real
``scipy.ndimage.gaussian_filter`` within the stated tolerance.  Only ``tests/`` import it.
"""

from __future__ import annotations

import numpy as np

GD = 5      # digits (base 256) of the axis-0 result
JMIN = 3    # digit products of significance below 256**JMIN are dropped in the axis-1 pass


def full_kernel(w_half: np.ndarray) -> np.ndarray:
    w_half = np.asarray(w_half, dtype=np.uint64)
    return np.concatenate([w_half[:0:-1], w_half])


def axis0_int(x: np.ndarray, w_half: np.ndarray, scale_bits: int) -> np.ndarray:
    """x (H, W) uint16 -> G1q (H, W) uint64: round(sum_t W[t] x[clamp(y+t)] / 2**(S-24)), 40 bits."""
    r = len(w_half) - 1
    k = full_kernel(w_half)
    xp = np.pad(x.astype(np.uint64), ((r, r), (0, 0)), mode="edge")
    h = x.shape[0]
    tot = np.zeros(x.shape, dtype=np.uint64)
    for i in range(2 * r + 1):
        tot += k[i] * xp[i : i + h]
    shift = np.uint64(scale_bits - 24)
    return (tot + (np.uint64(1) << (shift - np.uint64(1)))) >> shift


def digits_of(g: np.ndarray) -> np.ndarray:
    return np.stack([((g >> np.uint64(8 * s)) & np.uint64(0xFF)).astype(np.uint8) for s in range(GD)])


def axis1_float(g1q: np.ndarray, w_half: np.ndarray, scale_bits: int, in_scale: float) -> np.ndarray:
    """G1q (H, W) -> G_hi float64, exactly as the kernel's epilogue: zero-extended digit products summed in
    integers (d + s >= JMIN), one conversion, the clamped-edge taps added in float64, one multiply."""
    r = len(w_half) - 1
    h, w = g1q.shape
    k = full_kernel(w_half)
    wd = [((k >> np.uint64(8 * d)) & np.uint64(0xFF)) for d in range(4)]
    gs = [np.pad((g1q >> np.uint64(8 * s)) & np.uint64(0xFF), ((0, 0), (r, r))) for s in range(GD)]
    tot = np.zeros((h, w), dtype=np.uint64)
    for d in range(4):
        for s in range(GD):
            j = d + s
            if j < JMIN:
                continue
            acc = np.zeros((h, w), dtype=np.uint64)
            for i in range(2 * r + 1):
                if wd[d][i]:
                    acc += wd[d][i] * gs[s][:, i : i + w]
            tot += acc << np.uint64(8 * (j - JMIN))
    g = tot.astype(np.float64)
    suffix = np.zeros(r + 2, dtype=np.uint64)
    for j in range(r, -1, -1):
        suffix[j] = suffix[j + 1] + np.uint64(w_half[j])
    suffix_f = np.ldexp(suffix.astype(np.float64), -8 * JMIN)
    xs = np.arange(w)
    f_l = np.where(xs < r, suffix_f[np.minimum(xs + 1, r + 1)], 0.0)
    f_r = np.where(xs >= w - r, suffix_f[np.clip(w - xs, 0, r + 1)], 0.0)
    e_l = g1q[:, :1].astype(np.float64)
    e_r = g1q[:, -1:].astype(np.float64)
    edge = (f_l[None, :] * e_l + f_r[None, :] * e_r)
    # the kernel adds the edge term only in tiles that touch an edge; adding 0.0 elsewhere is the same number
    g = g + edge
    return g * np.ldexp(np.float64(in_scale), -(scale_bits + 24 - 8 * JMIN))
