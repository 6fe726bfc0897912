"""CPU pin of `oracle/tcgauss.py` (the integer pipeline of the tensor-core Gaussian, ref call site
src/arcadia_microscopy_tools/operations.py:91): with integer weights built the way `amt_tcg_create` builds them
(csrc/tcgauss.cu: round(w * 2^S), nudged so that they sum to 2^S exactly), the restatement stays within the stated
tolerance of the REAL scipy.ndimage.gaussian_filter, and the digit products it leaves out are as small as
`amt_tcg_error_bound` says.  The GPU tests then compare the product with this restatement bit for bit."""

import math

import numpy as np
import pytest
from scipy import ndimage as ndi

from oracle import tcgauss

TOL_G = 2e-10  # the tolerance tests/test_gpu_tcgauss.py uses against scipy


def integer_weights(sigma: float):
    """-> (W[0..r] uint64, S, float half weights): the weights the library reports through amt_tcg_weights."""
    r = int(4.0 * sigma + 0.5)
    x = np.arange(-r, r + 1, dtype=np.float64)
    w = np.exp(-0.5 * (x / sigma) ** 2)
    w /= w.sum()
    hw = w[r:]
    s_bits = 0
    while s_bits < 40 and math.ldexp(hw[0], s_bits + 1) < 4294967000.0:
        s_bits += 1
    big, frac, total = [], [], 0
    for t in range(r + 1):
        v = math.ldexp(hw[t], s_bits)
        f = math.floor(v + 0.5)
        big.append(int(f))
        frac.append(v - f)
        total += (1 if t == 0 else 2) * int(f)
    defect = (1 << s_bits) - total
    if defect % 2:
        step = 1 if defect > 0 else -1
        big[0] += step
        defect -= step
    while defect:
        step = 1 if defect > 0 else -1
        best = max(range(1, r + 1), key=lambda t: step * frac[t])
        big[best] += step
        frac[best] -= step
        defect -= 2 * step
    return np.array(big, dtype=np.uint64), s_bits, hw


@pytest.mark.parametrize("sigma", [16.0, 5.0, 1.5])
def test_weights_sum_to_a_power_of_two_and_fit_32_bits(sigma):
    big, s_bits, _ = integer_weights(sigma)
    assert int(big[0]) + 2 * int(big[1:].sum()) == 1 << s_bits
    assert int(big.max()) < 1 << 32


@pytest.mark.parametrize("sigma,shape", [(16.0, (200, 272)), (16.0, (130, 144)), (5.0, (150, 160))])
def test_integer_pipeline_against_scipy(sigma, shape):
    big, s_bits, _ = integer_weights(sigma)
    rng = np.random.default_rng(int(sigma * 10) + shape[0])
    x = rng.integers(0, 65536, size=shape).astype(np.uint16)
    g1 = tcgauss.axis0_int(x, big, s_bits)
    assert int(g1.max()) < 1 << 40  # five base-256 digits
    got = tcgauss.axis1_float(g1, big, s_bits, 1.0 / 65535.0)
    want = ndi.gaussian_filter(x.astype(np.float64) * (1.0 / 65535.0), sigma, mode="nearest", truncate=4.0)
    assert np.max(np.abs(got - want)) <= TOL_G
    # a constant image filters to itself up to the dropped digit products (the weights sum to 2^S exactly)
    flat = np.full(shape, 40000, dtype=np.uint16)
    g = tcgauss.axis1_float(tcgauss.axis0_int(flat, big, s_bits), big, s_bits, 1.0 / 65535.0)
    assert np.max(np.abs(g - 40000.0 / 65535.0)) <= 2e-11


def test_dropped_digit_products_are_what_the_error_bound_counts():
    """pass 2 leaves out (weight digit d) x (sample digit s) for d + s < JMIN; csrc/tcgauss.cu bounds their sum by
    sum 256^(d+s) * 255 * |digit d of W|_1 * 2^-(S+24) / 65535.  Measured here: the left-out products themselves."""
    big, s_bits, _ = integer_weights(16.0)
    rng = np.random.default_rng(3)
    x = rng.integers(0, 65536, size=(140, 160)).astype(np.uint16)
    g1 = tcgauss.axis0_int(x, big, s_bits)
    r = len(big) - 1
    k = tcgauss.full_kernel(big)
    w = x.shape[1]
    dropped = np.zeros(x.shape, dtype=np.uint64)  # < 2^41: exact
    bound = 0.0
    for d in range(4):
        wd = (k >> np.uint64(8 * d)) & np.uint64(0xFF)
        l1 = int(wd.sum())
        for s in range(tcgauss.GD):
            if d + s >= tcgauss.JMIN:
                continue
            gs = np.pad((g1 >> np.uint64(8 * s)) & np.uint64(0xFF), ((0, 0), (r, r)))
            acc = np.zeros(x.shape, dtype=np.uint64)
            for i in range(2 * r + 1):
                if wd[i]:
                    acc += wd[i] * gs[:, i : i + w]
            dropped += acc << np.uint64(8 * (d + s))
            bound += 255.0 * l1 * 2.0 ** (8 * (d + s) - (s_bits + 24)) / 65535.0
    worst = float(dropped.max()) * 2.0 ** -(s_bits + 24) / 65535.0
    assert 0.0 < worst <= bound
    assert bound < 6e-12  # 5.2e-12 for sigma = 16: part of amt_tcg_error_bound (5.3e-10)
