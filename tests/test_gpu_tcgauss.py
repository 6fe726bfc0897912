"""GPU parity of the tensor-core Gaussian (csrc/tcgauss.cu, ref: operations.py:91).

Three layers: (1) the integer weights the library reports against scipy's float64 weights, (2) the
kernels against the integer restatement in oracle/tcgauss.py — bit for bit, digits and float64
planes — and (3) the result against the real scipy.ndimage.gaussian_filter within TOL (absolute, on
the [0, 1] scale of img_as_float; the reference's tolerance for filtered planes is 1e-5 relative)."""

from __future__ import annotations

import numpy as np
import pytest
from scipy import ndimage as ndi

from oracle import filters, tcgauss

pytestmark = pytest.mark.gpu

from arcadia_microscopy_tools_b200 import _gpu  # noqa: E402

TOL_G = 2e-10   # |G_hi - scipy| on the [0, 1] scale (bound: 129 * 2^-37 * 2 passes + 2^-41 = 1.9e-9 at full contrast)
SCALE = 1.0 / 65535.0


def _image(seed, shape, kind):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 65536, size=shape).astype(np.uint16)
    if kind == "dim":
        return rng.poisson(300, size=shape).astype(np.uint16)
    yy, xx = np.mgrid[: shape[0], : shape[1]]
    img = 400 + 9000 * np.exp(-((yy - shape[0] / 3) ** 2 + (xx - shape[1] / 2) ** 2) / 900.0)
    return (img + rng.poisson(50, size=shape)).astype(np.uint16)


def test_integer_weights_match_scipy():
    tcg = _gpu.TensorCoreGaussian(16.0)
    w, s = tcg.int_weights.astype(np.float64), tcg.scale_bits
    assert s == 37 and tcg.radius == 64
    full = np.concatenate([w[:0:-1], w])
    assert int(full.sum()) == 2**s
    assert np.max(np.abs(np.ldexp(w, -s) - tcg.hw)) <= 2.0**-s
    assert int(tcg.int_weights.max()) < 2**32


@pytest.mark.parametrize("shape,kind", [((256, 256), "noise"), ((384, 272), "blob"), ((130, 144), "dim"),
                                        ((512, 1024), "noise")])
def test_axis0_digits_bit_exact(shape, kind):
    tcg = _gpu.TensorCoreGaussian(16.0)
    imgs = np.stack([_image(5 + i, shape, kind) for i in range(2)])
    digits = _gpu.to_host(tcg.axis0(_gpu.to_device(imgs)))
    for i in range(2):
        want = tcgauss.digits_of(tcgauss.axis0_int(imgs[i], tcg.int_weights, tcg.scale_bits))
        bad = np.argwhere(digits[i] != want)
        assert bad.size == 0, (shape, kind, i, len(bad), bad[:5])


@pytest.mark.parametrize("shape,kind", [((256, 256), "noise"), ((384, 272), "blob"), ((130, 144), "dim"),
                                        ((512, 1024), "blob"), ((128, 128), "noise"), ((250, 400), "blob"),
                                        ((1100, 144), "dim")])
def test_axis1_bit_exact_and_close_to_scipy(shape, kind):
    tcg = _gpu.TensorCoreGaussian(16.0)
    imgs = np.stack([_image(9 + i, shape, kind) for i in range(2)])
    dev = _gpu.to_device(imgs)
    digits = tcg.axis0(dev)
    ghi, mm, buckets = tcg.axis1(digits, None, SCALE, want_buckets=True)
    ghi = _gpu.to_host(ghi)
    mnmx = _gpu.minmax_values(mm, True)
    for i in range(2):
        g1q = tcgauss.axis0_int(imgs[i], tcg.int_weights, tcg.scale_bits)
        want = tcgauss.axis1_float(g1q, tcg.int_weights, tcg.scale_bits, SCALE)
        assert np.array_equal(ghi[i], want), (shape, kind, i, np.abs(ghi[i] - want).max())
        ref = ndi.gaussian_filter(imgs[i] * SCALE, 16.0, mode="nearest", truncate=4.0)
        err = np.abs(ghi[i] - ref).max()
        assert err <= TOL_G, (shape, kind, err)
        assert mnmx[i][0] == ghi[i].min() and mnmx[i][1] == ghi[i].max()


def test_dog_planes_against_oracle():
    """lo2d bit-exact; DoG = lo - G_hi within TOL_G of the oracle's (= scipy's) DoG."""
    tcg = _gpu.TensorCoreGaussian(16.0)
    imgs = np.stack([_image(21 + i, (256, 384), "blob") for i in range(3)])
    dev = _gpu.to_device(imgs)
    lo = _gpu.gauss_lo2d(dev, SCALE, 0.6)
    lo_h = _gpu.to_host(lo)
    for i in range(3):
        assert np.array_equal(lo_h[i], filters.gaussian(imgs[i], 0.6))
    dog, mm, _ = tcg.axis1(tcg.axis0(dev), lo, SCALE)
    dog = _gpu.to_host(dog)
    for i in range(3):
        want = filters.difference_of_gaussians(imgs[i], 0.6, 16.0)
        assert np.abs(dog[i] - want).max() <= TOL_G


@pytest.mark.parametrize("shape,kind", [((256, 256), "noise"), ((384, 272), "blob"), ((130, 144), "dim"),
                                        ((250, 400), "noise"), ((1100, 144), "dim"), ((128, 2048), "noise")])
@pytest.mark.parametrize("sigma_lo", [0.6, 0.3, 1.0])
def test_fused_narrow_gaussian_is_bit_identical_to_the_two_kernel_route(shape, kind, sigma_lo):
    """amt_tcg_axis1_dog (the executor's path: the narrow Gaussian computed inside the tensor-core kernel from a
    TMA-staged raw tile) against amt_gauss_lo2d + amt_tcg_axis1: planes, bucket codes and min / max bit for bit, for
    radii 1, 2 and 4, shapes with ragged tile edges, and the image border in every direction; the narrow Gaussian
    itself is pinned to scipy by test_dog_planes_against_oracle."""
    tcg = _gpu.TensorCoreGaussian(16.0)
    imgs = np.stack([_image(51 + i, shape, kind) for i in range(3)])
    dev = _gpu.to_device(imgs)
    digits = tcg.axis0(dev)
    lo = _gpu.gauss_lo2d(dev, SCALE, sigma_lo)
    want, mm_w, bk_w = tcg.axis1(digits, lo, SCALE, want_buckets=True)
    got, mm_g, bk_g = tcg.axis1_dog(digits, dev, sigma_lo, SCALE, want_buckets=True)
    want, got = _gpu.to_host(want), _gpu.to_host(got)
    bad = np.argwhere(want.view(np.uint64) != got.view(np.uint64))
    assert bad.size == 0, (shape, kind, sigma_lo, len(bad), bad[:6])
    assert np.array_equal(_gpu.to_host(bk_w), _gpu.to_host(bk_g))
    assert np.array_equal(_gpu.to_host(mm_w), _gpu.to_host(mm_g))
    ref = filters.difference_of_gaussians(imgs[0], sigma_lo, 16.0)
    assert np.abs(got[0] - ref).max() <= TOL_G


def test_plane_skipping():
    """skip_every / skip_offset leave the named planes untouched (the executor's segmentation channel)."""
    tcg = _gpu.TensorCoreGaussian(16.0)
    imgs = np.stack([_image(31 + i, (128, 128), "noise") for i in range(8)])
    dev = _gpu.to_device(imgs)
    digits = tcg.axis0(dev, 4, 1)
    ghi, _, _ = tcg.axis1(digits, None, SCALE, skip_every=4, skip_offset=1)
    ghi = _gpu.to_host(ghi)
    full, _, _ = tcg.axis1(tcg.axis0(dev), None, SCALE)
    full = _gpu.to_host(full)
    for i in range(8):
        if i % 4 == 1:
            assert not ghi[i].any()
        else:
            assert np.array_equal(ghi[i], full[i])


def test_full_size_plane_statistics():
    """2048 x 2048 (BASELINE config 2's plane): error against scipy and exactness against the restatement on a
    subsample of rows (the full restatement takes minutes on the CPU)."""
    tcg = _gpu.TensorCoreGaussian(16.0)
    img = _image(41, (2048, 2048), "blob")
    dev = _gpu.to_device(img[None])
    ghi, _, _ = tcg.axis1(tcg.axis0(dev), None, SCALE)
    ghi = _gpu.to_host(ghi)[0]
    ref = ndi.gaussian_filter(img * SCALE, 16.0, mode="nearest", truncate=4.0)
    assert np.abs(ghi - ref).max() <= TOL_G
