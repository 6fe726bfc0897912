"""Independent pins for the oracle legs DESIGN.md lists as restated from scikit-image's published
code (scikit-image itself is not installable here): OpenCV 4.13 ships its own implementations of the
same definitions, so agreement with it is evidence that the restatement is the textbook algorithm
and not a private variant.  CPU only; skipped when cv2 is absent.

* Otsu on integer images: cv2.threshold(..., THRESH_OTSU) maximises the same between-class variance
  over the exact 8-bit histogram; skimage returns the bin value t with foreground = (x > t), OpenCV the
  threshold with the same convention.
* 8-connected labelling, area, bbox, centroid: cv2.connectedComponentsWithStats.
* second-order central moments (the input of inertia tensor / axis lengths / eccentricity /
  orientation): cv2.moments (x = column, y = row, so OpenCV's mu20 is skimage's mu[0, 2]).
"""

from __future__ import annotations

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

import oracle  # noqa: E402
from oracle import labeling, regionprops, threshold  # noqa: E402


@pytest.mark.parametrize("seed", range(8))
def test_otsu_uint8_matches_opencv(seed):
    rng = np.random.default_rng(seed)
    kind = seed % 4
    if kind == 0:
        img = rng.gamma(2.0, 30.0, (200, 300))
    elif kind == 1:
        img = np.where(rng.random((180, 240)) < 0.3, rng.normal(180, 20, (180, 240)), rng.normal(60, 15, (180, 240)))
    elif kind == 2:
        img = rng.integers(0, 256, (128, 128)).astype(np.float64)
    else:
        img = np.where(rng.random((150, 150)) < 0.05, rng.normal(220, 10, (150, 150)), rng.normal(20, 5, (150, 150)))
    img8 = img.clip(0, 255).astype(np.uint8)
    t_cv, _ = cv2.threshold(img8, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
    assert float(threshold.threshold_otsu(img8)) == float(t_cv)
    assert np.array_equal(oracle.apply_threshold(img8), img8 > t_cv)


def _blobs(seed, shape, n):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[: shape[0], : shape[1]]
    m = np.zeros(shape, bool)
    for _ in range(n):
        cy, cx = rng.uniform(0, shape[0]), rng.uniform(0, shape[1])
        ry, rx = rng.uniform(2, 9, 2)
        m |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1
    m ^= rng.random(shape) < 0.01  # specks and pinholes: diagonal-only contacts exercise 8-connectivity
    return m


@pytest.mark.parametrize("seed,shape,n", [(1, (120, 160), 40), (2, (64, 257), 25), (3, (200, 200), 120)])
def test_labelling_area_bbox_centroid_match_opencv(seed, shape, n):
    mask = _blobs(seed, shape, n)
    n_cv, lab_cv, stats, cent = cv2.connectedComponentsWithStats(mask.astype(np.uint8), connectivity=8)
    lab = labeling.label(mask)
    assert int(lab.max()) == n_cv - 1
    # same partition: every oracle component maps onto exactly one OpenCV component and vice versa
    pairs = np.unique(np.stack([lab[mask], lab_cv[mask]], axis=1), axis=0)
    assert len(pairs) == n_cv - 1 and len(np.unique(pairs[:, 0])) == len(pairs) == len(np.unique(pairs[:, 1]))
    to_cv = dict(pairs.tolist())
    props = regionprops.regionprops_table(lab, properties=("label", "area", "bbox", "centroid"))
    for i, label in enumerate(props["label"]):
        k = to_cv[int(label)]
        x, y, w, h, area = stats[k]
        assert props["area"][i] == area
        assert (props["bbox-0"][i], props["bbox-1"][i], props["bbox-2"][i], props["bbox-3"][i]) == (y, x, y + h, x + w)
        assert np.allclose([props["centroid-0"][i], props["centroid-1"][i]], [cent[k][1], cent[k][0]], rtol=0, atol=1e-9)


def test_central_moments_and_axes_match_opencv():
    mask = _blobs(7, (160, 160), 30)
    lab = labeling.label(mask)
    props = regionprops.regionprops_table(
        lab, properties=("label", "area", "axis_major_length", "axis_minor_length", "eccentricity", "orientation"))
    for i, label in enumerate(props["label"]):
        crop = (lab == label).astype(np.uint8)
        m = cv2.moments(crop, binaryImage=True)
        if m["m00"] < 6:
            continue
        # skimage: inertia tensor [[mu02, -mu11], [-mu11, mu20]] / mu00 in (row, col) order, i.e. OpenCV's
        # [[mu20, -mu11], [-mu11, mu02]] / m00 in (x, y) order; eigenvalues l1 >= l2
        a, b, c = m["mu20"] / m["m00"], -m["mu11"] / m["m00"], m["mu02"] / m["m00"]
        half_tr, dev = 0.5 * (a + c), np.hypot(0.5 * (a - c), b)
        l1, l2 = half_tr + dev, max(half_tr - dev, 0.0)
        assert np.isclose(props["axis_major_length"][i], 4 * np.sqrt(l1), rtol=1e-9)
        assert np.isclose(props["axis_minor_length"][i], 4 * np.sqrt(l2), rtol=1e-7, atol=1e-9)
        if l1 > 0:
            assert np.isclose(props["eccentricity"][i], np.sqrt(1 - l2 / l1), rtol=1e-6, atol=1e-7)
        if dev > 1e-6 * half_tr:  # the axis direction is defined: compare as a line direction (mod pi)
            # skimage orientation: angle between the row axis and the major axis, in (-pi/2, pi/2]
            ang_cv = 0.5 * np.arctan2(2 * m["mu11"], m["mu20"] - m["mu02"])  # from the x (column) axis
            d = (props["orientation"][i] - (np.pi / 2 - ang_cv)) % np.pi
            assert min(d, np.pi - d) < 1e-6, (label, props["orientation"][i], ang_cv)


@pytest.mark.parametrize("sigma", [0.6, 1.0, 2.5, 16.0])
def test_gaussian_definition_matches_opencv(sigma):
    """The Gaussian the DoG is built from (radius int(4*sigma + 0.5), weights exp(-x^2 / 2 sigma^2) normalised
    to 1, edge-clamped borders) is also OpenCV's GaussianBlur with the same kernel size and BORDER_REPLICATE;
    only the summation order differs, so the two agree to rounding error."""
    from oracle import filters

    rng = np.random.default_rng(int(sigma * 10))
    img = rng.random((96, 120))
    r = int(4.0 * sigma + 0.5)
    got = filters.gaussian(img, sigma)
    want = cv2.GaussianBlur(img, (2 * r + 1, 2 * r + 1), sigmaX=sigma, sigmaY=sigma, borderType=cv2.BORDER_REPLICATE)
    assert np.allclose(got, want, rtol=0, atol=1e-13)
    dog = filters.difference_of_gaussians(img, 0.6, sigma) if sigma > 0.6 else None
    if dog is not None:
        lo = cv2.GaussianBlur(img, (5, 5), sigmaX=0.6, sigmaY=0.6, borderType=cv2.BORDER_REPLICATE)
        assert np.allclose(dog, lo - want, rtol=0, atol=1e-13)


def test_local_threshold_building_blocks_against_opencv():
    """The local-window methods rest on two filters; both restatements are checked against
    OpenCV's independent implementations (summation order differs: tolerance 1e-12 relative, written here):
    * threshold_local: scipy's Gaussian with mode='reflect' == cv2.GaussianBlur(ksize = 2*int(4s+0.5)+1, BORDER_REFLECT)
    * niblack / sauvola: the integral-image window mean / std with np.pad 'reflect' == cv2.boxFilter / sqrBoxFilter
      with BORDER_REFLECT_101."""
    from scipy import ndimage as ndi

    from oracle import threshold

    rng = np.random.default_rng(31)
    image = rng.random((90, 123)) * 1000.0
    for block_size in (5, 35):
        sigma = (block_size - 1) / 6.0
        radius = int(4.0 * sigma + 0.5)
        want = cv2.GaussianBlur(image, (2 * radius + 1, 2 * radius + 1), sigmaX=sigma, sigmaY=sigma, borderType=cv2.BORDER_REFLECT)
        got = threshold.threshold_local(image, block_size)  # offset 0: the smoothed image itself
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
        assert np.array_equal(got, ndi.gaussian_filter(image, sigma, mode="reflect"))
    u16 = rng.integers(0, 60000, (101, 77)).astype(np.uint16)
    for window in (3, 15, (5, 31)):
        wy, wx = (window, window) if np.isscalar(window) else window
        mean, std = threshold._mean_std(u16, window)
        f = u16.astype(np.float64)
        mean_cv = cv2.boxFilter(f, -1, (wx, wy), normalize=True, borderType=cv2.BORDER_REFLECT_101)
        sq_cv = cv2.sqrBoxFilter(f, -1, (wx, wy), normalize=True, borderType=cv2.BORDER_REFLECT_101)
        std_cv = np.sqrt(np.clip(sq_cv - mean_cv * mean_cv, 0, None))
        assert np.abs(mean - mean_cv).max() <= 1e-12 * mean.max() and np.abs(std - std_cv).max() <= 1e-9 * std.max()


def test_triangle_threshold_is_within_one_level_of_opencv():
    """cv2.threshold(THRESH_TRIANGLE) implements the same construction (peak, longer tail, farthest bin from the
    line) with its own end conventions: it lands exactly one grey level beside scikit-image's answer.  A loose,
    independent check that the restated scan finds the same corner of the histogram."""
    from oracle import threshold

    rng = np.random.default_rng(0)
    for _ in range(12):
        shape = (200, 180)
        fg = rng.random(shape) < rng.uniform(0.05, 0.4)
        img = np.where(fg, rng.normal(rng.uniform(120, 220), rng.uniform(5, 25), shape),
                       rng.gamma(2.0, rng.uniform(5, 20), shape)).clip(0, 255).astype(np.uint8)
        t_cv, _ = cv2.threshold(img, 0, 255, cv2.THRESH_BINARY | cv2.THRESH_TRIANGLE)
        assert abs(int(t_cv) - int(threshold.threshold_triangle(img))) <= 1


@pytest.mark.parametrize("seed", range(4))
def test_clear_border_and_relabel_sequential_against_independent_constructions(seed):
    """clear_border (bool and integer masks) and relabel_sequential are definitions rather than algorithms; the
    oracle's restatements are checked against constructions that share no code with them: OpenCV components that
    own a border pixel (per label value for integer masks: skimage clears the border-touching FRAGMENT of a
    label), and np.unique's inverse for the renumbering."""
    from conftest import random_blobs

    mask = random_blobs(100 + seed, (90, 120), 40)
    n, comp = cv2.connectedComponents(mask.astype(np.uint8), connectivity=8)
    touching = np.unique(np.concatenate([comp[0], comp[-1], comp[:, 0], comp[:, -1]]))
    want = mask & ~np.isin(comp, touching[touching > 0])
    got = labeling.clear_border(mask)
    assert got.dtype == np.bool_ and np.array_equal(got, want)
    # integer mask: label values in stripes so that labels touch each other and split into fragments
    values = np.where(mask, 7 + 3 * ((np.arange(120)[None, :] // 9) % 4) + 20 * ((np.arange(90)[:, None] // 30)), 0).astype(np.int64)
    want_int = values.copy()
    for v in np.unique(values)[1:]:
        _, comp_v = cv2.connectedComponents((values == v).astype(np.uint8), connectivity=8)
        touch_v = np.unique(np.concatenate([comp_v[0], comp_v[-1], comp_v[:, 0], comp_v[:, -1]]))
        want_int[np.isin(comp_v, touch_v[touch_v > 0])] = 0
    got_int = labeling.clear_border(values)
    assert got_int.dtype == values.dtype and np.array_equal(got_int, want_int)
    uniq, inverse = np.unique(want_int, return_inverse=True)
    assert uniq[0] == 0
    assert np.array_equal(labeling.relabel_sequential(want_int), inverse.reshape(want_int.shape))
