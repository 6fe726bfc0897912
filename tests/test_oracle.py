"""Pin the oracle against the real third-party code the reference runs (scipy / numpy) and
against the reference's own coarse test assertions (ref: tests/test_masks.py)."""

from __future__ import annotations

import numpy as np
import pytest
from scipy import ndimage as ndi
from scipy.spatial import ConvexHull

import oracle
from oracle import exposure, filters, labeling, percentile, regionprops, threshold

from conftest import make_label_image, random_blobs


@pytest.mark.parametrize("sigma", [0.6, 1.0, 3.3, 10.0, 16.0])
@pytest.mark.parametrize("dtype", [np.uint16, np.float64])
def test_gaussian_restatement_is_bit_identical_to_scipy(sigma, dtype):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 65535, size=(97, 131)).astype(dtype)
    f = filters.img_as_float(img)
    want = ndi.gaussian_filter(f, sigma, mode="nearest", truncate=4.0)
    got = filters.gaussian_restated(f, sigma)
    assert np.array_equal(want, got)


def test_gaussian_restatement_3d():
    rng = np.random.default_rng(4)
    vol = rng.random((9, 40, 33))
    assert np.array_equal(ndi.gaussian_filter(vol, 2.0, mode="nearest", truncate=4.0), filters.gaussian_restated(vol, 2.0))


def test_kernel_weights_match_scipy():
    from scipy.ndimage._filters import _gaussian_kernel1d

    for sigma in (0.6, 1.0, 16.0, 2.5):
        r = int(4.0 * sigma + 0.5)
        assert np.array_equal(_gaussian_kernel1d(sigma, 0, r), filters.gaussian_kernel1d(sigma))


@pytest.mark.parametrize("dtype", [np.uint16, np.float64])
def test_percentile_restatement(dtype):
    rng = np.random.default_rng(5)
    a = (rng.random((64, 77)) * 60000).astype(dtype)
    for q in [(0, 100), (1, 99), (2, 98), (33.3, 66.6), (50, 50.0001)]:
        assert np.array_equal(np.percentile(a, q), percentile.percentile_restated(a, q))
    assert np.percentile(a, 0) == percentile.percentile_restated(a, 0)


def test_rescale_intensity_uint16_computes_in_float64():
    a = np.array([[0, 100, 65535]], dtype=np.uint16)
    out = exposure.rescale_intensity(a, (50.5, 60000.25), (0, 1))
    assert out.dtype == np.float64
    want = (np.clip(a.astype(np.float64), 50.5, 60000.25) - 50.5) / (60000.25 - 50.5) * 1.0 + 0.0
    assert np.array_equal(out, want)


def test_float_histogram_restatement_matches_numpy():
    rng = np.random.default_rng(6)
    for _ in range(5):
        lo, hi = sorted(rng.normal(size=2) * 10)
        x = np.clip(rng.normal((lo + hi) / 2, (hi - lo) / 4, size=20000), lo, hi)
        x[:3] = [lo, hi, (lo + hi) / 2]
        h1, c1 = threshold.histogram_float_256(x)
        h2, c2 = threshold.histogram_float_restated(x)
        assert np.array_equal(h1, h2) and np.array_equal(c1, c2)
    # values sitting exactly on linspace edges
    e = np.linspace(0.0, 1.0, 257)
    x = np.concatenate([e, e[:-1] + 1e-17, np.nextafter(e[1:], 0)])
    assert np.array_equal(threshold.histogram_float_256(x)[0], threshold.histogram_float_restated(x)[0])


def test_otsu_separates_two_populations():
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.normal(100, 5, 5000), rng.normal(200, 5, 5000)])
    t = threshold.threshold_otsu(x)
    assert 120 < t < 180
    xi = np.clip(x, 0, 65535).astype(np.uint16)
    ti = threshold.threshold_otsu(xi)
    assert 110 < ti < 180  # exact-bin histogram: first maximum in the empty gap
    assert threshold.threshold_otsu(np.full((4, 4), 7, np.uint16)) == 7


def test_label_bool_is_scipy_full_connectivity():
    m = random_blobs(1, (80, 90), 40)
    lab = labeling.label(m)
    want, n = ndi.label(m, structure=np.ones((3, 3)))
    assert np.array_equal(lab, want) and lab.max() == n
    firsts = [np.flatnonzero(lab.ravel() == k)[0] for k in range(1, n + 1)]
    assert firsts == sorted(firsts)  # raster order of first pixel


def test_label_multivalue_against_per_value_scipy():
    rng = np.random.default_rng(8)
    img = rng.integers(0, 4, size=(40, 50)) * random_blobs(2, (40, 50), 30)
    lab = labeling.label(img)
    # same partition as labelling every value separately
    seen = {}
    for v in np.unique(img[img > 0]):
        lv, n = ndi.label(img == v, structure=np.ones((3, 3)))
        for k in range(1, n + 1):
            ids = np.unique(lab[lv == k])
            assert len(ids) == 1 and ids[0] not in seen
            seen[ids[0]] = True
    assert len(seen) == lab.max()
    firsts = [np.flatnonzero(lab.ravel() == k)[0] for k in range(1, lab.max() + 1)]
    assert firsts == sorted(firsts)


def test_clear_border_removes_only_touching_fragments():
    img = np.zeros((10, 10), dtype=np.int64)
    img[0:3, 0:3] = 5  # touches the border
    img[6:8, 6:8] = 5  # same value, separate fragment, interior
    img[4, 4] = 2
    out = labeling.clear_border(img)
    assert out[1, 1] == 0 and out[6, 6] == 5 and out[4, 4] == 2
    b = labeling.clear_border(img > 0)
    assert b.dtype == bool and not b[1, 1] and b[6, 6]
    assert np.array_equal(labeling.relabel_sequential(out), np.where(out == 5, 2, np.where(out == 2, 1, 0)))


def test_regionprops_reference_disc_expectations():
    """The reference's own coarse checks (test_masks.py:179-197, :263-295) + SURVEY 8a-10."""
    img = make_label_image((60, 60), [(15, 15, 6), (45, 45, 6)])
    props = oracle.cell_properties(img)
    assert np.allclose(np.c_[props["centroid_y"], props["centroid_x"]], [[15, 15], [45, 45]], atol=2)
    assert np.all(props["circularity"] > 0.85) and np.all(props["circularity"] <= 1.1)
    assert props["area"].tolist() == [109.0, 109.0]
    assert abs(props["perimeter"][0] - 35.31370849898476) < 1e-12
    three = make_label_image((80, 80), [(20, 20, 5), (20, 60, 8), (60, 40, 11)])
    p3 = oracle.cell_properties(three)
    assert p3["area"].tolist() == [69.0, 193.0, 373.0]
    assert p3["area_convex"].tolist() == [69.0, 201.0, 381.0]
    assert np.allclose(p3["perimeter"], [27.313708498984763, 48.97056274847714, 68.28427124746190])
    assert np.allclose(p3["axis_major_length"], [9.3375, 15.6755, 21.7911], atol=1e-3)
    assert np.allclose(p3["orientation"], -np.pi / 4)  # exactly symmetric discs


def _area_convex_qhull(img):
    """skimage's route: Qhull on the diamond offsets, then point-in-polygon on the grid."""
    rr, cc = np.nonzero(img)
    pts = np.concatenate([np.c_[rr - 0.5, cc], np.c_[rr + 0.5, cc], np.c_[rr, cc - 0.5], np.c_[rr, cc + 0.5]])
    pts = np.unique(pts, axis=0)
    hull = ConvexHull(pts)
    v = hull.points[hull.vertices]
    n = len(v)
    count = 0
    for r in range(img.shape[0]):
        for c in range(img.shape[1]):
            inside = True
            for i in range(n):  # hull.vertices are counter-clockwise in 2-D
                a, b = v[i], v[(i + 1) % n]
                if (b[0] - a[0]) * (c - a[1]) - (b[1] - a[1]) * (r - a[0]) < -1e-9:
                    inside = False
                    break
            count += inside
    return float(count)


def test_area_convex_matches_qhull_route():
    rng = np.random.default_rng(9)
    for seed in range(6):
        m = random_blobs(100 + seed, (24, 28), 3, rmax=7)
        lab, n = ndi.label(m, structure=np.ones((3, 3)))
        for k in range(1, n + 1):
            crop = lab[ndi.find_objects(lab)[k - 1]] == k
            assert regionprops.area_convex(crop) == _area_convex_qhull(crop)
    assert regionprops.area_convex(np.ones((1, 1), bool)) == 1.0
    del rng


def test_intensity_props_and_sum():
    rng = np.random.default_rng(42)
    img = make_label_image((60, 60), [(15, 15, 6), (45, 45, 6)])
    inten = rng.integers(100, 1000, size=img.shape).astype(np.uint16)
    t = regionprops.regionprops_table(img, intensity_image=inten, properties=["intensity_mean", "intensity_max", "intensity_min", "intensity_std", "intensity_sum"])
    v = inten[img == 1]
    assert t["intensity_sum"][0] == v.sum() and t["intensity_sum"].dtype == np.uint64
    assert t["intensity_mean"][0] == v.mean() and t["intensity_std"][0] == v.std()
    assert t["intensity_max"].dtype == np.float64 and t["intensity_max"][0] == v.max()


def test_ops_error_messages_follow_reference():
    x = np.arange(12, dtype=np.uint16).reshape(3, 4)
    with pytest.raises(ValueError, match="Invalid percentile range"):
        oracle.rescale_by_percentile(x, (5, 5))
    with pytest.raises(ValueError, match="must be smaller than high_sigma"):
        oracle.subtract_background_dog(x, 2, 1)
    with pytest.raises(ValueError, match="Percentile must be between 0 and 100"):
        oracle.subtract_background_dog(x, percentile=101)
    assert not oracle.apply_threshold(np.full((3, 3), 4, np.uint16)).any()
    with pytest.raises(ValueError, match="No cells remain"):
        oracle.process_mask(np.ones((5, 5), bool), True)
