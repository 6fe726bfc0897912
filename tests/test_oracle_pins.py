"""More pins for the oracle's restatements of scikit-image's own Python (last table row), each against a
construction that shares no code with the oracle: scipy.ndimage's measurement routines and NumPy's covariance /
eigen-decomposition for the 3-D per-object table, solid ellipsoids with known axes, hand-derived perimeters from the
published weight table, and the fixed-point / cross-entropy definition of Li's threshold."""

from __future__ import annotations

import math

import numpy as np
import pytest
import scipy.ndimage as ndi

import oracle
from oracle import regionprops as orp
from oracle import threshold as oth


def _blobs3d(seed: int, shape=(24, 40, 36), n=7):
    rng = np.random.default_rng(seed)
    zz, yy, xx = np.indices(shape)
    lab = np.zeros(shape, np.int32)
    for k in range(1, n + 1):
        c = [rng.uniform(4, s - 4) for s in shape]
        r = rng.uniform(2.5, 5.5, 3)
        inside = ((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2 <= 1.0
        lab[inside & (lab == 0)] = k
    vol = rng.integers(0, 65536, shape).astype(np.uint16)
    return lab, vol


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_3d_table_against_scipy_measurements_and_numpy_covariance(seed):
    lab, vol = _blobs3d(seed)
    got = orp.regionprops_table_3d(lab, {"A": vol})
    ids = np.unique(lab[lab > 0])
    assert np.array_equal(got["label"], ids)
    assert np.array_equal(got["area"], ndi.sum_labels(np.ones_like(lab), lab, ids))
    com = np.array(ndi.center_of_mass(np.ones_like(lab), lab, ids))
    for d in range(3):
        np.testing.assert_allclose(got[f"centroid-{d}"], com[:, d], rtol=1e-12)
    v64 = vol.astype(np.float64)
    assert np.array_equal(got["intensity_sum_A"], ndi.sum_labels(vol.astype(np.uint64), lab, ids).astype(np.uint64))
    np.testing.assert_allclose(got["intensity_mean_A"], ndi.mean(v64, lab, ids), rtol=1e-12)
    np.testing.assert_allclose(got["intensity_std_A"], ndi.standard_deviation(v64, lab, ids), rtol=1e-10)
    assert np.array_equal(got["intensity_max_A"], ndi.maximum(vol, lab, ids))
    assert np.array_equal(got["intensity_min_A"], ndi.minimum(vol, lab, ids))
    for row, k in enumerate(ids):
        pts = np.argwhere(lab == k)
        for d in range(3):
            assert got[f"bbox-{d}"][row] == pts[:, d].min() and got[f"bbox-{d + 3}"][row] == pts[:, d].max() + 1
        # inertia tensor = tr(S) I - S for the coordinates' population covariance S: eigenvalues tr(S) - lambda_i, descending
        lam = np.linalg.eigvalsh(np.cov(pts.T.astype(np.float64), bias=True))
        want = np.sort(lam.sum() - lam)[::-1]
        have = [got[f"inertia_tensor_eigvals-{d}"][row] for d in range(3)]
        np.testing.assert_allclose(have, want, rtol=1e-9, atol=1e-12)
        # full axis lengths of the ellipsoid with the same second moments: sqrt(20 lambda)
        assert got["axis_major_length"][row] == pytest.approx(math.sqrt(20 * lam.max()), rel=1e-9)


def test_3d_axis_lengths_of_a_solid_ellipsoid():
    """A digitised solid ellipsoid with semi-axes (6, 9, 14): the table's major axis is 2 x 14 and the minor axis 2 x 6
    up to the discretisation (a known answer that needs no code at all)."""
    zz, yy, xx = np.indices((20, 28, 40))
    lab = ((((zz - 9.5) / 6) ** 2 + ((yy - 13.5) / 9) ** 2 + ((xx - 19.5) / 14) ** 2) <= 1.0).astype(np.int32)
    got = orp.regionprops_table_3d(lab)
    assert got["axis_major_length"][0] == pytest.approx(28.0, rel=0.02)
    assert got["axis_minor_length"][0] == pytest.approx(12.0, rel=0.03)
    assert got["area"][0] == pytest.approx(4 / 3 * math.pi * 6 * 9 * 14, rel=0.02)


def test_perimeter_weights_on_hand_derived_shapes():
    """skimage's 4-neighbourhood perimeter weighs a border pixel by its border neighbours (Benkrid & Crookes): two edge
    neighbours -> 1, two diagonal neighbours -> sqrt 2, one of each -> (1 + sqrt 2) / 2.  Worked out by hand:
    * a filled a x b rectangle (a, b >= 3): its border is a one-pixel frame of 2a + 2b - 4 pixels, each with exactly two
      edge neighbours in the frame -> 2a + 2b - 4;
    * a one-pixel diagonal line of n pixels: the n - 2 inner pixels have two diagonal neighbours -> (n - 2) sqrt 2 (the
      end pixels, with one neighbour, carry no weight);
    * a one-pixel "knight" polyline (two right, one down-right, repeated): the joints have one edge and one diagonal
      neighbour."""
    for a, b in ((3, 3), (5, 9), (12, 4)):
        img = np.zeros((a + 4, b + 4), bool)
        img[2 : 2 + a, 2 : 2 + b] = True
        assert orp.perimeter(img) == pytest.approx(2 * a + 2 * b - 4, abs=1e-12)
    n = 9
    diag = np.zeros((n + 2, n + 2), bool)
    diag[np.arange(1, n + 1), np.arange(1, n + 1)] = True
    assert orp.perimeter(diag) == pytest.approx((n - 2) * math.sqrt(2), abs=1e-12)
    # x x .        pixels (0,0) (0,1) (1,2) (1,3) (2,4) (2,5): inner pixels all have one edge + one diagonal neighbour
    # . . x x
    pts = [(1, 1), (1, 2), (2, 3), (2, 4), (3, 5), (3, 6)]
    line = np.zeros((5, 8), bool)
    for p in pts:
        line[p] = True
    assert orp.perimeter(line) == pytest.approx(4 * (1 + math.sqrt(2)) / 2, abs=1e-12)


def _cross_entropy(values: np.ndarray, t: float) -> float:
    """Li & Lee's criterion: -sum_{x <= t} x log(mean_back) - sum_{x > t} x log(mean_fore) (constant terms dropped)."""
    back, fore = values[values <= t], values[values > t]
    if back.size == 0 or fore.size == 0 or back.mean() <= 0:
        return math.inf
    return float(-(back.sum() * math.log(back.mean()) + fore.sum() * math.log(fore.mean())))


@pytest.mark.parametrize("seed,dtype", [(0, np.uint16), (1, np.uint16), (2, np.float64), (3, np.float64)])
def test_li_threshold_is_the_fixed_point_and_a_minimum_of_the_cross_entropy(seed, dtype):
    rng = np.random.default_rng(seed)
    a = np.concatenate([rng.normal(900, 120, 6000), rng.normal(3200, 400, 3000)]).clip(1, 65535)
    image = a.astype(dtype).reshape(90, 100) if dtype == np.uint16 else (a / 65535.0).reshape(90, 100)
    t = float(oth.threshold_li(image.copy()))
    vals = image.astype(np.float64).ravel()
    shifted, ts = vals - vals.min(), t - vals.min()
    back, fore = shifted[shifted <= ts].mean(), shifted[shifted > ts].mean()
    t_fix = (back - fore) / (math.log(back) - math.log(fore))
    tol = 0.5 if dtype == np.uint16 else np.min(np.diff(np.unique(vals))) / 2
    assert abs(t_fix - ts) <= 2 * tol + 1e-12  # the iteration stopped at its fixed point
    # and the fixed point is a minimum of the criterion among the thresholds around it
    # (the criterion is flat wherever no sample lies, so compare its VALUE, not the location of the minimum)
    span = np.linspace(0.3 * ts, 2.5 * ts, 221)
    ce = np.array([_cross_entropy(shifted, s) for s in span])
    assert _cross_entropy(shifted, ts) <= ce.min() + 1e-9 * abs(ce.min())
    assert ce.max() > ce.min() + 1e-4 * abs(ce.min())  # ... and the scan does cover thresholds that are worse
    assert oracle is not None


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_float_image_otsu_is_the_maximiser_of_the_between_class_variance(seed):
    """The float-image path of `threshold_otsu` (256 bins of np.histogram over [min, max], threshold = a bin centre;
    what `apply_threshold` sees after `rescale_by_percentile`, ref: operations.py:186, 214): the returned threshold
    maximises Otsu's between-class variance w0 * w1 * (mu0 - mu1)^2 evaluated from the definition -- explicit sums over
    the two classes for every cut, no cumulative sums, no shared code -- and the mask it gives is the one OpenCV's Otsu
    gives on the same data quantised to its 256 bins."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(seed)
    n = 40000
    lo = rng.normal(0.2, 0.05 + 0.01 * seed, n)
    hi = rng.normal(0.7, 0.08, n // (2 + seed))
    x = np.clip(np.concatenate([lo, hi]), 0.0, 1.0)
    x = x[: (x.size // 100) * 100].reshape(-1, 100)
    t = float(oth.threshold_otsu(x))
    hist, edges = np.histogram(x, bins=256)
    centers = (edges[:-1] + edges[1:]) / 2.0
    assert np.min(np.abs(centers - t)) == 0.0  # a bin centre, bit for bit
    best, best_k = -1.0, -1
    for k in range(255):  # classes: bins 0..k and k+1..255
        w0, w1 = hist[: k + 1].sum(), hist[k + 1 :].sum()
        if w0 == 0 or w1 == 0:
            continue
        mu0 = float((hist[: k + 1] * centers[: k + 1]).sum()) / w0
        mu1 = float((hist[k + 1 :] * centers[k + 1 :]).sum()) / w1
        var = float(w0) * float(w1) * (mu0 - mu1) ** 2
        if var > best:
            best, best_k = var, k
    assert abs(int(np.argmin(np.abs(centers - t))) - best_k) <= 0  # the same cut
    # OpenCV on the image quantised to the same 256 bins: same foreground
    q = np.clip(((x - x.min()) / (x.max() - x.min()) * 256.0).astype(np.int64), 0, 255).astype(np.uint8)
    t_cv, _ = cv2.threshold(q, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
    mask_oracle = x > t
    mask_cv = q > t_cv
    assert np.count_nonzero(mask_oracle != mask_cv) <= 0.002 * x.size  # samples on the cut bin's edge may differ
