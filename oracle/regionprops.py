"""Oracle: ``skimage.measure.regionprops_table`` for the hot-path properties.
TEST INFRASTRUCTURE ONLY.

Restates the per-region arithmetic of scikit-image 0.25.2 ``_regionprops.py`` /
``_moments.py`` / ``_regionprops_utils.py`` / ``morphology/convex_hull.py`` as used by the
reference's ``SegmentationMask.cell_properties`` (``masks.py:286-289`` morphology,
``masks.py:317-326`` intensity).  SURVEY.md 8a item 10.  One region per present label, in
ascending label order, found through ``scipy.ndimage.find_objects`` exactly as skimage does.
"""

from __future__ import annotations

import math

import numpy as np
from scipy import ndimage as ndi

SQRT2 = math.sqrt(2)
_PERIM_WEIGHTS = np.zeros(50, dtype=np.float64)
_PERIM_WEIGHTS[[5, 7, 15, 17, 25, 27]] = 1
_PERIM_WEIGHTS[[21, 33]] = SQRT2
_PERIM_WEIGHTS[[13, 23]] = (1 + SQRT2) / 2
_STREL_4 = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=np.uint8)
_PERIM_KERNEL = np.array([[10, 2, 10], [2, 1, 2], [10, 2, 10]])

# properties whose table columns are integer typed (skimage COL_DTYPES); all others float64
_INT_PROPS = {"label", "bbox"}

SCALAR_PROPS = (
    "label",
    "area",
    "area_convex",
    "perimeter",
    "eccentricity",
    "solidity",
    "axis_major_length",
    "axis_minor_length",
    "orientation",
    "intensity_mean",
    "intensity_max",
    "intensity_min",
    "intensity_std",
    "intensity_sum",  # extension (north star): exact integer sum, not a skimage property
)


def moments_central(image: np.ndarray, center, order: int = 3) -> np.ndarray:
    """skimage ``moments_central``: per-axis ``(arange(n) - center)**p`` tables contracted
    with the 0/1 image by ``np.dot``."""
    calc = image.astype(np.float64, copy=False)
    for dim, dim_length in enumerate(image.shape):
        delta = np.arange(dim_length, dtype=np.float64) - center[dim]
        powers_of_delta = delta[:, np.newaxis] ** np.arange(order + 1, dtype=np.float64)
        calc = np.rollaxis(calc, dim, image.ndim)
        calc = np.dot(calc, powers_of_delta)
        calc = np.rollaxis(calc, -1, dim)
    return calc


def inertia_tensor(mu: np.ndarray, ndim: int) -> np.ndarray:
    """skimage ``inertia_tensor``: diagonal ``(sum(mu[corners2]) - mu[corners2]) / mu0``,
    off-diagonal ``-mu11 / mu0``."""
    mu0 = mu[(0,) * ndim]
    result = np.zeros((ndim, ndim), dtype=mu.dtype)
    corners2 = tuple(2 * np.eye(ndim, dtype=int))
    d = np.diag(result)
    d.flags.writeable = True
    d[:] = (np.sum(mu[corners2]) - mu[corners2]) / mu0
    import itertools

    for dims in itertools.combinations(range(ndim), 2):
        mu_index = np.zeros(ndim, dtype=int)
        mu_index[list(dims)] = 1
        result[dims] = -mu[tuple(mu_index)] / mu0
        result.T[dims] = -mu[tuple(mu_index)] / mu0
    return result


def inertia_tensor_eigvals(T: np.ndarray):
    eigvals = np.linalg.eigvalsh(T)
    eigvals = np.clip(eigvals, 0, None, out=eigvals)
    return sorted(eigvals, reverse=True)


def perimeter(image: np.ndarray) -> float:
    """skimage ``perimeter(image, neighborhood=4)`` on a 2-D binary crop."""
    img = image.astype(np.uint8)
    eroded = ndi.binary_erosion(img, _STREL_4, border_value=0)
    border = img - eroded
    code = ndi.convolve(border, _PERIM_KERNEL, mode="constant", cval=0)
    hist = np.bincount(code.ravel(), minlength=50)
    return float(hist @ _PERIM_WEIGHTS)


def _hull_monotone_chain(points: np.ndarray) -> np.ndarray:
    """Exact integer convex hull (Andrew's monotone chain), counter-clockwise, no collinear
    points.  ``points`` are int64 (already scaled so that every coordinate is an integer)."""
    pts = np.unique(points, axis=0)
    pts = [tuple(map(int, p)) for p in pts]
    if len(pts) <= 2:
        return np.array(pts, dtype=np.int64)

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower = []
    for p in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 0:
            lower.pop()
        lower.append(p)
    upper = []
    for p in reversed(pts):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 0:
            upper.pop()
        upper.append(p)
    return np.array(lower[:-1] + upper[:-1], dtype=np.int64)


def area_convex(image: np.ndarray) -> float:
    """skimage ``convex_hull_image(image).sum()`` for a 2-D crop: the hull of the four
    edge-midpoints ``(r+-0.5, c)``, ``(r, c+-0.5)`` of every pixel, then every integer grid
    point inside, on an edge of, or at a vertex of that polygon (include_borders=True).

    All coordinates are half-integers, so the test is done exactly in integers scaled by 2.
    The closed hull intersected with an integer row is an interval; its lattice points are
    counted per row.
    """
    rr, cc = np.nonzero(image)
    if rr.size == 0:
        return 0.0
    r2 = 2 * rr.astype(np.int64)
    c2 = 2 * cc.astype(np.int64)
    pts = np.concatenate(
        [
            np.stack([r2 - 1, c2], axis=1),
            np.stack([r2 + 1, c2], axis=1),
            np.stack([r2, c2 - 1], axis=1),
            np.stack([r2, c2 + 1], axis=1),
        ]
    )
    hull = _hull_monotone_chain(pts)
    nv = len(hull)
    total = 0
    for r in range(int(rr.min()), int(rr.max()) + 1):
        y = 2 * r
        lo = None  # exact rational bounds lo_n/lo_d <= c2 <= hi_n/hi_d
        hi = None
        for i in range(nv):
            (y0, x0), (y1, x1) = hull[i], hull[(i + 1) % nv]
            y0, x0, y1, x1 = int(y0), int(x0), int(y1), int(x1)
            if (y0 - y) * (y1 - y) > 0:
                continue
            if y0 == y1:
                cands = [(x0, 1), (x1, 1)]
            else:
                num = x0 * (y1 - y0) + (x1 - x0) * (y - y0)
                den = y1 - y0
                if den < 0:
                    num, den = -num, -den
                cands = [(num, den)]
            for n_, d_ in cands:
                if lo is None or n_ * lo[1] < lo[0] * d_:
                    lo = (n_, d_)
                if hi is None or n_ * hi[1] > hi[0] * d_:
                    hi = (n_, d_)
        if lo is None:
            continue
        # integer columns c with lo <= 2c <= hi
        c_min = -((-lo[0]) // (2 * lo[1]))  # ceil(lo / 2)
        c_max = hi[0] // (2 * hi[1])  # floor(hi / 2)
        c_min = max(c_min, 0)
        c_max = min(c_max, image.shape[1] - 1)
        if c_max >= c_min:
            total += c_max - c_min + 1
    return float(total)


def _region_scalar(prop: str, ctx: dict):
    """One scalar property of one region.  ``ctx`` caches the skimage intermediates."""
    img = ctx["image"]
    ndim = img.ndim
    if prop == "label":
        return ctx["label"]
    if prop == "area":
        return float(np.sum(img) * 1.0)
    if prop == "area_convex":
        if "area_convex" not in ctx:
            ctx["area_convex"] = area_convex(img)
        return ctx["area_convex"]
    if prop == "solidity":
        return _region_scalar("area", ctx) / _region_scalar("area_convex", ctx)
    if prop == "perimeter":
        return perimeter(img)
    if prop in ("eccentricity", "axis_major_length", "axis_minor_length", "orientation"):
        if "T" not in ctx:
            coords = np.argwhere(img)
            centroid_local = tuple(coords.mean(axis=0))
            mu = moments_central(img.astype(np.uint8), centroid_local, order=3)
            ctx["T"] = inertia_tensor(mu, ndim)
            ctx["eig"] = inertia_tensor_eigvals(ctx["T"])
        T, ev = ctx["T"], ctx["eig"]
        if prop == "axis_major_length":
            if ndim == 2:
                return 4 * math.sqrt(ev[0])
            return math.sqrt(10 * (ev[0] + ev[1] - ev[2]))
        if prop == "axis_minor_length":
            if ndim == 2:
                return 4 * math.sqrt(ev[-1])
            return math.sqrt(10 * max(-ev[0] + ev[1] + ev[2], 0.0))
        if prop == "eccentricity":
            l1, l2 = ev
            if l1 == 0:
                return 0.0
            return math.sqrt(1 - l2 / l1)
        a, b, b, c = T.flat
        if a - c == 0:
            return math.pi / 4.0 if b < 0 else -math.pi / 4.0
        return 0.5 * math.atan2(-2 * b, c - a)
    vals = ctx["intensity"][img]
    if prop == "intensity_mean":
        return float(np.mean(vals, axis=0))
    if prop == "intensity_max":
        return float(np.max(vals, axis=0))
    if prop == "intensity_min":
        return float(np.min(vals, axis=0))
    if prop == "intensity_std":
        return float(np.std(vals, axis=0))
    if prop == "intensity_sum":
        return int(vals.sum(dtype=np.uint64))
    raise ValueError(f"oracle.regionprops: unsupported property {prop!r}")


def regionprops_table(label_image: np.ndarray, intensity_image=None, properties=("label", "bbox")):
    """Dict of 1-D columns, skimage naming (``centroid-0``, ``bbox-2`` ...), skimage column
    dtypes (label/bbox int64, everything else float64; ``intensity_sum`` uint64)."""
    label_image = np.asarray(label_image)
    if not np.issubdtype(label_image.dtype, np.integer):
        raise TypeError("Non-integer label_image types are ambiguous")
    ndim = label_image.ndim
    objects = ndi.find_objects(label_image)
    regions = [(i + 1, sl) for i, sl in enumerate(objects) if sl is not None]
    n = len(regions)
    out: dict[str, np.ndarray] = {}
    columns: list[tuple[str, str, int | None]] = []
    for prop in properties:
        if prop == "bbox":
            for k in range(2 * ndim):
                out[f"bbox-{k}"] = np.empty(n, dtype=np.int64)
        elif prop == "centroid":
            for k in range(ndim):
                out[f"centroid-{k}"] = np.empty(n, dtype=np.float64)
        elif prop == "label":
            out[prop] = np.empty(n, dtype=np.int64)
        elif prop == "intensity_sum":
            out[prop] = np.empty(n, dtype=np.uint64)
        elif prop in SCALAR_PROPS:
            out[prop] = np.empty(n, dtype=np.float64)
        else:
            raise ValueError(f"oracle.regionprops: unsupported property {prop!r}")
    del columns
    for row, (lab, sl) in enumerate(regions):
        img = label_image[sl] == lab
        ctx = {"image": img, "label": lab}
        if intensity_image is not None:
            ctx["intensity"] = np.asarray(intensity_image)[sl]
        for prop in properties:
            if prop == "bbox":
                for k in range(ndim):
                    out[f"bbox-{k}"][row] = sl[k].start
                    out[f"bbox-{k + ndim}"][row] = sl[k].stop
            elif prop == "centroid":
                coords = np.argwhere(img) + np.array([s.start for s in sl])
                cen = coords.astype(np.float64).mean(axis=0)
                for k in range(ndim):
                    out[f"centroid-{k}"][row] = cen[k]
            else:
                out[prop][row] = _region_scalar(prop, ctx)
    return out


def regionprops_table_3d(label_volume: np.ndarray, intensity_volumes: dict | None = None) -> dict:
    """3-D restatement (BASELINE config 4; no reference entry point, SURVEY.md note N3) of
    ``regionprops_table(label_volume, intensity_image=..., properties=(label, area, bbox, centroid,
    inertia_tensor_eigvals, axis_major_length, axis_minor_length, intensity_*))``: same per-region
    code path as the 2-D properties above (find_objects crop, local-coordinate central moments,
    inertia tensor, ``eigvalsh``) with skimage's 3-D axis-length formulas."""
    intensity_volumes = intensity_volumes or {}
    objects = ndi.find_objects(label_volume)
    rows = []
    for i, sl in enumerate(objects):
        if sl is None:
            continue
        label = i + 1
        crop = label_volume[sl] == label
        area = float(np.sum(crop))
        coords = np.argwhere(crop)
        local_centroid = coords.mean(axis=0)
        mu = moments_central(crop.astype(np.uint8), local_centroid, order=2)
        ev = inertia_tensor_eigvals(inertia_tensor(mu, 3))
        row = {"label": label, "area": area}
        for d in range(3):
            row[f"bbox-{d}"] = sl[d].start
            row[f"bbox-{d + 3}"] = sl[d].stop
            row[f"centroid-{d}"] = sl[d].start + local_centroid[d]
            row[f"inertia_tensor_eigvals-{d}"] = ev[d]
        row["axis_major_length"] = math.sqrt(10 * (ev[0] + ev[1] - ev[2]))
        row["axis_minor_length"] = math.sqrt(10 * max(-ev[0] + ev[1] + ev[2], 0))
        for name, vol in intensity_volumes.items():
            vals = vol[sl][crop]
            row[f"intensity_sum_{name}"] = int(vals.sum(dtype=np.uint64))
            row[f"intensity_mean_{name}"] = float(np.mean(vals))
            row[f"intensity_max_{name}"] = float(np.max(vals))
            row[f"intensity_min_{name}"] = float(np.min(vals))
            row[f"intensity_std_{name}"] = float(np.std(vals))
        rows.append(row)
    if not rows:
        return {}
    return {k: np.array([r[k] for r in rows]) for k in rows[0]}
