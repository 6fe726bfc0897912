"""Cell outlines (ref: ``masks.py:68-115``, ``masks.py:229-245``).

CPU part: the oracle against the reference's own expectations (``tests/test_masks.py:86-149``) and
against the real ``cv2.findContours``; the product's host-side contour assembly against the oracle.
GPU part (``-m gpu``): ``SegmentationMask.cell_outlines`` for both extractors against the oracle, point
for point."""

from __future__ import annotations

import numpy as np
import pytest
from scipy import ndimage as ndi

from oracle import labeling
from oracle import outlines as oracle_outlines

from conftest import make_label_image, random_blobs

from arcadia_microscopy_tools_b200 import masks


def _square_keys_numpy(lab: np.ndarray) -> np.ndarray:
    """What ``amt_outline_squares`` + sort returns, by brute force."""
    h, w = lab.shape
    out = []
    for r0 in range(h - 1):
        for c0 in range(w - 1):
            q = [int(lab[r0, c0]), int(lab[r0, c0 + 1]), int(lab[r0 + 1, c0]), int(lab[r0 + 1, c0 + 1])]
            for value in sorted(set(q) - {0}):
                case = sum((q[i] == value) << i for i in range(4))
                if case != 15:
                    out.append((value << 34) | (r0 << 19) | (c0 << 4) | case)
    return np.sort(np.array(out, dtype=np.uint64))


def _touching_labels(seed: int, shape) -> np.ndarray:
    """Blobs cut into vertical stripes: labels that touch each other and fragments of one label."""
    blobs = random_blobs(seed, shape, 25)
    lab, _ = ndi.label(blobs, structure=np.ones((3, 3)))
    striped = np.where(lab > 0, lab * 3 + (np.arange(shape[1])[None, :] // 6) % 3, 0)
    return labeling.relabel_sequential(striped).astype(np.int64)


# ------------------------------------------------------------------ oracle vs the reference's assertions
def test_oracle_skimage_outlines_meet_the_reference_expectations():
    interior = make_label_image((50, 50), [(25, 25, 8)])
    multi = make_label_image((60, 60), [(15, 15, 6), (45, 45, 6)])
    assert len(oracle_outlines.extract_outlines_skimage(interior)) == 1
    two = oracle_outlines.extract_outlines_skimage(multi)
    assert len(two) == 2
    (outline,) = oracle_outlines.extract_outlines_skimage(interior)
    assert outline.ndim == 2 and outline.shape[1] == 2 and len(outline) > 0
    assert np.issubdtype(outline.dtype, np.floating)
    assert outline.min() >= 0 and outline[:, 0].max() < 50 and outline[:, 1].max() < 50
    assert np.linalg.norm(outline.mean(axis=0) - [25, 25]) < 2.0
    np.testing.assert_array_almost_equal(outline[0], outline[-1])  # closed
    assert np.linalg.norm(two[0].mean(axis=0) - [15, 15]) < 2.0 and np.linalg.norm(two[1].mean(axis=0) - [45, 45]) < 2.0
    (near_border,) = oracle_outlines.extract_outlines_skimage(make_label_image((50, 50), [(4, 25, 4)]))
    assert len(near_border) > 0
    speck = np.zeros((5, 5), dtype=np.int64)
    speck[2, 2] = 1
    (tiny,) = oracle_outlines.extract_outlines_skimage(speck)
    assert tiny.ndim == 2 and tiny.shape[1] == 2
    # a one-pixel cell: the diamond through its four edge midpoints, closed; the start follows from the
    # joining order (segments of squares (1,1), (1,2), (2,1), (2,2): the third one is prepended)
    assert tiny.tolist() == [[2.5, 2.0], [2.0, 1.5], [1.5, 2.0], [2.0, 2.5], [2.5, 2.0]]
    # scikit-image's own docstring example: a[0, 0] = 1 in a 3x3 image -> [[0, 0.5], [0.5, 0]]
    corner = np.zeros((3, 3))
    corner[0, 0] = 1
    assert [c.tolist() for c in oracle_outlines.find_contours(corner, 0.5)] == [[[0.0, 0.5], [0.5, 0.0]]]


def test_oracle_cellpose_outlines_are_cv2_borders():
    """The cellpose leg is the real OpenCV routine; check the wrapper's conventions on a disc."""
    multi = make_label_image((60, 60), [(15, 15, 6), (45, 45, 6), (30, 5, 1)])
    got = oracle_outlines.extract_outlines_cellpose(multi)
    assert len(got) == 3 and got[2].shape == (0, 2)  # the r=1 speck has a 1-point border: dropped
    for outline, centre in zip(got[:2], [(15, 15), (45, 45)]):
        assert outline.dtype.kind == "i" and outline.shape[1] == 2
        assert np.all(multi[outline[:, 0], outline[:, 1]] == multi[centre])  # (y, x) order, on the cell
        assert np.linalg.norm(outline.mean(axis=0) - centre) < 1.0


# ------------------------------------------------------------------ product host assembly vs oracle
@pytest.mark.parametrize("seed", range(6))
def test_contours_from_square_keys_match_the_oracle(seed):
    lab = _touching_labels(seed, (40 + 3 * seed, 52)) if seed % 2 else make_label_image((48, 40), [(12, 12, 7), (30, 25, 9), (2, 30, 3)])
    n = int(lab.max())
    want = oracle_outlines.extract_outlines_skimage(lab)
    got = masks._contours_from_square_keys(_square_keys_numpy(lab), n)
    assert len(got) == len(want) == n
    for a, b in zip(got, want):
        assert a.dtype == np.float64 and a.shape == b.shape and np.array_equal(a, b)


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("extractor", ["cellpose", "skimage"])
def test_cell_outlines_match_the_oracle(extractor):
    from arcadia_microscopy_tools_b200.synthetic import make_fov

    reference = oracle_outlines.extract_outlines_cellpose if extractor == "cellpose" else oracle_outlines.extract_outlines_skimage
    cases = [
        make_label_image((50, 50), [(25, 25, 8)]),
        make_label_image((60, 60), [(15, 15, 6), (45, 45, 6)]),
        make_label_image((50, 50), [(4, 25, 4)]),
        make_label_image((80, 80), [(20, 20, 5), (20, 60, 8), (60, 40, 11), (70, 70, 1)]),
        _touching_labels(1, (97, 131)),
        _touching_labels(2, (64, 200)),
        make_fov(5, 2, 384, 512, n_cells=120)[1].astype(np.int64),
    ]
    speck = np.zeros((5, 5), dtype=np.int64)
    speck[2, 2] = 1
    cases.append(speck)
    for i, lab in enumerate(cases):
        m = masks.SegmentationMask(lab, remove_edge_cells=False, outline_extractor=extractor)
        want = reference(m.label_image)
        got = m.cell_outlines
        assert len(got) == len(want) == m.num_cells, (i, len(got), len(want))
        for k, (a, b) in enumerate(zip(got, want)):
            assert a.shape == b.shape and a.dtype.kind == b.dtype.kind and np.array_equal(a, b), (extractor, i, k)


@pytest.mark.gpu
def test_outlines_of_ragged_threshold_components():
    """A noisy threshold mask: a few large, ragged components with hundreds of local tops, holes and specks (the
    border follower must start each outer border exactly once, at its raster-first pixel)."""
    rng = np.random.default_rng(12)
    field = ndi.gaussian_filter(rng.random((384, 416)), 2.0)
    lab, n = ndi.label(field > np.percentile(field, 55), structure=np.ones((3, 3)))
    assert n > 20 and np.bincount(lab.ravel())[1:].max() > 5000
    lab = lab.astype(np.int64)
    for extractor, reference in (("cellpose", oracle_outlines.extract_outlines_cellpose),
                                 ("skimage", oracle_outlines.extract_outlines_skimage)):
        got = masks.SegmentationMask(lab, remove_edge_cells=False, outline_extractor=extractor).cell_outlines
        want = reference(lab)
        assert len(got) == len(want) == n
        for k, (a, b) in enumerate(zip(got, want)):
            assert a.shape == b.shape and np.array_equal(a, b), (extractor, k)


@pytest.mark.gpu
def test_module_level_outline_helpers_like_the_reference():
    """ref: tests/test_masks.py:86-149 imports ``_extract_outlines_skimage`` from ``masks`` and feeds it label images."""
    multi = make_label_image((60, 60), [(15, 15, 6), (45, 45, 6)])
    for helper, want in ((masks._extract_outlines_skimage, oracle_outlines.extract_outlines_skimage),
                         (masks._extract_outlines_cellpose, oracle_outlines.extract_outlines_cellpose)):
        got = helper(multi)
        assert len(got) == 2 and all(np.array_equal(a, b) for a, b in zip(got, want(multi)))
    speck = np.zeros((5, 5), dtype=np.int64)
    speck[2, 2] = 1
    outlines = masks._extract_outlines_skimage(speck)
    assert len(outlines) == 1 and outlines[0].ndim == 2 and outlines[0].shape[1] == 2
    assert masks._extract_outlines_skimage(np.zeros((8, 8), dtype=np.int64)) == []


@pytest.mark.gpu
def test_cell_outlines_reference_contract():
    """ref: tests/test_masks.py:86-149, through SegmentationMask(outline_extractor='skimage')."""
    m = masks.SegmentationMask(make_label_image((50, 50), [(25, 25, 8)]), remove_edge_cells=False, outline_extractor="skimage")
    (outline,) = m.cell_outlines
    assert outline.ndim == 2 and outline.shape[1] == 2 and len(outline) > 0 and np.issubdtype(outline.dtype, np.floating)
    assert outline.min() >= 0 and outline.max() < 50
    assert np.linalg.norm(outline.mean(axis=0) - [25, 25]) < 2.0
    np.testing.assert_array_almost_equal(outline[0], outline[-1])
    two = masks.SegmentationMask(make_label_image((60, 60), [(15, 15, 6), (45, 45, 6)]), remove_edge_cells=False,
                                 outline_extractor="skimage").cell_outlines
    assert len(two) == 2
    assert np.linalg.norm(two[0].mean(axis=0) - [15, 15]) < 2.0 and np.linalg.norm(two[1].mean(axis=0) - [45, 45]) < 2.0
    edge = masks.SegmentationMask(make_label_image((50, 50), [(4, 25, 4)]), remove_edge_cells=False, outline_extractor="skimage")
    assert len(edge.cell_outlines[0]) > 0
    kept = edge.filter("area", min_value=1)
    assert kept.outline_extractor == "skimage"


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4])
def test_find_contours_visits_every_crossing_edge_exactly_once(seed):
    """An independent pin for the marching-squares restatement (`oracle.outlines.find_contours`, the "skimage"
    extractor of ref masks.py:82-115).  On a 0/1 image at level 0.5 every contour vertex is the midpoint of two
    4-adjacent pixels with different values, and the contours together pass through every such midpoint exactly once;
    a segment joins the midpoints of two edges of one 2x2 square (length sqrt(1/2) or 1); a contour that does not reach
    the image border is closed; around a high island and around a hole the contours wind in opposite directions."""
    rng = np.random.default_rng(seed)
    img = (ndi.gaussian_filter(rng.normal(size=(48, 56)), 2.0) > 0.02).astype(np.uint8)
    img[20:30, 20:30] = 1
    img[23:27, 23:27] = 0  # a hole
    contours = oracle_outlines.find_contours(img, 0.5)
    want = set()
    h, w = img.shape
    for r in range(h):
        for c in range(w):
            if c + 1 < w and img[r, c] != img[r, c + 1]:
                want.add((float(r), c + 0.5))
            if r + 1 < h and img[r, c] != img[r + 1, c]:
                want.add((r + 0.5, float(c)))
    seen = []
    for contour in contours:
        closed = np.array_equal(contour[0], contour[-1])
        pts = contour[:-1] if closed else contour
        seen.extend(map(tuple, pts.tolist()))
        steps = np.linalg.norm(np.diff(contour, axis=0), axis=1)
        assert np.all(np.isclose(steps, np.sqrt(0.5)) | np.isclose(steps, 1.0))
        if not closed:  # an open contour starts and ends on the image border
            for end in (contour[0], contour[-1]):
                assert end[0] in (0.0, h - 1.0) or end[1] in (0.0, w - 1.0)
    assert len(seen) == len(set(seen))  # no vertex twice
    assert set(seen) == want            # and none missing

    def signed_area(p):
        return 0.5 * float(np.sum(p[:-1, 0] * p[1:, 1] - p[1:, 0] * p[:-1, 1]))

    island = np.zeros((12, 12), np.uint8)
    island[3:9, 3:9] = 1
    ring = island.copy()
    ring[5:7, 5:7] = 0
    (outer,) = oracle_outlines.find_contours(island, 0.5)
    both = oracle_outlines.find_contours(ring, 0.5)
    assert len(both) == 2
    areas = sorted(signed_area(c) for c in both)
    assert areas[0] * areas[1] < 0                                   # opposite windings
    assert np.sign(signed_area(outer)) == np.sign(signed_area(max(both, key=len)))  # island and ring: same outer winding
    assert abs(abs(signed_area(outer)) - 35.5) < 1e-9               # 6x6 pixels minus the four cut corners (4 x 1/8)
