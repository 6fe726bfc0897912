"""Oracle: connected-component labelling, clear_border, relabel_sequential.
TEST INFRASTRUCTURE ONLY.

Restates the scikit-image calls of the reference's ``_process_mask``
(``src/arcadia_microscopy_tools/masks.py:38-65``): ``ski.segmentation.clear_border`` (:56),
``ski.measure.label`` (:63), ``ski.segmentation.relabel_sequential`` (:65).  SURVEY.md 8a
items 7-9.
"""

from __future__ import annotations

import numpy as np
from scipy import ndimage as ndi
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components


def _label_bool(image: np.ndarray):
    """skimage ``_label_bool``: ``scipy.ndimage.label`` with the full (connectivity = ndim)
    structuring element; int32 output numbered in raster order of each component's first
    pixel."""
    structure = ndi.generate_binary_structure(image.ndim, image.ndim)
    lab, n = ndi.label(image, structure=structure)
    return lab, int(n)


def _label_multivalue(image: np.ndarray):
    """skimage ``clabel`` semantics for integer images: maximal regions of EQUAL non-zero
    value under full connectivity, background 0, numbered in raster order of first pixel.

    Built as connected components of the pixel graph whose edges join fully-connected
    neighbours holding the same non-zero value.
    """
    shape = image.shape
    n = image.size
    idx = np.arange(n, dtype=np.int64).reshape(shape)
    rows = []
    cols = []
    # half of the (3^ndim - 1) neighbour offsets is enough for an undirected graph
    offsets = [off for off in np.ndindex(*([3] * image.ndim)) if off > tuple([1] * image.ndim)]
    for off in offsets:
        d = [o - 1 for o in off]
        sl_a = tuple(slice(max(0, -k), s - max(0, k)) for k, s in zip(d, shape))
        sl_b = tuple(slice(max(0, k), s - max(0, -k)) for k, s in zip(d, shape))
        a = image[sl_a]
        b = image[sl_b]
        m = (a == b) & (a != 0)
        rows.append(idx[sl_a][m])
        cols.append(idx[sl_b][m])
    r = np.concatenate(rows)
    c = np.concatenate(cols)
    graph = coo_matrix((np.ones(r.size, dtype=np.int8), (r, c)), shape=(n, n)).tocsr()
    _, comp = connected_components(graph, directed=False)
    comp = comp.reshape(shape)
    fg = image != 0
    out = np.zeros(shape, dtype=np.int64)
    if fg.any():
        comp_fg = comp[fg]
        uniq, first = np.unique(comp_fg, return_index=True)
        order = np.argsort(first, kind="stable")  # raster order of first pixel
        rank = np.empty(uniq.size, dtype=np.int64)
        rank[order] = np.arange(1, uniq.size + 1)
        out[fg] = rank[np.searchsorted(uniq, comp_fg)]
    return out, int(out.max())


def label(image: np.ndarray, return_num: bool = False):
    """``skimage.measure.label(image)`` with default arguments (background 0, full
    connectivity)."""
    image = np.asarray(image)
    if image.dtype == np.bool_:
        lab, n = _label_bool(image)
    else:
        lab, n = _label_multivalue(image)
    return (lab, n) if return_num else lab


def clear_border(labels: np.ndarray) -> np.ndarray:
    """``skimage.segmentation.clear_border(labels)`` (buffer_size 0, bgval 0): re-label by
    (value, connectivity), drop every re-labelled component with a pixel on the image border,
    keep the input dtype (bool stays bool)."""
    labels = np.asarray(labels)
    out = labels.copy()
    borders = np.zeros(out.shape, dtype=bool)
    for d in range(out.ndim):
        sl = [slice(None)] * out.ndim
        sl[d] = slice(0, 1)
        borders[tuple(sl)] = True
        sl[d] = slice(-1, None)
        borders[tuple(sl)] = True
    lab, number = label(out, return_num=True)
    borders_indices = np.unique(lab[borders])
    label_mask = np.isin(np.arange(number + 1), borders_indices)
    mask = label_mask[lab.reshape(-1)].reshape(lab.shape)
    out[mask] = 0
    return out


def relabel_sequential(label_field: np.ndarray) -> np.ndarray:
    """``skimage.segmentation.relabel_sequential(label_field)[0]``: 0 -> 0 and the i-th
    smallest non-zero value -> i."""
    label_field = np.asarray(label_field)
    in_vals = np.unique(label_field)
    nz = in_vals[in_vals != 0]
    out = np.searchsorted(nz, label_field) + 1
    out[label_field == 0] = 0
    return out.astype(label_field.dtype if label_field.dtype.kind in "iu" else np.int64)
