"""GPU parity on the shapes of BASELINE.json configs 3 and 4 (scaled down so the oracle finishes in
seconds): a (T, C, Y, X) time-lapse preprocessed and labelled frame by frame through
``MicroscopyImage.apply_pipeline`` / ``Pipeline(parallel=True)``, and a (Z, C, Y, X) confocal
stack with per-slice filters plus 3-D per-label quantification."""

from __future__ import annotations

import numpy as np
import pytest

import oracle
from oracle.regionprops import regionprops_table_3d

pytestmark = pytest.mark.gpu

from arcadia_microscopy_tools_b200 import operations  # noqa: E402
from arcadia_microscopy_tools_b200.channels import DAPI, FITC  # noqa: E402
from arcadia_microscopy_tools_b200.masks import SegmentationMask  # noqa: E402
from arcadia_microscopy_tools_b200.microscopy import MicroscopyImage  # noqa: E402
from arcadia_microscopy_tools_b200.pipeline import ImageOperation, Pipeline  # noqa: E402
from arcadia_microscopy_tools_b200.synthetic import make_fov  # noqa: E402
from arcadia_microscopy_tools_b200.volumes import quantify_label_volume  # noqa: E402


def _stack(n, c, shape, seed):
    frames = [make_fov(seed + i, c, shape[0], shape[1], 25)[0] for i in range(n)]
    return np.stack(frames)  # (n, C, Y, X)


def test_config3_timelapse_per_frame_preprocess_and_label():
    """T x 2 channels (Ti2-E-shaped): every frame of both channels through DoG + percentile
    rescale, Otsu on each frame, labels per frame -- the non-contiguous (T, Y, X) channel view of
    a (T, C, Y, X) block goes straight into Pipeline(parallel=True)."""
    T, shape = 6, (128, 160)
    data = _stack(T, 2, shape, 300)
    image = MicroscopyImage.from_arrays(data, [DAPI, FITC], "TCYX")
    pre = Pipeline([ImageOperation(operations.subtract_background_dog, low_sigma=0.6, high_sigma=16.0, percentile=0),
                    ImageOperation(operations.rescale_by_percentile, percentile_range=(1, 99), out_range=(0, 1))],
                   parallel=True)
    seg = Pipeline(pre.operations + [ImageOperation(operations.apply_threshold)], parallel=True)
    for ch_index, channel in enumerate((DAPI, FITC)):
        view = image.get_channel_intensities(channel)
        assert view.shape == (T, *shape) and not view.flags.c_contiguous
        got_pre = image.apply_pipeline(pre, channel)
        got_seg = image.apply_pipeline(seg, channel)
        for t in range(T):
            want = oracle.rescale_by_percentile(oracle.subtract_background_dog(data[t, ch_index], 0.6, 16.0, 0), (1, 99), (0, 1))
            assert np.array_equal(got_pre[t], want), (channel.name, t)
            want_mask = oracle.apply_threshold(want)
            assert np.array_equal(got_seg[t], want_mask), (channel.name, t)
            if ch_index == 0:
                labels = SegmentationMask(got_seg[t], {DAPI: data[t, 0]}, remove_edge_cells=True).label_image
                assert np.array_equal(labels, oracle.process_mask(want_mask, True)), t


def test_config4_zstack_per_slice_filters_and_3d_quantification():
    """Z x 4 channels (Stellaris-shaped): per-slice preprocessing of every channel, per-slice Otsu,
    then the per-object table of a 3-D label volume over all four raw channels."""
    Z, C, shape = 8, 4, (96, 96)
    data = _stack(Z, C, shape, 400)
    names = ["brightfield", "dapi", "fitc", "tritc"]
    pre = Pipeline([ImageOperation(operations.subtract_background_dog, low_sigma=1.0, high_sigma=8.0, percentile=10),
                    ImageOperation(operations.rescale_by_percentile, percentile_range=(0.5, 99.5))], parallel=True)
    for c in range(C):
        got = pre(data[:, c])
        for z in range(Z):
            want = oracle.rescale_by_percentile(oracle.subtract_background_dog(data[z, c], 1.0, 8.0, 10), (0.5, 99.5))
            assert np.array_equal(got[z], want), (c, z)
    # 3-D objects: ellipsoids spanning several slices
    rng = np.random.default_rng(41)
    zz, yy, xx = np.mgrid[0:Z, 0:shape[0], 0:shape[1]]
    vol = np.zeros((Z, *shape), dtype=np.int32)
    for i in range(30):
        ctr = rng.uniform([0, 0, 0], [Z, *shape])
        rad = rng.uniform([1.0, 3, 3], [3.0, 8, 8])
        m = ((zz - ctr[0]) / rad[0]) ** 2 + ((yy - ctr[1]) / rad[1]) ** 2 + ((xx - ctr[2]) / rad[2]) ** 2 <= 1
        vol[m & (vol == 0)] = i + 1
    uniq = np.unique(vol[vol > 0])
    lut = np.zeros(int(vol.max()) + 1, dtype=np.int64)
    lut[uniq] = np.arange(1, uniq.size + 1)
    chans = {n: np.ascontiguousarray(data[:, c]) for c, n in enumerate(names)}
    got = quantify_label_volume(vol, chans)
    want = regionprops_table_3d(lut[vol], chans)
    rename = {"centroid-0": "centroid_z", "centroid-1": "centroid_y", "centroid-2": "centroid_x"}
    for key, w in want.items():
        g = got[rename.get(key, key)]
        if key in ("label", "area") or key.startswith(("bbox", "intensity_sum", "intensity_max", "intensity_min")):
            assert np.array_equal(g.astype(np.float64), np.asarray(w, dtype=np.float64)), key
        else:
            assert np.allclose(g, w, rtol=1e-5, atol=1e-9 * max(1.0, float(np.abs(w).max()))), key
