"""GPU parity at the NAMED shapes of BASELINE.json configs 2, 3 and 4 (the small-shape files next to this one run the
same comparisons where the oracle needs seconds; these need about a minute of CPU each).

config 2: one full 4 x 2048 x 2048 field of view + its ~2 k-cell int64 label mask through the executor (host-fed C-ABI
          call and device-resident call) against the oracle: thresholds, labels, counts, integer columns bit for bit,
          float columns to rtol 1e-5, the thresholded channel's plane bit for bit, the others to 1e-8 (bench.py runs the
          same check on three FOVs and prints it as `parity_checked`).
config 3: one 2 x 2048 x 2048 frame pair through Pipeline(parallel=True), preprocess + Otsu + labels per frame.
config 4: per-slice preprocessing of a 4 x 1024 x 1024 slice set, and the 3-D per-object table of a 64 x 1024 x 1024
          label volume over four channels."""

from __future__ import annotations

import numpy as np
import pytest

import oracle
from oracle.regionprops import regionprops_table_3d

pytestmark = pytest.mark.gpu

import bench  # noqa: E402
from arcadia_microscopy_tools_b200 import operations  # noqa: E402
from arcadia_microscopy_tools_b200.batch import FovBatchExecutor, FovPipelineConfig  # noqa: E402
from arcadia_microscopy_tools_b200.channels import DAPI  # noqa: E402
from arcadia_microscopy_tools_b200.masks import SegmentationMask  # noqa: E402
from arcadia_microscopy_tools_b200.pipeline import ImageOperation, Pipeline  # noqa: E402
from arcadia_microscopy_tools_b200.synthetic import make_fov  # noqa: E402
from arcadia_microscopy_tools_b200.volumes import quantify_label_volume  # noqa: E402


@pytest.mark.parametrize("plane_filter", ["tensor_core", "fma"])
def test_config2_full_fov_against_oracle(plane_filter):
    _, batch, results = bench.cpu_baseline_single(1)
    cfg = FovPipelineConfig(n_channels=4, height=2048, width=2048, seg_channel=bench.SEG_CHANNEL, chunk_fovs=1,
                            max_labels=4096, max_label_value=int(batch[0][1].max()), given_label_dtype=np.int64,
                            plane_filter=plane_filter)
    with FovBatchExecutor(cfg) as ex:
        assert ex.uses_tensor_cores == (plane_filter == "tensor_core")
        report = bench.parity_check(ex, batch, results, [n.upper() for n in bench.NAMES])
    assert report["ok"], report["failures"]
    assert report["max_plane_abs_err_other_channels"] <= (bench.PLANE_ATOL_TENSOR_CORE if plane_filter == "tensor_core" else 1e-12)
    assert len(results[0]["props_thr"]["label"]) > 500 and len(results[0]["props_given"]["label"]) > 1500


def test_config3_full_size_frame_pair():
    fov, _, _ = make_fov(20263000, 2, 2048, 2048, 2000)
    frames = fov  # (2, 2048, 2048): the two channels of one time point, independent slices of a parallel pipeline
    pre = Pipeline([ImageOperation(operations.subtract_background_dog, low_sigma=0.6, high_sigma=16.0, percentile=0),
                    ImageOperation(operations.rescale_by_percentile, percentile_range=(1, 99), out_range=(0, 1))],
                   parallel=True)
    seg = Pipeline(pre.operations + [ImageOperation(operations.apply_threshold)], parallel=True)
    got_pre, got_seg = pre(frames), seg(frames)
    for c in range(2):
        want = oracle.rescale_by_percentile(oracle.subtract_background_dog(frames[c], 0.6, 16.0, 0), (1, 99), (0, 1))
        assert np.array_equal(got_pre[c], want), c
        want_mask = oracle.apply_threshold(want)
        assert np.array_equal(got_seg[c], want_mask), c
    labels = SegmentationMask(got_seg[1], {DAPI: frames[1]}, remove_edge_cells=True).label_image
    assert np.array_equal(labels, oracle.process_mask(oracle.apply_threshold(
        oracle.rescale_by_percentile(oracle.subtract_background_dog(frames[1], 0.6, 16.0, 0), (1, 99), (0, 1))), True))


def test_config4_full_size_slices_and_volume():
    Z, C, H, W = 64, 4, 1024, 1024
    slice_set, _, _ = make_fov(20264000, C, H, W, 500)  # one z position, four channels
    pre = Pipeline([ImageOperation(operations.subtract_background_dog, low_sigma=0.6, high_sigma=16.0, percentile=0),
                    ImageOperation(operations.rescale_by_percentile, percentile_range=(1, 99))], parallel=True)
    got = pre(slice_set)
    for c in range(C):
        want = oracle.rescale_by_percentile(oracle.subtract_background_dog(slice_set[c], 0.6, 16.0, 0), (1, 99))
        assert np.array_equal(got[c], want), c
    # the 3-D label volume: ellipsoids spanning several slices, four raw channels
    rng = np.random.default_rng(44)
    vol = np.zeros((Z, H, W), dtype=np.int32)
    n_obj = 400
    for i in range(n_obj):
        ctr = rng.uniform([2, 20, 20], [Z - 2, H - 20, W - 20])
        rad = rng.uniform([1.5, 6, 6], [4.0, 16, 16])
        z0, z1 = int(max(ctr[0] - rad[0], 0)), int(min(ctr[0] + rad[0] + 1, Z))
        y0, y1 = int(ctr[1] - rad[1]), int(ctr[1] + rad[1] + 1)
        x0, x1 = int(ctr[2] - rad[2]), int(ctr[2] + rad[2] + 1)
        zz, yy, xx = np.mgrid[z0:z1, y0:y1, x0:x1]
        m = ((zz - ctr[0]) / rad[0]) ** 2 + ((yy - ctr[1]) / rad[1]) ** 2 + ((xx - ctr[2]) / rad[2]) ** 2 <= 1
        sub = vol[z0:z1, y0:y1, x0:x1]
        sub[m & (sub == 0)] = i + 1
    uniq = np.unique(vol[vol > 0])
    lut = np.zeros(int(vol.max()) + 1, dtype=np.int64)
    lut[uniq] = np.arange(1, uniq.size + 1)
    names = ["brightfield", "dapi", "fitc", "tritc"]
    chans = {n: rng.integers(100, 4000, size=(Z, H, W)).astype(np.uint16) for n in names}
    got = quantify_label_volume(vol, chans)
    want = regionprops_table_3d(lut[vol], chans)
    rename = {"centroid-0": "centroid_z", "centroid-1": "centroid_y", "centroid-2": "centroid_x"}
    assert len(want["label"]) > 300
    for key, w in want.items():
        g = got[rename.get(key, key)]
        if key in ("label", "area") or key.startswith(("bbox", "intensity_sum", "intensity_max", "intensity_min")):
            assert np.array_equal(g.astype(np.float64), np.asarray(w, dtype=np.float64)), key
        else:
            assert np.allclose(g, w, rtol=1e-5, atol=1e-9 * max(1.0, float(np.abs(w).max()))), key
