"""GPU parity: 3-D per-object quantification (BASELINE config 4) against the oracle's restatement of
skimage's 3-D regionprops (integers bit-exact, floats within 1e-5 relative)."""

from __future__ import annotations

import numpy as np
import pytest

from oracle.regionprops import regionprops_table_3d

pytestmark = pytest.mark.gpu

from arcadia_microscopy_tools_b200.volumes import quantify_label_volume, relabel_volume  # noqa: E402


def ellipsoid_volume(shape, n_obj, rng):
    """Non-overlapping labelled ellipsoids (some clipped by the volume border), labels not consecutive."""
    d, h, w = shape
    vol = np.zeros(shape, dtype=np.int32)
    zz, yy, xx = np.mgrid[0:d, 0:h, 0:w]
    for i in range(n_obj):
        c = rng.uniform([0, 0, 0], [d, h, w])
        r = rng.uniform([1.5, 3, 3], [4, 9, 9])
        m = ((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2 <= 1
        m &= vol == 0
        vol[m] = 3 * i + 2
    return vol


@pytest.mark.parametrize("shape,n_obj", [((12, 96, 80), 25), ((5, 40, 131), 10), ((33, 64, 64), 40)])
def test_quantify_label_volume_matches_oracle(shape, n_obj):
    rng = np.random.default_rng(shape[0])
    vol = ellipsoid_volume(shape, n_obj, rng)
    chans = {"dapi": rng.integers(0, 65535, size=shape).astype(np.uint16),
             "fitc": (rng.gamma(2.0, 900.0, size=shape)).clip(0, 65535).astype(np.uint16)}
    got = quantify_label_volume(vol, chans)
    # oracle on the sequentially relabelled volume (what regionprops sees after relabel_sequential)
    uniq = np.unique(vol[vol > 0])
    lut = np.zeros(int(vol.max()) + 1, dtype=np.int64)
    lut[uniq] = np.arange(1, uniq.size + 1)
    seq = lut[vol]
    relabeled, k = relabel_volume(vol)
    assert k == uniq.size and np.array_equal(relabeled.cpu().numpy(), seq)
    want = regionprops_table_3d(seq, chans)
    assert len(got["label"]) == uniq.size
    rename = {"centroid-0": "centroid_z", "centroid-1": "centroid_y", "centroid-2": "centroid_x"}
    for key, w in want.items():
        g = got[rename.get(key, key)]
        if key in ("label", "area") or key.startswith(("bbox", "intensity_sum", "intensity_max", "intensity_min")):
            assert np.array_equal(g.astype(np.float64), np.asarray(w, dtype=np.float64)), key
        else:
            atol = 1e-9 * max(1.0, float(np.abs(w).max()))
            assert np.allclose(g, w, rtol=1e-5, atol=atol), (key, np.abs(g - w).max())


def test_quantify_label_volume_no_channels_and_validation():
    vol = np.zeros((4, 16, 16), dtype=np.int64)
    vol[1:3, 2:6, 3:9] = 7
    got = quantify_label_volume(vol)
    assert got["label"].tolist() == [1] and got["area"].tolist() == [48.0]
    assert [int(got[f"bbox-{i}"][0]) for i in range(6)] == [1, 2, 3, 3, 6, 9]
    assert np.allclose([got["centroid_z"][0], got["centroid_y"][0], got["centroid_x"][0]], [1.5, 3.5, 5.5])
    with pytest.raises(ValueError):
        quantify_label_volume(np.zeros((8, 8), dtype=np.int32))
    with pytest.raises(TypeError):
        quantify_label_volume(vol, {"dapi": np.zeros(vol.shape, dtype=np.float32)})
