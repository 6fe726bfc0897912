"""The committed golden fixture (reference test data, config 1) against the oracle."""

from __future__ import annotations

import hashlib

import numpy as np

import oracle
from oracle import threshold


def test_fixture_is_the_reference_nd2_pixels(golden):
    fov = golden["fov"]
    assert fov.shape == (4, 256, 256) and fov.dtype == np.uint16
    digest = hashlib.sha256(fov.tobytes()).hexdigest()
    assert digest == str(golden["fov_sha256"]) and digest.startswith("c9d2dd8f") and digest.endswith("2d1ff6dd")
    stats = [(int(c.min()), int(c.max()), int(c.sum())) for c in fov]
    assert stats == [(7370, 17636, 845349334), (1048, 16117, 100104950), (129, 2943, 31846305), (98, 777, 11379399)]


def test_raw_otsu_known_answers(golden):
    fov = golden["fov"]
    ts = [int(threshold.threshold_otsu(fov[c])) for c in range(4)]
    assert ts == golden["raw_otsu"].tolist() == [12407, 2742, 968, 262]
    assert [int((fov[c] > ts[c]).sum()) for c in range(4)] == [53179, 1297, 3715, 1338]


def test_workload_known_answers_dapi(golden):
    dapi = golden["fov"][1]
    x = oracle.subtract_background_dog(dapi, 0.6, 16.0, percentile=0)
    P = oracle.rescale_by_percentile(x, (1, 99), (0, 1))
    assert hashlib.sha256(np.ascontiguousarray(P).tobytes()).hexdigest() == str(golden["bg0/pre_sha256"][1])
    mask = oracle.apply_threshold(P)
    labels = oracle.process_mask(mask, True)
    assert np.array_equal(labels, golden["bg0/labels_thr"])
    assert labels.max() == 26 and float(threshold.threshold_otsu(P)) == 0.490234375
    props = oracle.cell_properties(labels, {"fitc": golden["fov"][2]}, ["label", "area"], ["intensity_sum"])
    assert props["area"][:5].tolist() == [38, 483, 28, 4, 6]
    assert props["intensity_sum_fitc"][:3].tolist() == [43134, 187370, 27834]
