"""Golden vectors for the SURVEY 8f rank-4 rows on the reference's own fixture (config 1):
``tests/golden/config1_thresholds_outlines.npz`` (made by ``tests/golden/make_golden_rank4.py``).
CPU: the oracle reproduces them; GPU (``-m gpu``): the product reproduces them through its public API."""

from __future__ import annotations

import hashlib
from pathlib import Path

import numpy as np
import pytest

import oracle
from oracle import outlines as oracle_outlines
from oracle import threshold

GOLDEN4 = Path(__file__).parent / "golden" / "config1_thresholds_outlines.npz"
LOCAL_CASES = [("niblack", {}), ("niblack", {"window_size": 31, "k": 0.1}), ("sauvola", {}),
               ("sauvola", {"window_size": (7, 25), "k": 0.3}), ("local", {"block_size": 35}),
               ("local", {"block_size": 51, "offset": -15.0})]


@pytest.fixture(scope="module")
def golden4():
    with np.load(GOLDEN4) as z:
        return {k: z[k] for k in z.files}


def _mask_digest(mask) -> str:
    return hashlib.sha256(np.packbits(np.ascontiguousarray(mask)).tobytes()).hexdigest()


def _outline_digest(items) -> str:
    h = hashlib.sha256()
    for a in items:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


def _images(golden):
    fov = golden["fov"]
    pre = oracle.rescale_by_percentile(oracle.subtract_background_dog(fov[1], 0.6, 16.0, percentile=0), (1, 99))
    return {"dapi_u16": fov[1], "fitc_u16": fov[2], "dapi_pre_f64": pre}


def test_known_threshold_values_of_the_fixture(golden4):
    """Independent anchors: Otsu of raw DAPI / FITC and of the preprocessed plane are the SURVEY 8c values;
    ISODATA agrees with Otsu on these bimodal histograms; the raw histograms have no second mode (minimum fails)."""
    d = dict(zip(golden4["dapi_u16/methods"].tolist(), golden4["dapi_u16/thresholds"].tolist()))
    f = dict(zip(golden4["fitc_u16/methods"].tolist(), golden4["fitc_u16/thresholds"].tolist()))
    p = dict(zip(golden4["dapi_pre_f64/methods"].tolist(), golden4["dapi_pre_f64/thresholds"].tolist()))
    assert d["otsu"] == 2742 and f["otsu"] == 968 and p["otsu"] == 0.490234375
    assert d["isodata"] == 2742 and f["isodata"] == 968
    assert np.isnan(d["minimum"]) and np.isnan(f["minimum"]) and p["minimum"] == 0.048828125
    assert d["mean"] == 100104950 / 65536 and f["mean"] == 31846305 / 65536  # channel sums of SURVEY 8c


def test_oracle_reproduces_threshold_goldens(golden, golden4):
    for name, image in _images(golden).items():
        methods = golden4[f"{name}/methods"].tolist()
        for m, want, fg in zip(methods, golden4[f"{name}/thresholds"], golden4[f"{name}/foreground"]):
            if np.isnan(want):
                with pytest.raises(RuntimeError):
                    getattr(threshold, f"threshold_{m}")(image.copy())
                continue
            assert float(getattr(threshold, f"threshold_{m}")(image.copy())) == want, (name, m)
            assert int(oracle.apply_threshold(image.copy(), m).sum()) == fg, (name, m)
    for i, (method, kw) in enumerate(LOCAL_CASES):
        mask = oracle.apply_threshold(golden["fov"][1], method, **kw)
        assert _mask_digest(mask) == str(golden4[f"local/{i}/sha256"]) and mask.sum() == golden4[f"local/{i}/foreground"]


def test_oracle_reproduces_outline_goldens(golden, golden4):
    for name in ("thr", "given"):
        lab = golden[f"bg0/labels_{name}"].astype(np.int64)
        for extractor, func in (("cellpose", oracle_outlines.extract_outlines_cellpose),
                                ("skimage", oracle_outlines.extract_outlines_skimage)):
            items = func(lab)
            assert [len(a) for a in items] == golden4[f"outlines/{name}/{extractor}/points"].tolist()
            assert _outline_digest(items) == str(golden4[f"outlines/{name}/{extractor}/sha256"])


@pytest.mark.gpu
def test_gpu_thresholds_reproduce_goldens(golden, golden4):
    from arcadia_microscopy_tools_b200 import operations

    for name, image in _images(golden).items():
        if name == "dapi_pre_f64":  # the product's own preprocessing chain (bit-identical to the oracle's)
            image = operations.rescale_by_percentile(operations.subtract_background_dog(golden["fov"][1], 0.6, 16.0, percentile=0), (1, 99))
        methods = golden4[f"{name}/methods"].tolist()
        for m, want, fg in zip(methods, golden4[f"{name}/thresholds"], golden4[f"{name}/foreground"]):
            if np.isnan(want):
                with pytest.raises(RuntimeError, match="Unable to find two maxima"):
                    operations.apply_threshold(image, m)
                continue
            mask = operations.apply_threshold(image, m)
            assert mask.sum() == fg and np.array_equal(mask, np.asarray(image) > want), (name, m)
    for i, (method, kw) in enumerate(LOCAL_CASES):
        mask = operations.apply_threshold(golden["fov"][1], method, **kw)
        assert _mask_digest(mask) == str(golden4[f"local/{i}/sha256"]), (method, kw)


@pytest.mark.gpu
def test_gpu_outlines_reproduce_goldens(golden, golden4):
    from arcadia_microscopy_tools_b200.masks import SegmentationMask

    for name in ("thr", "given"):
        lab = golden[f"bg0/labels_{name}"].astype(np.int64)
        for extractor in ("cellpose", "skimage"):
            items = SegmentationMask(lab, remove_edge_cells=False, outline_extractor=extractor).cell_outlines
            assert [len(a) for a in items] == golden4[f"outlines/{name}/{extractor}/points"].tolist(), (name, extractor)
            assert _outline_digest(items) == str(golden4[f"outlines/{name}/{extractor}/sha256"]), (name, extractor)
