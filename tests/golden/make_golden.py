"""Regenerate tests/golden/*.npz (run in the build container, where /root/reference exists).

Inputs: the reference's own test fixture ``tests/data/example-multichannel.nd2`` (config 1 of
BASELINE.json), decoded with the package's raw ND2 reader.  Outputs: the pixel block itself
(so the GPU box, which has no /root/reference, can run config 1) and the oracle's answers for
workload W on it.  The oracle is pinned against scipy / numpy in tests/test_oracle.py and
reproduces every provisional known answer of SURVEY.md 8c (asserted below).

    python tests/golden/make_golden.py
"""

from __future__ import annotations

import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402
from arcadia_microscopy_tools_b200.nd2_raw import read_nd2  # noqa: E402
from oracle import filters, regionprops, threshold  # noqa: E402

REF_DATA = Path("/root/reference/src/arcadia_microscopy_tools/tests/data")
CHANNELS = ["BRIGHTFIELD", "DAPI", "FITC", "TRITC"]
TABLE_PROPS = ["label", "area", "bbox", "centroid", "axis_major_length", "axis_minor_length", "eccentricity",
               "orientation", "perimeter", "area_convex"]
INT_PROPS = ["intensity_sum", "intensity_mean", "intensity_max", "intensity_min", "intensity_std"]


def workload(fov: np.ndarray, seg: int, bg_pct: float, given: np.ndarray | None = None) -> dict:
    out: dict[str, np.ndarray] = {}
    lv, p1, p2, sha = [], [], [], []
    pre = []
    for c in range(fov.shape[0]):
        dog = filters.difference_of_gaussians(fov[c], 0.6, 16.0)
        level = np.percentile(dog, bg_pct)
        clipped = np.clip(dog - level, 0, None)
        a, b = np.percentile(clipped, (1, 99))
        P = oracle.rescale_by_percentile(clipped, (1, 99), (0, 1))
        lv.append(level), p1.append(a), p2.append(b)
        sha.append(hashlib.sha256(np.ascontiguousarray(P).tobytes()).hexdigest())
        pre.append(P)
    out["levels"], out["p1"], out["p2"] = np.array(lv), np.array(p1), np.array(p2)
    out["pre_sha256"] = np.array(sha)
    out["threshold"] = np.array(threshold.threshold_otsu(pre[seg]))
    mask = oracle.apply_threshold(pre[seg])
    labels = oracle.process_mask(mask, True)
    out["labels_thr"] = labels.astype(np.int32)

    def tables(lab, prefix):
        morph = regionprops.regionprops_table(lab, properties=TABLE_PROPS)
        for k, v in morph.items():
            out[f"{prefix}/{k}"] = v
        for c, name in enumerate(CHANNELS[: fov.shape[0]]):
            t = regionprops.regionprops_table(lab, intensity_image=fov[c], properties=INT_PROPS)
            for k, v in t.items():
                out[f"{prefix}/{k}_{name.lower()}"] = v

    tables(labels, "thr")
    if given is not None:
        lab_g = oracle.process_mask(given.astype(np.int64), True)
        out["labels_given"] = lab_g.astype(np.int32)
        tables(lab_g, "given")
    return out


def main() -> None:
    fov = read_nd2(REF_DATA / "example-multichannel.nd2")
    assert fov.shape == (4, 256, 256) and fov.dtype == np.uint16
    digest = hashlib.sha256(fov.tobytes()).hexdigest()
    assert digest.startswith("c9d2dd8f") and digest.endswith("2d1ff6dd"), digest
    gold: dict[str, np.ndarray] = {"fov": fov, "fov_sha256": np.array(digest)}
    # raw-image Otsu (exact-bin) per channel: thresholds, foreground, components
    raw_t, raw_fg, raw_nc, raw_nc_cb = [], [], [], []
    for c in range(4):
        t = threshold.threshold_otsu(fov[c])
        m = fov[c] > t
        raw_t.append(int(t)), raw_fg.append(int(m.sum()))
        raw_nc.append(int(oracle.labeling.label(m, return_num=True)[1]))
        raw_nc_cb.append(int(oracle.labeling.label(oracle.labeling.clear_border(m), return_num=True)[1]))
    assert raw_t == [12407, 2742, 968, 262] and raw_fg == [53179, 1297, 3715, 1338], (raw_t, raw_fg)
    assert raw_nc == [53, 20, 21, 111] and raw_nc_cb == [47, 20, 21, 110], (raw_nc, raw_nc_cb)
    gold["raw_otsu"], gold["raw_fg"] = np.array(raw_t), np.array(raw_fg)
    gold["raw_ncomp"], gold["raw_ncomp_cleared"] = np.array(raw_nc), np.array(raw_nc_cb)
    # a Cellpose-like integer mask for config 1: the raw-DAPI Otsu components dilated by value
    from scipy import ndimage as ndi

    seeds = oracle.labeling.label(fov[1] > raw_t[1])
    given = ndi.grey_dilation(seeds, size=(5, 5)).astype(np.int32)
    gold["given"] = given
    for bg in (0, 90):
        w = workload(fov, seg=1, bg_pct=bg, given=given)
        for k, v in w.items():
            gold[f"bg{bg}/{k}"] = v
    w0 = {k[4:]: v for k, v in gold.items() if k.startswith("bg0/")}
    assert repr(float(w0["levels"][1])) == "-0.010282287478435684"
    assert repr(float(w0["p1"][1])) == "0.003493070777886204" and repr(float(w0["p2"][1])) == "0.034058555701019226"
    assert float(w0["threshold"]) == 0.490234375 and int(w0["labels_thr"].max()) == 26
    assert w0["thr/area"][:5].tolist() == [38, 483, 28, 4, 6] and w0["thr/area"].sum() == 1945
    assert w0["thr/intensity_sum_fitc"][:3].tolist() == [43134, 187370, 27834]
    w90 = {k[5:]: v for k, v in gold.items() if k.startswith("bg90/")}
    assert repr(float(w90["levels"][1])) == "0.0011692927313822956" and float(w90["threshold"]) == 0.392578125
    assert int(w90["labels_thr"].max()) == 23 and w90["thr/area"][:5].tolist() == [22, 433, 13, 2, 107]
    out = Path(__file__).with_name("config1_multichannel.npz")
    np.savez_compressed(out, **gold)
    print(f"wrote {out} ({out.stat().st_size / 1024:.0f} KiB, {len(gold)} arrays)")


if __name__ == "__main__":
    main()
