"""Regenerate tests/golden/config1_thresholds_outlines.npz: the oracle's answers for the SURVEY 8f rank-4
rows (all ten ``apply_threshold`` methods, both outline extractors) on the reference's own fixture
(config 1, already stored in config1_multichannel.npz, so this script needs no /root/reference).

    python tests/golden/make_golden_rank4.py
"""

from __future__ import annotations

import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402
from oracle import outlines, threshold  # noqa: E402

HERE = Path(__file__).parent
GLOBAL_METHODS = ["otsu", "li", "yen", "isodata", "mean", "minimum", "triangle"]
LOCAL_CASES = [("niblack", {}), ("niblack", {"window_size": 31, "k": 0.1}), ("sauvola", {}),
               ("sauvola", {"window_size": (7, 25), "k": 0.3}), ("local", {"block_size": 35}),
               ("local", {"block_size": 51, "offset": -15.0})]


def mask_digest(mask: np.ndarray) -> str:
    return hashlib.sha256(np.packbits(np.ascontiguousarray(mask)).tobytes()).hexdigest()


def outline_digest(items: list[np.ndarray]) -> str:
    h = hashlib.sha256()
    for a in items:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


def threshold_value(image: np.ndarray, method: str) -> float:
    func = getattr(threshold, f"threshold_{method}")
    try:
        return float(func(image.copy()))
    except RuntimeError:  # threshold_minimum: "Unable to find two maxima in histogram"
        return float("nan")


def main() -> None:
    with np.load(HERE / "config1_multichannel.npz") as z:
        fov = z["fov"]
        labels_thr = z["bg0/labels_thr"].astype(np.int64)
        labels_given = z["bg0/labels_given"].astype(np.int64)
    gold: dict[str, np.ndarray] = {}
    pre = oracle.rescale_by_percentile(oracle.subtract_background_dog(fov[1], 0.6, 16.0, percentile=0), (1, 99))
    for name, image, methods in (("dapi_u16", fov[1], GLOBAL_METHODS), ("fitc_u16", fov[2], GLOBAL_METHODS),
                                 ("dapi_pre_f64", pre, ["otsu", "yen", "isodata", "minimum", "triangle"])):
        values = np.array([threshold_value(image, m) for m in methods])
        gold[f"{name}/methods"] = np.array(methods)
        gold[f"{name}/thresholds"] = values
        gold[f"{name}/foreground"] = np.array([int((image > v).sum()) if np.isfinite(v) else -1 for v in values])
    for i, (method, kw) in enumerate(LOCAL_CASES):
        mask = oracle.apply_threshold(fov[1], method, **kw)
        gold[f"local/{i}/foreground"] = np.array(int(mask.sum()))
        gold[f"local/{i}/sha256"] = np.array(mask_digest(mask))
    for name, lab in (("thr", labels_thr), ("given", labels_given)):
        for extractor, func in (("cellpose", outlines.extract_outlines_cellpose), ("skimage", outlines.extract_outlines_skimage)):
            items = func(lab)
            assert len(items) == lab.max()
            gold[f"outlines/{name}/{extractor}/points"] = np.array([len(a) for a in items])
            gold[f"outlines/{name}/{extractor}/sha256"] = np.array(outline_digest(items))
    out = HERE / "config1_thresholds_outlines.npz"
    np.savez_compressed(out, **gold)
    print(f"wrote {out} ({out.stat().st_size / 1024:.1f} KiB, {len(gold)} arrays)")
    for k in ("dapi_u16", "fitc_u16", "dapi_pre_f64"):
        print(k, dict(zip(gold[f"{k}/methods"].tolist(), gold[f"{k}/thresholds"].tolist())))
    print({k: v.tolist() for k, v in gold.items() if k.endswith("/points")})


if __name__ == "__main__":
    main()
