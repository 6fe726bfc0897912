"""Pin the committed golden vectors to the REAL reference.

Needs an environment where scikit-image (the reference pins 0.25.2) imports — which is neither this build container
nor the GPU box (`profiles/r02_skimage_probe.txt`: no wheel in /opt/wheelhouse, no index reachable), so this script
has never run; it is committed so that the day such an environment exists one command settles the oracle's
"parity unpinned" legs:

    python tests/golden/make_golden_from_reference.py          # compare the reference with the committed goldens
    python tests/golden/make_golden_from_reference.py --write  # and rewrite them from the reference

It imports `operations` and `masks` straight from `/root/reference/src/arcadia_microscopy_tools/` (module files, not the
package `__init__`, which pulls matplotlib / cellpose), runs workload W of SURVEY.md 8d on the reference's own fixture
`tests/data/example-multichannel.nd2` exactly as `make_golden.py` runs it through the oracle, and reports every key
whose value differs from the committed `config1_multichannel.npz`.
"""

from __future__ import annotations

import hashlib
import importlib.util
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF_SRC = Path("/root/reference/src/arcadia_microscopy_tools")
GOLDEN = Path(__file__).parent / "config1_multichannel.npz"
CHANNELS = ["BRIGHTFIELD", "DAPI", "FITC", "TRITC"]
TABLE_PROPS = ["label", "area", "bbox", "centroid", "axis_major_length", "axis_minor_length", "eccentricity",
               "orientation", "perimeter", "area_convex"]
INT_PROPS = ["intensity_mean", "intensity_max", "intensity_min", "intensity_std"]


def _load(name: str):
    spec = importlib.util.spec_from_file_location(f"_ref_{name}", REF_SRC / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main() -> int:
    try:
        import skimage  # noqa: F401
        from skimage import measure
    except ImportError as err:  # the state of every environment this repository has seen so far
        print(f"scikit-image is not importable here ({err}); nothing compared, goldens untouched")
        return 0
    ops = _load("operations")
    masks = _load("masks")
    with np.load(GOLDEN) as z:
        gold = {k: z[k] for k in z.files}
    fov = gold["fov"]
    report: list[str] = []
    fresh: dict[str, np.ndarray] = {}
    for tag, bg in (("bg0", 0.0), ("bg90", 90.0)):
        pre = [ops.rescale_by_percentile(ops.subtract_background_dog(fov[c], 0.6, 16.0, bg), (1, 99), (0, 1)) for c in range(4)]
        fresh[f"{tag}/pre_sha256"] = np.array([hashlib.sha256(np.ascontiguousarray(p).tobytes()).hexdigest() for p in pre])
        from skimage import filters as skf

        fresh[f"{tag}/threshold"] = np.array(skf.threshold_otsu(pre[1]))
        labels = masks._process_mask(ops.apply_threshold(pre[1]), True)
        fresh[f"{tag}/labels_thr"] = labels.astype(np.int32)
        morph = measure.regionprops_table(labels, properties=TABLE_PROPS)
        for k, v in morph.items():
            fresh[f"{tag}/thr/{k}"] = v
        for c, name in enumerate(CHANNELS):
            t = measure.regionprops_table(labels, intensity_image=fov[c], properties=INT_PROPS)
            for k, v in t.items():
                fresh[f"{tag}/thr/{k}_{name.lower()}"] = v
    for key, val in fresh.items():
        if key not in gold:
            report.append(f"{key}: not in the committed golden file")
        elif val.dtype.kind in "fc":
            if not np.allclose(val, gold[key], rtol=1e-12, atol=0, equal_nan=True):
                report.append(f"{key}: max |d| = {np.nanmax(np.abs(val - gold[key])):.3e}")
        elif not np.array_equal(val, gold[key]):
            report.append(f"{key}: differs")
    print(f"{len(fresh)} keys compared with the real reference, {len(report)} differ")
    for line in report:
        print("  ", line)
    if "--write" in sys.argv:
        np.savez_compressed(GOLDEN, **(gold | fresh))
        print(f"rewrote {GOLDEN.name} from the reference")
    return 1 if report else 0


if __name__ == "__main__":
    sys.exit(main())
