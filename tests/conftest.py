"""Shared fixtures.  GPU tests are marked ``gpu``; everything else runs on the CPU box."""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
if str(Path(__file__).parent) not in sys.path:
    sys.path.insert(0, str(Path(__file__).parent))  # test helpers (nd2_synth)

GOLDEN = Path(__file__).parent / "golden" / "config1_multichannel.npz"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    with np.load(GOLDEN) as z:
        return {k: z[k] for k in z.files}


def draw_disk(center, radius, shape):
    """``skimage.draw.disk``: pixels with ((y-cy)/r)^2 + ((x-cx)/r)^2 < 1, clipped to shape."""
    cy, cx = center
    yy, xx = np.mgrid[: shape[0], : shape[1]]
    inside = ((yy - cy) / radius) ** 2 + ((xx - cx) / radius) ** 2 < 1
    return np.nonzero(inside)


def make_label_image(shape=(50, 50), cells=None):
    """Disk-shaped cells, as the reference's test fixture builder (test_masks.py:14-30)."""
    label_image = np.zeros(shape, dtype=np.int64)
    if cells is None:
        cells = [(shape[0] // 2, shape[1] // 2, 8)]
    for label, (cy, cx, r) in enumerate(cells, start=1):
        rr, cc = draw_disk((cy, cx), r, shape)
        label_image[rr, cc] = label
    return label_image


def random_blobs(seed, shape, n_blobs, rmax=9):
    """Random boolean mask of overlapping discs and squares (stress input for labelling)."""
    rng = np.random.default_rng(seed)
    m = np.zeros(shape, dtype=bool)
    for _ in range(n_blobs):
        cy, cx = rng.integers(0, shape[0]), rng.integers(0, shape[1])
        r = rng.integers(1, rmax)
        if rng.random() < 0.5:
            rr, cc = draw_disk((cy, cx), r, shape)
            m[rr, cc] = True
        else:
            m[max(cy - r, 0) : cy + r, max(cx - r, 0) : cx + r] = True
    return m
