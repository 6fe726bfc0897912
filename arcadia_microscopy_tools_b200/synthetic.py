"""Seeded synthetic fields of view for tests and the benchmark.

Per FOV: Poisson background per channel plus ~``n_cells`` non-overlapping elliptical cells with a
Gaussian intensity profile, clipped to uint16; and the matching Cellpose-like integer label
mask (pixel centre inside the ellipse, labels 1..K in generation order, a few cells touching
the border).  Host-side NumPy only: this is input generation, not part of the measured path.
"""

from __future__ import annotations

import numpy as np

BACKGROUND_LAMBDA = (300.0, 400.0, 250.0, 150.0)
CHANNEL_GAIN = (0.35, 1.0, 0.6, 0.25)


def _place_cells(rng: np.random.Generator, height: int, width: int, n_cells: int, r_lo: float, r_hi: float):
    """Rejection-sample non-overlapping centres on a coarse occupancy grid."""
    cell = int(2 * r_hi) + 1
    gh, gw = height // cell + 1, width // cell + 1
    grid: dict[tuple[int, int], list[tuple[float, float, float]]] = {}
    placed: list[tuple[float, float, float]] = []
    attempts = 0
    while len(placed) < n_cells and attempts < 40 * n_cells:
        attempts += 1
        r = rng.uniform(r_lo, r_hi)
        # ~5 % of the cells are allowed to straddle the image border
        margin = -0.5 * r if rng.random() < 0.05 else r + 1.0
        cy = rng.uniform(margin, height - 1 - margin)
        cx = rng.uniform(margin, width - 1 - margin)
        gy, gx = int(cy) // cell, int(cx) // cell
        ok = True
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                for (oy, ox, orad) in grid.get((gy + dy, gx + dx), ()):
                    if (oy - cy) ** 2 + (ox - cx) ** 2 < (r + orad + 1.5) ** 2:
                        ok = False
                        break
                if not ok:
                    break
            if not ok:
                break
        if ok:
            grid.setdefault((gy, gx), []).append((cy, cx, r))
            placed.append((cy, cx, r))
    del gh, gw
    return placed


def make_cell_layer(seed: int, height: int, width: int, n_cells: int, r_lo: float = 6.0, r_hi: float = 14.0):
    """-> (float64 (H, W) unit-amplitude cell layer, int32 (H, W) label mask, K)."""
    rng = np.random.default_rng(seed)
    layer = np.zeros((height, width), dtype=np.float64)
    labels = np.zeros((height, width), dtype=np.int32)
    cells = _place_cells(rng, height, width, n_cells, r_lo, r_hi)
    k = 0
    for cy, cx, r in cells:
        ratio = rng.uniform(0.6, 1.0)
        theta = rng.uniform(0.0, np.pi)
        amp = rng.uniform(1500.0, 12000.0)
        a, b = r, r * ratio
        y0, y1 = max(int(cy - r) - 1, 0), min(int(cy + r) + 2, height)
        x0, x1 = max(int(cx - r) - 1, 0), min(int(cx + r) + 2, width)
        if y0 >= y1 or x0 >= x1:
            continue
        yy, xx = np.mgrid[y0:y1, x0:x1]
        dy, dx = yy - cy, xx - cx
        u = dx * np.cos(theta) + dy * np.sin(theta)
        v = -dx * np.sin(theta) + dy * np.cos(theta)
        q = (u / a) ** 2 + (v / b) ** 2
        inside = q < 1.0
        if not inside.any():
            continue
        k += 1
        layer[y0:y1, x0:x1] += amp * np.exp(-1.5 * q) * (q < 4.0)
        sub = labels[y0:y1, x0:x1]
        sub[inside & (sub == 0)] = k
    return layer, labels, k


def make_fov(seed: int, n_channels: int = 4, height: int = 2048, width: int = 2048, n_cells: int = 2000):
    """One seeded FOV: (uint16 (C, H, W), int32 (H, W) labels, K)."""
    layer, labels, k = make_cell_layer(seed, height, width, n_cells)
    rng = np.random.default_rng(seed + 1_000_003)
    fov = np.empty((n_channels, height, width), dtype=np.uint16)
    for c in range(n_channels):
        lam = BACKGROUND_LAMBDA[c % 4]
        gain = CHANNEL_GAIN[c % 4]
        img = rng.poisson(lam, size=(height, width)).astype(np.float64) + gain * layer
        fov[c] = np.clip(np.rint(img), 0, 65535).astype(np.uint16)
    return fov, labels, k
