"""Oracle: ``np.percentile`` (method='linear').  TEST INFRASTRUCTURE ONLY.

Called by the reference at ``operations.py:47`` and ``operations.py:94``.  ``percentile`` is
the real numpy routine; ``percentile_restated`` spells out the rank / lerp arithmetic the
CUDA select kernels reproduce and is pinned against numpy in
tests/test_oracle.py.
"""

from __future__ import annotations

import numpy as np


def percentile(a: np.ndarray, q):
    return np.percentile(a, q)


def rank_pair(n: int, q: float) -> tuple[int, int, float]:
    """(lo, hi, gamma) for percentile q over n elements: numpy's virtual index
    ``v = (n-1) * (q/100)``, ``lo = floor(v)``, ``hi = min(lo+1, n-1)``, ``gamma = v - lo``."""
    quant = np.true_divide(np.float64(q), 100.0)
    v = (n - 1) * quant
    lo = int(np.floor(v))
    lo = min(max(lo, 0), n - 1)
    hi = min(lo + 1, n - 1)
    gamma = float(v - lo)
    return lo, hi, gamma


def lerp(a: float, b: float, t: float) -> float:
    """numpy ``_lerp``: ``a + (b-a)*t`` if t < 0.5 else ``b - (b-a)*(1-t)``; exact b when
    t == 1 is implied by the second form."""
    a = np.float64(a)
    b = np.float64(b)
    t = np.float64(t)
    diff = b - a
    if t >= 0.5:
        out = b - diff * (1 - t)
    else:
        out = a + diff * t
    return float(out)


def percentile_restated(a: np.ndarray, q):
    """Sort-based restatement.  uint16 inputs: order statistics are promoted to float64
    before the lerp (numpy subtracts in the array dtype for ``b - a`` only when both are
    arrays of that dtype; here ``take`` results are lerped as float64 after
    ``np.percentile`` converts via ``asanyarray(..., float)`` weights) — verified equal to
    ``np.percentile`` on uint16 in tests."""
    flat = np.sort(np.asarray(a).reshape(-1))
    n = flat.size
    scalar = np.isscalar(q)
    qs = [q] if scalar else list(q)
    out = []
    for qq in qs:
        lo, hi, g = rank_pair(n, qq)
        out.append(lerp(float(flat[lo]), float(flat[hi]), g))
    return out[0] if scalar else np.array(out)
