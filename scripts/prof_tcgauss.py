#!/usr/bin/env python
"""One launch of each kernel of the tensor-core DoG path on 8 planes of 2048 x 2048, for ncu."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from arcadia_microscopy_tools_b200 import _gpu, _lib  # noqa: E402

H = W = 2048
planes = 8
rng = np.random.default_rng(0)
img = (rng.poisson(300, size=(planes, H, W)) + (rng.random((planes, H, W)) < 0.01) * 5000).astype(np.uint16)
x = _gpu.to_device(img)
tcg = _gpu.TensorCoreGaussian(16.0)
for _ in range(2):
    lo = _gpu.gauss_lo2d(x, 1 / 65535.0, 0.6)
    dig = tcg.axis0(x)
    out, mm, b = tcg.axis1(dig, lo, 1 / 65535.0, want_buckets=True)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
