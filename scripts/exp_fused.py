#!/usr/bin/env python
"""Debug aid: the fused pass 2 on a small image against the two-kernel route."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from arcadia_microscopy_tools_b200 import _gpu  # noqa: E402

h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 256)
sig = float(sys.argv[3]) if len(sys.argv) > 3 else 0.6
rng = np.random.default_rng(1)
imgs = rng.integers(0, 65536, size=(2, h, w)).astype(np.uint16)
tcg = _gpu.TensorCoreGaussian(16.0)
dev = _gpu.to_device(imgs)
digits = tcg.axis0(dev)
lo = _gpu.gauss_lo2d(dev, 1 / 65535.0, sig)
want, mm_w, bk_w = tcg.axis1(digits, lo, 1 / 65535.0, want_buckets=True)
torch.cuda.synchronize()
print("two-kernel route ok")
import os
from arcadia_microscopy_tools_b200 import _lib
_lib.check(_lib.load().amt_tune(b"tcg_debug", int(os.environ.get("DBG", "0"), 0)))
got, mm_g, bk_g = tcg.axis1_dog(digits, dev, sig, 1 / 65535.0, want_buckets=True)
torch.cuda.synchronize()
want, got = _gpu.to_host(want), _gpu.to_host(got)
bad = np.argwhere(want != got)
print("mismatches", len(bad), bad[:10], "max abs diff", np.abs(want - got).max())
