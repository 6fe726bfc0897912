#!/usr/bin/env python
"""The fused pass 2 (amt_tcg_axis1_dog) alone on 32 planes of 2048^2, twice: the program behind its ncu capture."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
from arcadia_microscopy_tools_b200 import _gpu, _lib  # noqa: E402

C, H, W = 4, 2048, 2048
dev = torch.device("cuda", 0)
lib = _lib.load()
fovs, given, max_label = bench.build_device_batch(8, 4, dev)
x = fovs.reshape(32, H, W)
tcg = _gpu.TensorCoreGaussian(16.0)
digits = tcg.axis0(x)
for _ in range(2):
    out, mm, bk = tcg.axis1_dog(digits, x, 0.6, 1 / 65535.0, want_buckets=True)
torch.cuda.synchronize()
print("ok")
