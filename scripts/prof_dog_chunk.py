#!/usr/bin/env python
"""The five DoG kernels of one executor chunk (32 planes of 2048 x 2048: 8 exact, 24 on the tensor cores), launched
exactly as bench.py's time_kernels launches them — one warm-up and one measured launch each — for
`ncu --set full` (profiles/r02_ncu_dog_kernels.json) and the launch list."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
from arcadia_microscopy_tools_b200 import _gpu, _lib  # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.load()
fovs, given, max_label = bench.build_device_batch(8, 2, dev)
hw = (_gpu.gaussian_half_weights(0.6), _gpu.gaussian_half_weights(16.0))
tcg = _gpu.TensorCoreGaussian(16.0)
k = bench.time_kernels(lib, _gpu, fovs, hw, steps=1, warmup=1, tcg=tcg, decision_exact=True)
print({key: round(v["ms"], 4) for key, v in k.items() if isinstance(v, dict)})
