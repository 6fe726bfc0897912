#!/usr/bin/env python
"""Experiment: lo2d / axis0 / axis1 alone (32 planes of 2048^2) under amt_tune("tcg_debug") masks given on the command line."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
from arcadia_microscopy_tools_b200 import _gpu, _lib  # noqa: E402
from scripts.bench_tcgauss import timed  # noqa: E402

C, H, W = 4, 2048, 2048
SCALE = 1.0 / 65535.0
dev = torch.device("cuda", 0)
lib = _lib.load()
fovs, given, max_label = bench.build_device_batch(8, 4, dev)
planes = 8 * C
x = fovs[:8].reshape(planes, H, W)
tcg = _gpu.TensorCoreGaussian(16.0)
hw_lo = _gpu.gaussian_half_weights(0.6)
d_lo = torch.from_numpy(hw_lo).to(dev)
lo = torch.empty((planes, H, W), dtype=torch.float64, device=dev)
digits = torch.empty((planes, 5, H, W), dtype=torch.uint8, device=dev)
out = torch.empty((planes, H, W), dtype=torch.float64, device=dev)
buckets = torch.empty((planes, H, W), dtype=torch.int16, device=dev)
mm = torch.empty((planes, 2), dtype=torch.int64, device=dev)
st = _gpu.stream_ptr()
p = _gpu.ptr
res = {}
ref = None
for mask in [int(a, 0) for a in sys.argv[1:]] or [0]:
    _lib.check(lib.amt_tune(b"tcg_debug", mask))
    r = {"lo2d_ms": timed(lambda: _lib.check(lib.amt_gauss_lo2d(p(x), SCALE, p(lo), planes, H, W, p(d_lo), len(hw_lo) - 1, 0, 0, st))),
         "axis0_ms": timed(lambda: _lib.check(lib.amt_tcg_axis0(tcg.handle, p(x), planes, H, W, p(digits), 0, 0, st))),
         "axis1_ms": timed(lambda: _lib.check(lib.amt_tcg_axis1(tcg.handle, p(digits), p(lo), SCALE, p(out), planes, H, W, p(buckets), p(mm), 0, 0, st)))}
    r["axis1_dog_ms"] = timed(lambda: _lib.check(lib.amt_tcg_axis1_dog(tcg.handle, p(digits), p(x), p(d_lo), len(hw_lo) - 1, SCALE, p(out), planes, H, W, p(buckets), p(mm), 0, 0, st)))
    chk = (float(lo.sum()), float(out.sum()), int(buckets.view(torch.uint8).sum()))
    if ref is None:
        ref = chk
    r["same_as_first"] = chk == ref
    res[hex(mask)] = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()}
_lib.check(lib.amt_tune(b"tcg_debug", 0))
print(json.dumps(res))
