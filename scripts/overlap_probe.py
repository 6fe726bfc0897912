"""Does an HBM-bound kernel overlap with the FP64-bound DoG when they run on two streams?
Times A (DoG passes) alone, B (streaming subtraction kernels) alone and A || B.  Tooling only.

    AMT_TUNE=dog_variant=1,dog_ctas=1 python scripts/overlap_probe.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from arcadia_microscopy_tools_b200 import _gpu, _lib as L  # noqa: E402

lib = L.load()
planes, H, W = 32, 2048, 2048
dev = torch.device("cuda", 0)
img = torch.randint(0, 32767, (planes, H, W), dtype=torch.int16, device=dev)
hw_lo, hw_hi = _gpu.gaussian_half_weights(0.6), _gpu.gaussian_half_weights(16.0)
d_lo, d_hi = torch.from_numpy(hw_lo).to(dev), torch.from_numpy(hw_hi).to(dev)
tmp_lo = torch.empty((planes, H, W), dtype=torch.float64, device=dev)
tmp_hi = torch.empty_like(tmp_lo)
out = torch.empty_like(tmp_lo)
mm = torch.empty((planes, 2), dtype=torch.int64, device=dev)
a = torch.rand((planes, H, W), dtype=torch.float64, device=dev)
b = torch.rand_like(a)
c = torch.empty_like(a)
lo_p, hi_p = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
sA = torch.cuda.Stream(priority=0)
import os

sB = torch.cuda.Stream(priority=int(os.environ.get("PROBE_B_PRIORITY", "-1")))
NA, NB = 6, 24


def run_a():
    with torch.cuda.stream(sA):
        for _ in range(NA):
            L.check(lib.amt_dog2d(_gpu.ptr(img), L.AMT_U16, 1.0 / 65535.0, _gpu.ptr(out), planes, H, W, _gpu.ptr(d_lo),
                                  len(hw_lo) - 1, _gpu.ptr(d_hi), len(hw_hi) - 1, _gpu.ptr(tmp_lo), _gpu.ptr(tmp_hi),
                                  _gpu.ptr(mm), sA.cuda_stream))


def run_b():
    with torch.cuda.stream(sB):
        for _ in range(NB):
            L.check(lib.amt_sub_f64(_gpu.ptr(a), _gpu.ptr(b), _gpu.ptr(c), a.numel(), sB.cuda_stream))


def timed(fns):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sA.wait_event(e0)
    sB.wait_event(e0)
    for f in fns:
        f()
    ea, eb = torch.cuda.Event(), torch.cuda.Event()
    ea.record(sA)
    eb.record(sB)
    torch.cuda.current_stream().wait_event(ea)
    torch.cuda.current_stream().wait_event(eb)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for f in (run_a, run_b):
    f()
torch.cuda.synchronize()
ta = min(timed([run_a]) for _ in range(3))
tb = min(timed([run_b]) for _ in range(3))
tab = min(timed([run_a, run_b]) for _ in range(3))
tba = min(timed([run_b, run_a]) for _ in range(3))
gb = NB * 3 * a.numel() * 8 / 1e9
print(f"A (DoG x{NA}) {ta:.2f} ms | B (sub x{NB}, {gb:.1f} GB) {tb:.2f} ms = {gb / tb * 1e3:.0f} GB/s | "
      f"A||B {tab:.2f} ms, B||A {tba:.2f} ms | serial {ta + tb:.2f} ms, ideal {max(ta, tb):.2f} ms")
