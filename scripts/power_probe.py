#!/usr/bin/env python
"""Is the fused pass 2 power-limited?  Runs `amt_tcg_axis1_dog` (32 planes of 2048 x 2048) in a loop for a few seconds
per tcg_debug mask while a thread samples NVML power, SM clock and throttle reasons every 5 ms.  One JSON line."""
import json
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import pynvml  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from arcadia_microscopy_tools_b200 import _gpu, _lib  # noqa: E402

C, H, W = 4, 2048, 2048
SCALE = 1.0 / 65535.0
dev = torch.device("cuda", 0)
lib = _lib.load()
fovs, given, max_label = bench.build_device_batch(8, 4, dev)
planes = 8 * C
x = fovs.reshape(planes, H, W)
tcg = _gpu.TensorCoreGaussian(16.0)
hw_lo = _gpu.gaussian_half_weights(0.6)
d_lo = torch.from_numpy(hw_lo).to(dev)
digits = torch.empty((planes, 5, H, W), dtype=torch.uint8, device=dev)
out = torch.empty((planes, H, W), dtype=torch.float64, device=dev)
buckets = torch.empty((planes, H, W), dtype=torch.int16, device=dev)
mm = torch.empty((planes, 2), dtype=torch.int64, device=dev)
st = _gpu.stream_ptr()
p = _gpu.ptr
pynvml.nvmlInit()
hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
limit_w = pynvml.nvmlDeviceGetEnforcedPowerLimit(hnd) / 1000.0


def a0():
    _lib.check(lib.amt_tcg_axis0(tcg.handle, p(x), planes, H, W, p(digits), 0, 0, st))


def a1():
    _lib.check(lib.amt_tcg_axis1_dog(tcg.handle, p(digits), p(x), p(d_lo), len(hw_lo) - 1, SCALE, p(out), planes, H, W,
                                     p(buckets), p(mm), 0, 0, st))


def copy():
    out.copy_(out2)


out2 = torch.ones_like(out)


def sample_while(fn, seconds=2.5):
    samples = []
    stop = threading.Event()

    def sampler():
        while not stop.is_set():
            samples.append((pynvml.nvmlDeviceGetPowerUsage(hnd) / 1000.0,
                            pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM),
                            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(hnd)))
            time.sleep(0.005)

    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    th = threading.Thread(target=sampler)
    th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    n = 0
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    stop.set()
    th.join()
    s = samples[len(samples) // 5:]
    pw = np.array([a for a, _, _ in s])
    ck = np.array([b for _, b, _ in s])
    reasons = 0
    for _, _, r in s:
        reasons |= r
    return {"ms": round(e0.elapsed_time(e1) / n, 4), "power_w_median": float(np.median(pw)), "power_w_max": float(pw.max()),
            "sm_mhz_median": float(np.median(ck)), "sm_mhz_min": float(ck.min()), "throttle_reasons_or": hex(reasons), "samples": len(s)}


res = {"power_limit_w": limit_w}
a0()
res["axis0"] = sample_while(a0)
for mask in (0, 1, 32, 2, 35):
    _lib.check(lib.amt_tune(b"tcg_debug", mask))
    res[f"fused_dbg{mask}"] = sample_while(a1)
_lib.check(lib.amt_tune(b"tcg_debug", 0))
res["copy_f64_planes"] = sample_while(copy)
print(json.dumps(res))
