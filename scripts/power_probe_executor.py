#!/usr/bin/env python
"""Does the whole executor run at the power cap?  64 FOVs of config 2 through run_device in a loop for ~4 s while a thread
samples NVML power / SM clock / throttle reasons every 5 ms; then the same with the DoG only / the remainder only is
approximated by the per-kernel probes (scripts/power_probe.py).  One JSON line."""
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
from arcadia_microscopy_tools_b200 import _lib  # noqa: E402
from arcadia_microscopy_tools_b200.batch import FovBatchExecutor, FovPipelineConfig  # noqa: E402

dev = torch.device("cuda", 0)
n_fov = 64
fovs, given, max_label = bench.build_device_batch(n_fov, 4, dev)
pynvml.nvmlInit()
hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
res = {"power_limit_w": pynvml.nvmlDeviceGetEnforcedPowerLimit(hnd) / 1000.0}
lib = _lib.load()


def sample_while(fn, seconds):
    samples, stop = [], threading.Event()

    def sampler():
        while not stop.is_set():
            samples.append((time.perf_counter(), pynvml.nvmlDeviceGetPowerUsage(hnd) / 1000.0,
                            pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM),
                            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(hnd)))
            time.sleep(0.005)

    th = threading.Thread(target=sampler)
    th.start()
    t0 = time.perf_counter()
    ms = []
    while time.perf_counter() - t0 < seconds:
        ms.append(fn())
    stop.set()
    th.join()
    s = [x for x in samples if x[0] - t0 > 1.0]  # NVML's power reading lags: drop the first second
    pw, ck = np.array([x[1] for x in s]), np.array([x[2] for x in s])
    capped = sum(1 for x in s if x[3] & 0x4)
    return {"ms_first": round(ms[0], 3), "ms_median_after_1s": round(float(np.median(ms[len(ms) // 3:])), 3), "batches": len(ms),
            "power_w_median": float(np.median(pw)), "power_w_max": float(pw.max()), "sm_mhz_median": float(np.median(ck)),
            "sm_mhz_min": float(ck.min()), "share_of_samples_power_capped": round(capped / max(len(s), 1), 3)}


for mode, tune in (("default", {}), ("float64_everywhere", {"exec_tc": 0})):
    for k, v in tune.items():
        _lib.check(lib.amt_tune(k.encode(), v))
    cfg = FovPipelineConfig(n_channels=4, height=2048, width=2048, seg_channel=1, chunk_fovs=32, max_labels=4096,
                            max_label_value=max_label)
    with FovBatchExecutor(cfg, device=0) as ex:
        out = ex.alloc_outputs(n_fov)
        for _ in range(2):
            ex.run_device(fovs, given, out, sync=True)
        res[mode] = sample_while(lambda: ex.run_device(fovs, given, out, sync=True), 4.0)
    for k in tune:
        _lib.check(lib.amt_tune(k.encode(), 1))
print(json.dumps(res))
