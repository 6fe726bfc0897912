#!/usr/bin/env python
"""Per-kernel timing of the tensor-core DoG path next to the float64 strip kernels (CUDA events, 32 planes of
2048 x 2048 = one executor chunk), then the whole executor with the path on and off.  One JSON line."""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import bench  # noqa: E402
from arcadia_microscopy_tools_b200 import _gpu, _lib  # noqa: E402
from arcadia_microscopy_tools_b200.batch import FovBatchExecutor, FovPipelineConfig  # noqa: E402

C, H, W = 4, 2048, 2048
SCALE = 1.0 / 65535.0


def timed(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    n_fov = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 32
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    fovs, given, max_label = bench.build_device_batch(n_fov, 4, dev)
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
    for name, (every, off) in {"24_of_32_planes": (4, 1), "32_planes": (0, 0)}.items():
        n_sel = planes - planes // every if every else planes
        px = n_sel * H * W
        t_lo = timed(lambda: _lib.check(lib.amt_gauss_lo2d(p(x), SCALE, p(lo), planes, H, W, p(d_lo), len(hw_lo) - 1, every, off, st)))
        t_a0 = timed(lambda: _lib.check(lib.amt_tcg_axis0(tcg.handle, p(x), planes, H, W, p(digits), every, off, st)))
        t_a1 = timed(lambda: _lib.check(lib.amt_tcg_axis1(tcg.handle, p(digits), p(lo), SCALE, p(out), planes, H, W, p(buckets), p(mm), every, off, st)))
        res[name] = {
            "lo2d_ms": t_lo, "axis0_ms": t_a0, "axis1_ms": t_a1,
            "lo2d_gbs": px * 10 / t_lo / 1e6, "axis0_gbs": px * 7 / t_a0 / 1e6, "axis1_gbs": px * 23 / t_a1 / 1e6,
            "axis0_tmacs": px * 4 * 2 * 256 / t_a0 / 1e9, "axis1_tmacs": px * 17 * 256 / t_a1 / 1e9,
            "us_per_plane": 1e3 * (t_lo + t_a0 + t_a1) / n_sel,
        }
    # where the time goes: the same launches with parts switched off (amt_tune "tcg_debug": 1 = no MMAs,
    # 2 = no epilogue arithmetic / stores, 4 = no TMEM loads)
    dec = {}
    for mask in (0, 64, 1, 2, 3):
        _lib.check(lib.amt_tune(b"tcg_debug", mask))
        dec[f"dbg{mask}"] = {
            "axis0_ms": timed(lambda: _lib.check(lib.amt_tcg_axis0(tcg.handle, p(x), planes, H, W, p(digits), 0, 0, st)), 5, 2),
            "axis1_ms": timed(lambda: _lib.check(lib.amt_tcg_axis1(tcg.handle, p(digits), p(lo), SCALE, p(out), planes, H, W, p(buckets), p(mm), 0, 0, st)), 5, 2)}
    _lib.check(lib.amt_tune(b"tcg_debug", 0))
    res["decompose_32_planes"] = dec
    if "--kernels-only" in sys.argv:
        print(json.dumps(res))
        return
    # whole executor, device-resident
    cfg = FovPipelineConfig(n_channels=C, height=H, width=W, seg_channel=1, chunk_fovs=8, max_labels=4096,
                            max_label_value=max_label)
    with FovBatchExecutor(cfg, device=0) as ex:
        o = ex.alloc_outputs(n_fov)
        for tc_on in (1, 0, 1):
            _lib.check(lib.amt_tune(b"exec_tc", tc_on))
            for _ in range(2):
                ex.run_device(fovs, given, o, sync=True)
            ms = [ex.run_device(fovs, given, o, sync=True) for _ in range(5)]
            key = f"executor_tc{tc_on}"
            res.setdefault(key, []).append({"ms_per_chunk": float(np.median(ms)) / (n_fov / 8),
                                            "gpix_s": n_fov * C * H * W / (float(np.median(ms)) * 1e-3) / 1e9,
                                            "counts_thr_sum": int(o["counts_thr"].sum())})
        res["uses_tensor_cores"] = ex.uses_tensor_cores
    print(json.dumps(res))


if __name__ == "__main__":
    main()
