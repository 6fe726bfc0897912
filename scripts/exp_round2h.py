#!/usr/bin/env python
"""A/B of a kernel variant on one chunk's worth of data (8 label images of 2048 x 2048), checked bit for bit against the
variant it replaces, then the whole executor both ways.
  dx_collect_threads 256 | 1024 : CTA size of the decision-exact collect pass
(Earlier uses of this script: tcg_p1_warps, region_wide, ccl_touch_filter, region_prefetch; see
profiles/r02_tcgauss_experiments.md, steps 17-22.)
One JSON line."""

from __future__ import annotations

import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import bench  # noqa: E402
from arcadia_microscopy_tools_b200 import _gpu, _lib  # noqa: E402
from arcadia_microscopy_tools_b200.batch import FovBatchExecutor, FovPipelineConfig  # noqa: E402

C, H, W = 4, 2048, 2048


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
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    tune = lambda k, v: _lib.check(lib.amt_tune(k, v), "amt_tune")  # noqa: E731
    n_fov = 64
    fovs, given, max_label = bench.build_device_batch(n_fov, 8, dev)
    res = {}
    # (the sections of the experiments that were measured and reverted are gone with their switches; the log has the numbers)
    # ---- the executor (64 FOVs, chunks of 32) with the decision-exact collect pass at 256 / 1024 threads per CTA
    cfg = FovPipelineConfig(n_channels=C, height=H, width=W, seg_channel=bench.SEG_CHANNEL, chunk_fovs=32, max_labels=4096,
                            max_label_value=max_label)
    for thr in (256, 1024, 256, 1024):
        tune(b"dx_collect_threads", thr)
        with FovBatchExecutor(cfg, device=0) as ex:
            out = ex.alloc_outputs(n_fov)
            for _ in range(2):
                ex.run_device(fovs, given, out, sync=True)
            ms = [ex.run_device(fovs, given, out, sync=True) for _ in range(5)]
        res.setdefault(f"executor_ms_per_8_fov_collect{thr}", []).append(sum(ms) / len(ms) / (n_fov / 8))
    tune(b"dx_collect_threads", 1024)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
