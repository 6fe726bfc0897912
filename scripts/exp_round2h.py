#!/usr/bin/env python
"""A/B of a kernel variant on one chunk's worth of data (8 label images of 2048 x 2048), checked bit for bit against the
variant it replaces, then the whole executor both ways.
  ccl_touch_filter 0 | 1 : integer masks with border clearing labelled everywhere | only where a border value occurs
(Earlier uses of this script: tcg_p1_warps, region_wide; see profiles/r02_tcgauss_experiments.md, steps 17-19.)
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
    # ---- labelling of the given integer masks (clear_border + relabel_sequential), one chunk of 8
    labels_in = given[:8].contiguous()
    outs = {}
    for filt in (0, 1):
        tune(b"ccl_touch_filter", filt)
        f = lambda: _gpu.label(labels_in, 2, True, max_value=int(max_label))  # noqa: E731
        res[f"label_given_ms_filter{filt}"] = timed(f)
        outs[filt] = f()
    res["label_given_bit_identical"] = bool(torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]))
    res["label_given_counts"] = outs[1][1].tolist()
    del outs
    # ---- the executor both ways (64 FOVs, chunks of 32)
    ref = None
    cfg = FovPipelineConfig(n_channels=C, height=H, width=W, seg_channel=bench.SEG_CHANNEL, chunk_fovs=32, max_labels=4096,
                            max_label_value=max_label)
    for filt in (0, 1, 0, 1):
        tune(b"ccl_touch_filter", filt)
        with FovBatchExecutor(cfg, device=0) as ex:
            out = ex.alloc_outputs(n_fov, labels=True)
            for _ in range(2):
                ex.run_device(fovs, given, out, sync=True)
            ms = [ex.run_device(fovs, given, out, sync=True) for _ in range(5)]
            snap = {k: out[k].clone() for k in ("tables_thr", "tables_given", "counts_thr", "counts_given", "thresholds", "labels_given")}
        if ref is None:
            ref = snap
        def rows_equal(k, cnt):  # tables are max_labels wide: only the first count columns of an image are written
            a, b = ref[k], snap[k]
            return all(torch.equal(a[i][:, :int(c)].view(torch.int64), b[i][:, :int(c)].view(torch.int64)) for i, c in enumerate(ref[cnt].tolist()))

        same = (all(torch.equal(ref[k], snap[k]) for k in ("counts_thr", "counts_given", "labels_given"))
                and torch.equal(ref["thresholds"].view(torch.int64), snap["thresholds"].view(torch.int64))
                and rows_equal("tables_thr", "counts_thr") and rows_equal("tables_given", "counts_given"))
        res.setdefault(f"executor_ms_per_8_fov_filter{filt}", []).append(sum(ms) / len(ms) / (n_fov / 8))
        res[f"executor_same_filter{filt}"] = bool(same)
    tune(b"ccl_touch_filter", 1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
