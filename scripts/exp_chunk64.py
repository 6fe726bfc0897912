#!/usr/bin/env python
"""Executor chunk size 32 against 64 fields of view (128 FOVs of config 2, device-resident).  One JSON line."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from arcadia_microscopy_tools_b200.batch import FovBatchExecutor, FovPipelineConfig  # noqa: E402

dev = torch.device("cuda", 0)
n_fov = 128
fovs, given, max_label = bench.build_device_batch(n_fov, 8, dev)
res = {}
for rep in range(2):
    for chunk in (32, 64):
        cfg = FovPipelineConfig(n_channels=4, height=2048, width=2048, seg_channel=1, chunk_fovs=chunk, max_labels=4096,
                                max_label_value=max_label)
        with FovBatchExecutor(cfg, device=0) as ex:
            out = ex.alloc_outputs(n_fov)
            for _ in range(2):
                ex.run_device(fovs, given, out, sync=True)
            ms = [ex.run_device(fovs, given, out, sync=True) for _ in range(6)]
            res.setdefault(f"chunk_{chunk}", []).append({"ms_per_8_fov": round(float(np.median(ms)) / (n_fov / 8), 4),
                                                         "counts": int(out["counts_thr"].sum()),
                                                         "mem_gb": round(torch.cuda.mem_get_info(0)[0] / 1e9, 1)})
        torch.cuda.empty_cache()
print(json.dumps(res))
