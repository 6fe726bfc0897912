#!/usr/bin/env python
"""Per-stage device time of the executor (one chunk-sized batch repeated), retry count, in the given modes."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from arcadia_microscopy_tools_b200.batch import FovBatchExecutor, FovPipelineConfig  # noqa: E402

n_fov = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
fovs, given, max_label = bench.build_device_batch(n_fov, 4, dev)
res = {}
for name, kw in {"decision_exact": {}, "float64_seg": dict(seg_plane_filter="float64"), "fma": dict(plane_filter="fma")}.items():
    cfg = FovPipelineConfig(n_channels=4, height=2048, width=2048, seg_channel=1, chunk_fovs=8, max_labels=4096,
                            max_label_value=max_label, **kw)
    with FovBatchExecutor(cfg, device=0) as ex:
        out = ex.alloc_outputs(n_fov)
        for _ in range(2):
            ex.run_device(fovs, given, out)
        ms = [ex.run_device(fovs, given, out) for _ in range(5)]
        ex.set_profiling(True)
        ex.run_device(fovs, given, out)
        st, chunks = ex.stage_ms()
        ex.set_profiling(False)
        res[name] = {"ms_per_chunk": float(np.median(ms)) / (n_fov / 8), "retries": ex.retry_count,
                     "stage_ms_per_chunk": {k: round(v / chunks, 4) for k, v in st.items()},
                     "counts": int(out["counts_thr"].sum())}
print(json.dumps(res))
