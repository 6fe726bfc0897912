#!/usr/bin/env python
"""One executor batch of `n_fov` FOVs (default 16 = two 8-FOV chunks of config 2) in the default decision-exact mode,
run `reps` times: the program behind the launch list (`ncu --metrics gpu__time_duration.sum --kernel-name-base
demangled -k regex:amt::`) and the `ncu --set full` capture of the non-DoG kernels."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
from arcadia_microscopy_tools_b200.batch import FovBatchExecutor, FovPipelineConfig  # noqa: E402

n_fov = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
fovs, given, max_label = bench.build_device_batch(n_fov, 4, dev)
cfg = FovPipelineConfig(n_channels=4, height=2048, width=2048, seg_channel=1, chunk_fovs=8, max_labels=4096,
                        max_label_value=max_label)
with FovBatchExecutor(cfg, device=0) as ex:
    out = ex.alloc_outputs(n_fov)
    ms = [ex.run_device(fovs, given, out) for _ in range(reps)]
print({"ms_per_batch": [round(m, 3) for m in ms], "counts": int(out["counts_thr"].sum())})
