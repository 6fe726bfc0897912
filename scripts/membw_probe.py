#!/usr/bin/env python
"""Read-only, write-only and copy bandwidth of this GPU with plain torch ops (calibration of the roofline's denominator
for kernels whose traffic is not half reads, half writes)."""
import json

import torch

dev = torch.device("cuda", 0)
n = 1 << 30  # 1 Gi float32 = 4 GiB
a = torch.empty(n, dtype=torch.float32, device=dev).fill_(1.0)
b = torch.empty_like(a)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


res = {}
ms = timed(lambda: b.copy_(a))
res["copy_gbs_read_plus_write"] = 2 * a.numel() * 4 / ms / 1e6
ms = timed(lambda: b.fill_(2.0))
res["fill_gbs_write_only"] = a.numel() * 4 / ms / 1e6
ms = timed(lambda: b.zero_())
res["memset_gbs_write_only"] = a.numel() * 4 / ms / 1e6
ms = timed(lambda: a.sum())
res["sum_gbs_read_only"] = a.numel() * 4 / ms / 1e6
a16 = a.view(torch.int16)[: n]
ms = timed(lambda: torch.add(a[: n // 4], 1.0, out=b[: n // 4]))
res["add_scalar_gbs_read_plus_write"] = 2 * (n // 4) * 4 / ms / 1e6
print(json.dumps(res))
