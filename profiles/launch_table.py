"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one line per kernel
name with launch count, total and mean duration, and share of the listed launches.

    python profiles/launch_table.py gpurun_out/launches.csv [first] [count]
"""
import csv
import re
import sys
from collections import OrderedDict


def main() -> None:
    path = sys.argv[1]
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    count = int(sys.argv[3]) if len(sys.argv) > 3 else None
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    data = []
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
        full = r[ki]
        if "at::" in full or "cub::" in full or full.startswith(("void at", "void (anonymous")):
            continue  # torch kernels of the synthetic input generator, not the library
        name = re.match(r"(?:void )?(?:amt::)?([A-Za-z0-9_]+)", full).group(1)
        if name in ("native", "cuda", "vectorized_elementwise_kernel", "distribution_elementwise_grid_stride_kernel"):
            continue
        data.append((name, float(r[vi].replace(",", "")) * scale))
    data = data[first:first + count] if count else data[first:]
    agg: "OrderedDict[str, list]" = OrderedDict()
    for name, us in data:
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += us
    total = sum(v[1] for v in agg.values())
    print(f"{'kernel':28s} {'launches':>8s} {'total_us':>10s} {'mean_us':>9s} {'share':>6s}")
    for name, (n, us) in agg.items():
        print(f"{name:28s} {n:8d} {us:10.1f} {us / n:9.1f} {100 * us / total:5.1f}%")
    print(f"{'TOTAL':28s} {len(data):8d} {total:10.1f}")


if __name__ == "__main__":
    main()
