"""Summarise an `ncu --metrics gpu__time_duration.sum --kernel-name-base demangled -k regex:amt:: --csv` launch list: one line
per kernel (template arguments kept) with launch count, total and mean duration, share of the listed launches, grid and
block size.  `reps` = how many times the profiled program repeated its batch: only the LAST repetition is tabulated.

    python profiles/launch_table.py gpurun_out/launches.csv [reps]
"""
import csv
import re
import sys
from collections import OrderedDict


def main() -> None:
    path = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    data = []
    for r in rows[hi + 1:]:
        if len(r) <= vi or re.search(r"\bat::|cub::|native::", r[ki]):
            continue  # torch kernels of the synthetic input generator, if the list was not filtered with -k regex:amt::
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
        m = re.match(r"(?:void )?(?:amt::)?((?:\w+::)*\w+(?:<[^(]*>)?)\(", r[ki])
        name = m.group(1) if m else r[ki][:60]
        name = name.replace("(int)", "").replace("(bool)", "")
        data.append((name[:64], float(r[vi].replace(",", "")) * scale, r[gi], r[bi]))
    data = data[len(data) - len(data) // reps:]
    agg: "OrderedDict[str, list]" = OrderedDict()
    for name, us, g, b in data:
        a = agg.setdefault(name, [0, 0.0, g, b])
        a[0] += 1
        a[1] += us
    total = sum(v[1] for v in agg.values())
    print(f"{'kernel':64s} {'launches':>8s} {'total_us':>10s} {'mean_us':>9s} {'share':>6s}  grid  block")
    for name, (n, us, g, b) in agg.items():
        print(f"{name:64s} {n:8d} {us:10.1f} {us / n:9.1f} {100 * us / total:5.1f}%  {g} {b}")
    print(f"{'TOTAL':64s} {sum(v[0] for v in agg.values()):8d} {total:10.1f}")


if __name__ == "__main__":
    main()
