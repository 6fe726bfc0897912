"""Markdown table of one executor chunk from the condensed ncu capture (profiles/ncu_to_json.py output):
per kernel the launches, total duration, share, DRAM GB/s against the measured HBM peak, issue-slot and
FP64-pipe occupancy.

    python profiles/kernel_table.py profiles/r01_ncu_full_one_chunk.json > profiles/r01_kernel_table.md
"""
import json
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def main() -> None:
    d = json.loads(Path(sys.argv[1]).read_text())
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
    rows: "OrderedDict[str, dict]" = OrderedDict()
    for k in d["kernels"]:
        name = k["kernel"].split("(")[0].replace("void ", "").replace("amt::", "").replace("(int)", "").replace("(bool)", "")
        r = rows.setdefault(name, {"n": 0, "t": 0.0, "bytes": 0.0, "issue": 0.0, "fp64": 0.0})
        r["n"] += 1
        r["t"] += k["duration"]
        r["bytes"] += k.get("dram_bytes", 0.0)
        r["issue"] += k.get("issue_active_pct", 0.0) * k["duration"]
        r["fp64"] += k.get("fp64_pipe_active_pct", 0.0) * k["duration"]
    total = sum(r["t"] for r in rows.values())
    print("# One 8-FOV chunk (32 planes / 8 label images of 2048x2048), `ncu --set full`, B200\n")
    print(f"Source: `{d.get('source', '?')}` condensed into `{Path(sys.argv[1]).name}`.  HBM peak (MEASURED_PEAKS.json, copy): "
          f"{peak} GB/s.  Times are cold-cache and serialised; shares agree with the live launch list "
          "(`r01_launches_final.txt`).\n")
    print("| kernel | launches | total µs | share | DRAM GB/s | % of HBM peak | issue slots busy | FP64 pipe busy |")
    print("|---|---|---|---|---|---|---|---|")
    for name, r in rows.items():
        gbs = r["bytes"] / r["t"] / 1e9
        print(f"| `{name}` | {r['n']} | {r['t'] * 1e6:.1f} | {100 * r['t'] / total:.1f}% | {gbs:.0f} | {100 * gbs / peak:.1f}% | "
              f"{r['issue'] / r['t']:.0f}% | {r['fp64'] / r['t']:.0f}% |")
    print(f"| **total** | {sum(r['n'] for r in rows.values())} | {total * 1e6:.1f} | | | | | |")


if __name__ == "__main__":
    main()
