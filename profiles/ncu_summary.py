"""Print the key per-kernel metrics of an .ncu-rep (raw page): duration, DRAM bytes, FP64 pipe,
issue utilisation, instruction count, occupancy limits and the top warp-stall reasons.

    python profiles/ncu_summary.py gpurun_out/x.ncu-rep [kernel-regex]
"""
import csv
import io
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "launch__block_size", "sm__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def main() -> None:
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        if pat and not pat.search(r[kn]):
            continue
        print("-----", r[kn][:150])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w}: {r[i]} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active\.ratio", h)
            if m:
                try:
                    stalls.append((float(r[i]), m.group(1)))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  stalls (warps per issue-active cycle):", ", ".join(f"{n}={v:.2f}" for v, n in stalls[:7]))


if __name__ == "__main__":
    main()
