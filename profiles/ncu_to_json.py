"""Condense an .ncu-rep (--set full) into a small JSON: per kernel launch the duration, DRAM bytes,
pipe / issue utilisation, instruction count, occupancy and the top stall reasons.  bench.py reads
`dram_bytes` of the dominant kernel from the committed JSON for `roofline.traffic`.

    python profiles/ncu_to_json.py gpurun_out/x.ncu-rep profiles/r01_x.json
"""
import csv
import io
import json
import re
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def main() -> None:
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    kernels = []
    for r in rows[2:]:
        k = {"kernel": re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("amt::", "")}
        for col, name in KEYS.items():
            if col in hdr:
                i = hdr.index(col)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                k[name] = v * UNIT_SCALE.get(units[i], 1.0) if name in ("duration", "dram_read", "dram_write") else v
        if "dram_read" in k and "dram_write" in k:
            k["dram_bytes"] = k["dram_read"] + k["dram_write"]
            if k.get("duration"):
                k["dram_gbs"] = k["dram_bytes"] / k["duration"] / 1e9
        stalls = []
        for i, h in enumerate(hdr):
            m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active\.ratio", h)
            if m:
                try:
                    stalls.append((float(r[i]), m.group(1)))
                except ValueError:
                    pass
        k["top_stalls"] = {n: round(v, 2) for v, n in sorted(stalls, reverse=True)[:5]}
        kernels.append(k)
    json.dump({"source": rep.split("/")[-1], "how": "ncu --set full --clock-control none (cold-cache, serialised launches)",
               "kernels": kernels}, open(out, "w"), indent=1)
    print(f"{len(kernels)} launches -> {out}")


if __name__ == "__main__":
    main()
