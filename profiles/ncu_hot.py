"""Hottest SASS instructions (by warp stall samples) and opcode mix of one kernel of an .ncu-rep.

    python profiles/ncu_hot.py rep.ncu-rep <kernel index> [n]
"""
import csv
import io
import subprocess
import sys
from collections import Counter


def blocks(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    out, cur = [], None
    for r in csv.reader(io.StringIO(txt)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            out.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    return out


def main():
    rep, idx = sys.argv[1], int(sys.argv[2])
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    b = blocks(rep)[idx]
    hdr, data = b["rows"][0], [r for r in b["rows"][1:] if len(r) > 10]
    ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    istall = hdr.index("Warp Stall Sampling (All Samples)")
    print(b["name"][:100])
    tot = sum(int(r[ie]) for r in data)
    print("warp instructions", tot, "samples", sum(int(r[isamp]) for r in data))
    oc = Counter()
    for r in data:
        parts = r[ia].split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        oc[op.split(".")[0]] += int(r[ie])
    print(oc.most_common(22))
    for k, r in enumerate(data):
        r.append(k)
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:n]:
        print(f"{r[isamp]:>6} {r[ie]:>9} #{r[-1]:<5} {r[ia][:90]}")


if __name__ == "__main__":
    main()
