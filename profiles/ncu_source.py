"""Per-instruction view of one kernel of an .ncu-rep (source page): executed counts, stall
samples and the top stall reasons, restricted to the hottest loop unless --all is given.

    python profiles/ncu_source.py rep.ncu-rep <kernel index> [--all]
"""
import csv
import io
import subprocess
import sys


def blocks(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    out, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            out.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    return out


def main():
    rep, idx = sys.argv[1], int(sys.argv[2])
    show_all = "--all" in sys.argv
    b = blocks(rep)[idx]
    hdr, data = b["rows"][0], [r for r in b["rows"][1:] if len(r) > 10]
    ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    print(b["name"][:120])
    tot = sum(int(r[isamp]) for r in data)
    mx = max(int(r[ie]) for r in data)
    sel = data if show_all else [r for r in data if int(r[ie]) == mx]
    print(f"total samples {tot}; shown instrs {len(sel)} samples {sum(int(r[isamp]) for r in sel)} (exec count {mx})")
    agg = {}
    for r in sel:
        for i, h in stall_cols:
            agg[h] = agg.get(h, 0) + int(r[i] or 0)
    print(sorted(agg.items(), key=lambda x: -x[1])[:8])
    for r in sel:
        st = sorted(((h[6:], int(r[i] or 0)) for i, h in stall_cols if int(r[i] or 0) > 0), key=lambda x: -x[1])
        print(f"{int(r[ie]):9d} {int(r[isamp]):5d} {r[ia].strip()[:64]:64s} {st[:3]}")


if __name__ == "__main__":
    main()
