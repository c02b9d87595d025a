"""Reads an `ncu --page source --csv` dump and prints, per kernel, the instructions that collect the most warp
samples (with their dominant stall reasons) -- enough to tell which warp role of a warp-specialised kernel waits where.

    python tools/ncu_roles.py gpurun_out/x_source.csv [kernel_index] [min_pct]
"""
import csv
import sys

csv.field_size_limit(10 ** 9)


def sections(path):
    secs, cur = [], None
    for r in csv.reader(open(path)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            secs.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    # ncu prints every kernel twice (two views); keep the first of each pair
    return secs[::2] if len(secs) > 1 and secs[0]["name"] == secs[1]["name"] and len(secs) % 2 == 0 else secs


def main():
    secs = sections(sys.argv[1])
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
    s = secs[which]
    h = s["rows"][0]
    idx = {n: i for i, n in enumerate(h)}
    data = s["rows"][1:]
    S = [int(r[idx["# Samples"]] or 0) for r in data]
    tot = sum(S)
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    print(s["name"][:90], "| kernels in file:", len(secs), "| samples", tot)
    marks = ("UTMALDG", "UTCHMMA", "MUFU.TANH", "LDTM", "UTMASTG", "UTCBAR", "BAR.SYNC")
    seen = set()
    for i, r in enumerate(data):
        src = r[idx["Source"]].strip()
        m = next((k for k in marks if k in src), None)
        first_mark = m is not None and m not in seen
        if first_mark:
            seen.add(m)
        if S[i] >= tot * min_pct / 100 or first_mark:
            top = sorted(((int(r[idx[st]] or 0), st[6:]) for st in stalls), reverse=True)[:2]
            print(f"{i:5d} {S[i]:6d} {100 * S[i] / tot:5.1f}%  {src[:72]:72s} {top}")


if __name__ == "__main__":
    main()
