#!/usr/bin/env python
"""Aggregate the warp-stall samples of one kernel of an ncu report (`ncu -i X.ncu-rep --page source
--csv > src.csv`) over address regions: share of samples, instructions per warp, FP64 instructions
per warp and the top stall reasons of each region.
usage: ncu_regions.py src.csv name:hexoffset [name:hexoffset ...]   (offsets from the kernel start,
ascending; a region runs to the next offset)"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h, data = rows[1], [r for r in rows[2:] if len(r) > 10]
    ia, isamp, iex, isrc = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    stalls = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    base = int(data[0][ia], 16)
    marks = [(a.split(":")[0], int(a.split(":")[1], 16)) for a in sys.argv[2:]] or [("all", 0)]
    marks.append(("end", 1 << 40))
    nwarp = int(data[0][iex] or 1) or 1          # the first instruction runs once per warp
    tot = sum(int(r[isamp] or 0) for r in data)
    for (name, lo), (_, hi) in zip(marks[:-1], marks[1:]):
        rs = [r for r in data if lo <= int(r[ia], 16) - base < hi]
        s = sum(int(r[isamp] or 0) for r in rs)
        ex = sum(int(r[iex] or 0) for r in rs)
        fp = sum(int(r[iex] or 0) for r in rs
                 if any(op in r[isrc].split(";")[0].split()[:2] for op in ("DFMA", "DMUL", "DADD")))
        st = {n: sum(int(r[i] or 0) for r in rs) for i, n in stalls}
        top = sorted(st.items(), key=lambda x: -x[1])[:5]
        print(f"{name:26s} samples {100 * s / max(tot, 1):5.1f}%  inst/warp {ex / nwarp:7.1f}  fp64/warp {fp / nwarp:7.1f}  "
              + " ".join(f"{n[6:]}:{100 * v / max(s, 1):.0f}%" for n, v in top))


if __name__ == "__main__":
    main()
