#!/usr/bin/env python
"""Per-instruction warp-stall samples of a kernel from an .ncu-rep (ncu --set full): the hottest SASS instructions with their top stall
reasons, and the totals per stall reason.   usage: python tools/ncu_hot.py REPORT.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
    data = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr)]
    i_s, i_src = hdr.index("# Samples"), hdr.index("Source")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[i_s]) for r in data) or 1
    agg = {h: sum(int(r[hdr.index(h)]) for r in data) for h in stalls}
    print(f"{len(data)} instructions, {tot} samples; by reason: " + ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for i in sorted(range(len(data)), key=lambda i: -int(data[i][i_s]))[:top]:
        r = data[i]
        st = sorted(((h[6:], int(r[hdr.index(h)])) for h in stalls), key=lambda kv: -kv[1])[:2]
        print(f"{i:5d} {100 * int(r[i_s]) / tot:5.1f}%  {r[i_src].strip()[:80]:80s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
