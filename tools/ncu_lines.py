#!/usr/bin/env python
"""Warp-stall samples of an .ncu-rep (captured with --import-source on) aggregated per CUDA source line.
Usage: tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    agg = defaultdict(lambda: [0, 0, ""])
    cur_file, hdr = "", None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if not d["Line No"]:
                continue                       # SASS rows under a source line
            try:
                s, ex = int(d["# Samples"]), int(d["Instructions Executed"])
            except (ValueError, KeyError):
                continue
            key = (cur_file, int(d["Line No"]))
            agg[key][0] += s
            agg[key][1] += ex
            if not agg[key][2]:
                agg[key][2] = r[1].strip()[:110]
    tot = sum(v[0] for v in agg.values()) or 1
    for (f, ln), (s, ex, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5.1f%% %10d  %s:%d  %s" % (100.0 * s / tot, ex, f, ln, src))


if __name__ == "__main__":
    main()
