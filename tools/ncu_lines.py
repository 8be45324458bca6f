#!/usr/bin/env python3
"""per-source-line instruction and stall-sample shares of an .ncu-rep captured with
--import-source on (development aid): ncu_lines.py file.ncu-rep [min_pct]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname = ""
    hdr = None
    agg = {}
    order = []
    cur = None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 4 and r[0] == "Line No":
            hdr = r
            ia = hdr.index("Address")
            ii = hdr.index("Instructions Executed")
            isamp = hdr.index("# Samples")
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        if r[0] != "":
            cur = (fname, int(r[0]), r[1].strip())
            if cur not in agg:
                agg[cur] = [0, 0, 0]
                order.append(cur)
            continue
        if cur is None or r[ia] in ("", "..."):
            continue
        try:
            agg[cur][0] += int(r[ii])
            agg[cur][1] += int(r[isamp])
            agg[cur][2] += 1
        except ValueError:
            pass
    ti = sum(v[0] for v in agg.values())
    ts = sum(v[1] for v in agg.values())
    print("total warp instructions %d, samples %d" % (ti, ts))
    for key in order:
        v = agg[key]
        pi, ps = 100.0 * v[0] / max(ti, 1), 100.0 * v[1] / max(ts, 1)
        if pi >= minpct or ps >= minpct:
            print("%-16s %4d  inst %5.1f%%  samples %5.1f%%  sass %4d | %s" % (key[0], key[1], pi, ps, v[2], key[2][:90]))


if __name__ == "__main__":
    main()
