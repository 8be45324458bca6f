#!/usr/bin/env python3
"""basic-block instruction shares of an .ncu-rep (development aid): ncu_blocks.py file.ncu-rep units [min_pct]
units = number of work items of the launch (windows, reads): prints instructions per unit"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.8
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ia, ii, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ins = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        try:
            ins.append((r[ia].strip(), int(r[ii]), int(r[isamp])))
        except ValueError:
            pass
    tot = sum(x[1] for x in ins)
    ts = sum(x[2] for x in ins)
    print("%d sass instrs; executed %d = %.1f per unit" % (len(ins), tot, tot / units))
    blocks, cur = [], None
    for i, (s, c, sm) in enumerate(ins):
        if cur is None or c != cur[1]:
            cur = [i, c, 0, 0, {}]
            blocks.append(cur)
        cur[2] += 1
        cur[3] += sm
        op = s.split()[0] if not s.startswith("@") else s.split()[1]
        op = op.split(".")[0]
        cur[4][op] = cur[4].get(op, 0) + 1
    for b in blocks:
        share = 100.0 * b[1] * b[2] / max(tot, 1)
        if share >= minpct:
            ops = sorted(b[4].items(), key=lambda kv: -kv[1])[:7]
            print("idx %5d n=%4d exec/unit=%7.2f  inst %5.1f%% (%6.1f/unit) samp %5.1f%%  %s"
                  % (b[0], b[2], b[1] / units, share, b[1] * b[2] / units, 100.0 * b[3] / max(ts, 1),
                     " ".join("%s:%d" % kv for kv in ops)))


if __name__ == "__main__":
    main()
