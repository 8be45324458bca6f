#!/usr/bin/env python3
"""prints selected metrics of an .ncu-rep (development aid): ncu_metrics.py file.ncu-rep [regex]"""
import csv
import re
import subprocess
import sys

DEFAULT = (r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|gpu__dram_throughput.avg.pct|smsp__issue_active.avg.pct|"
           r"smsp__inst_executed.sum$|sm__inst_executed_pipe_(alu|fma|fmaheavy|fmalite|lsu|xu|uniform|adu|cbu)\.avg.pct_of_peak_sustained_active|"
           r"sm__pipe_(alu|fma|fmaheavy|fmalite|xu)_cycles_active.avg.pct_of_peak_sustained_active|sm__warps_active.avg.pct|"
           r"launch__(registers_per_thread|grid_size|block_size|occupancy_limit)|smsp__average_warp.*issue_stalled.*_per_warp_active|"
           r"smsp__warp_issue_stalled.*_per_warp_active|sm__cycles_elapsed.avg$|smsp__thread_inst_executed_per_inst_executed|"
           r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|smsp__inst_executed_op_shared|sm__inst_executed.avg.per_cycle_active")

def main():
    rep = sys.argv[1]
    rx = re.compile(sys.argv[2] if len(sys.argv) > 2 else DEFAULT)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[head.index("Kernel Name")] if "Kernel Name" in head else "")
        for i, k in enumerate(head):
            if rx.search(k):
                print("  %-90s %-12s %s" % (k, units[i], r[i]))

if __name__ == "__main__":
    main()
