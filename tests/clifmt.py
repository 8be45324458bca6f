"""Per-read output lines of the reference's `query ... -tophits -queryids` (classification.cpp:583-632,
printing.cpp:305-360) formatted from result arrays, so that array-level results (oracle, C-ABI) can be
compared with what the reference BINARY printed (tests/golden/cli_*.txt)."""
import numpy as np

from metacache_mpi_b200 import synth

COL = "\t|\t"


def taxon_name(tabs, idx):
    """names as make_files() of test_gpu_cli_dropin.py / synth.write_* produce them"""
    if idx < tabs.ntargets:
        return "SYN%06d.1" % idx
    return "taxon%d" % int(tabs.taxid[idx])


def format_lines(tabs, truth, ot, oh, on, best, lowest, highest=19, first_id=1):
    """one line per read pair: qid | header | tophits | classification"""
    out = []
    for i in range(len(on)):
        tops = []
        for c in range(int(on[i])):
            if oh[i, c] == 0:
                break
            t = int(ot[i, c])
            if lowest == 0:
                tops.append("%s:%d" % (taxon_name(tabs, t), oh[i, c]))
            else:
                a = int(tabs.lineage[t, lowest]) if tabs.rank[t] < lowest else t
                tops.append("%s:%d" % (int(tabs.taxid[a]) if a >= 0 else taxon_name(tabs, t), oh[i, c]))
        b = int(best[i])
        if b < 0 or tabs.rank[b] > highest:
            cls = "--"
        else:
            r = max(lowest, int(tabs.rank[b]))
            x = int(tabs.lineage[b, r])
            cls = "%s:%s" % (synth.RANK_NAMES[r], taxon_name(tabs, x) if x >= 0 else "--")
        out.append(COL.join(["%d" % (first_id + i), "read%d_g%d/1" % (i, int(truth[i])), ",".join(tops), cls]))
    return out


def golden_read_lines(path):
    return [l for l in open(path).read().splitlines() if not l.startswith("#") and COL in l]


def golden_summary(path):
    """{'unclassified': n, rank name: n, ...} from the '# classified' block"""
    out = {}
    for l in open(path).read().splitlines():
        if l.startswith("# unclassified:") or (l.startswith("#   ") and "% (" in l):
            name = l[2:].split()[0].rstrip(":")
            out[name] = int(l[l.rindex("(") + 1:l.rindex(")")])
    return out


def stats_summary(stats, total):
    """the same dictionary from the 22 counters (printing.cpp:522-641 prints these ranks only)"""
    out = {}
    if stats[21] > 0:
        out["unclassified"] = int(stats[21])
    for r in (0, 3, 4, 6, 10, 12, 14, 16, 18, 19, 20):
        if stats[r] > 0:
            out[synth.RANK_NAMES[r]] = int(stats[r])
    return out
