"""Regenerates the golden command-line outputs of tests/test_gpu_cli_dropin.py from the REFERENCE
binary (oracle/_ref/metacache_mpi_ref, built from /root/reference by oracle/Makefile, run as
`mpirun -np 2` through oracle/mpirun_shim.py).  Needs no GPU.  Usage: python tests/golden/make_cli_golden.py"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402

load_package()
import test_gpu_cli_dropin as t  # noqa: E402

CASES = {
    "cli_np2_species_maxcand4.txt": t.QFLAGS,
    "cli_np2_abundances.txt": ["-pairfiles", "-lowest", "species", "-maxcand", "4", "-hitmin", "4", "-hitdiff", "30",
                               "-abundances", "-abundance-per", "species"],
    "cli_np2_abundance_per_genus.txt": ["-pairfiles", "-lowest", "species", "-maxcand", "4", "-hitmin", "4", "-hitdiff", "30",
                                        "-highest", "family", "-abundances", "-abundance-per", "genus"],
    "cli_np2_abundance_two_levels.txt": ["-pairfiles", "-lowest", "species", "-maxcand", "8", "-hitmin", "2", "-hitdiff", "3",
                                         "-abundances", "-abundance-per", "species"],
    "cli_np2_hits_per_seq.txt": ["-pairfiles", "-lowest", "species", "-maxcand", "4", "-hitmin", "4", "-hitdiff", "80", "-hits-per-seq"],
}

# the reference's DEFAULT flags (-lowest sequence, -maxcand 2, hitmin 5, hitdiff 1.0): at np >= 2 its uint32 wire format
# loses the sequence-level taxa of every rank but the receiver; np = 4 adds the reduction tree.
# "mixed": genomes 1 and 4 carry no taxid, so at -lowest species their candidates stay sequence-level while the others are lifted.
COMPAT_CASES = {
    ("cli_np2_default_flags.txt", 2, False): ["-pairfiles", "-tophits", "-queryids"],
    ("cli_np4_default_flags.txt", 4, False): ["-pairfiles", "-tophits", "-queryids"],
    ("cli_np4_mixed_lowest_species.txt", 4, True): t.QFLAGS,
    ("cli_np2_mixed_lowest_species.txt", 2, True): t.QFLAGS,
}

if __name__ == "__main__":
    assert os.path.exists(t.REF), "build the reference first: make -C oracle"
    for (name, np_, mixed), flags in COMPAT_CASES.items():
        with tempfile.TemporaryDirectory() as d:
            t.make_files(d, no_taxid=(1, 4) if mixed else ())
            t.run([sys.executable, t.SHIM, "-np", str(np_), "--", t.REF, "build", "rdb", "genomes", "-taxonomy", "tax"], d)
            t.run([sys.executable, t.SHIM, "-np", str(np_), "--", t.REF, "query", "rdb", "r1.fq", "r2.fq"] + flags +
                  ["-threads", "1", "-query-limit", "300", "-out", "res.txt"], d)
            with open(os.path.join(ROOT, "tests", "golden", name), "w") as f:
                f.write("\n".join(t.stable_lines(os.path.join(d, "res.txt"))) + "\n")
            print("wrote", name)
    with tempfile.TemporaryDirectory() as d:
        # merge mode: the reference's own query results of two half databases, merged by the reference
        t.make_split_files(d)
        for nm in ("A", "B"):
            t.run([sys.executable, t.SHIM, "-np", "2", "--", t.REF, "build", "db" + nm, "g" + nm, "-taxonomy", "tax"], d)
            t.run([sys.executable, t.SHIM, "-np", "2", "--", t.REF, "query", "db" + nm, "r1.fq", "r2.fq", "-pairfiles"] + t.MERGE_FLAGS +
                  ["-threads", "1", "-query-limit", "300", "-out", "res%s.txt" % nm], d)
        t.run([sys.executable, t.SHIM, "-np", "1", "--", t.REF, "merge", "resA.txt", "resB.txt", "-taxonomy", "tax"] + t.MERGE_FLAGS +
              ["-out", "merged.txt"], d)
        with open(os.path.join(ROOT, "tests", "golden", "cli_merge_two_databases.txt"), "w") as f:
            f.write("\n".join(t.stable_lines(os.path.join(d, "merged.txt"))) + "\n")
        print("wrote cli_merge_two_databases.txt")
    with tempfile.TemporaryDirectory() as d:
        t.make_assembly_summary_files(d)
        t.run([sys.executable, t.SHIM, "-np", "2", "--", t.REF, "build", "rdb", "genomes", "-taxonomy", "tax"], d)
        t.run([sys.executable, t.SHIM, "-np", "2", "--", t.REF, "query", "rdb", "r1.fq", "r2.fq"] + t.QFLAGS +
              ["-threads", "1", "-query-limit", "300", "-out", "res.txt"], d)
        with open(os.path.join(ROOT, "tests", "golden", "cli_np2_assembly_summary.txt"), "w") as f:
            f.write("\n".join(t.stable_lines(os.path.join(d, "res.txt"))) + "\n")
        print("wrote cli_np2_assembly_summary.txt")
    with tempfile.TemporaryDirectory() as d:
        t.make_unranked_files(d)
        t.run([sys.executable, t.SHIM, "-np", "2", "--", t.REF, "build", "rdb", "genomes", "-taxonomy", "tax"], d)
        t.run([sys.executable, t.SHIM, "-np", "2", "--", t.REF, "query", "rdb", "r1.fq", "r2.fq"] + t.QFLAGS +
              ["-threads", "1", "-query-limit", "300", "-out", "res.txt"], d)
        with open(os.path.join(ROOT, "tests", "golden", "cli_np2_rank_unranked.txt"), "w") as f:
            f.write("\n".join(t.stable_lines(os.path.join(d, "res.txt"))) + "\n")
        print("wrote cli_np2_rank_unranked.txt")
    with tempfile.TemporaryDirectory() as d:
        t.make_files(d)
        t.run([sys.executable, t.SHIM, "-np", "2", "--", t.REF, "build", "rdb", "genomes", "-taxonomy", "tax"], d)
        for name, flags in CASES.items():
            t.run([sys.executable, t.SHIM, "-np", "2", "--", t.REF, "query", "rdb", "r1.fq", "r2.fq"] + flags +
                  ["-threads", "1", "-query-limit", "300", "-out", "res.txt"], d)
            with open(os.path.join(ROOT, "tests", "golden", name), "w") as f:
                f.write("\n".join(t.stable_lines(os.path.join(d, "res.txt"))) + "\n")
            print("wrote", name)
