"""BASELINE.json configs[0] end to end through the two command lines (development aid / evidence):
10 synthetic genomes x 5 Mbp, 50 000 read pairs 2 x 150 bp, the authors' query flags, P = 2 shards.
Runs metacache-mpi_b200/metacache_b200 (GPU) and oracle/_ref/metacache_mpi_ref (reference, CPU,
`mpirun -np 2` through the FIFO shim), checks that the query outputs are identical and prints the
wall times.  Needs oracle/_ref (built by `python __graft_entry__.py` where /root/reference exists)."""
import os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
load_package()
from metacache_mpi_b200 import synth

EXE = os.path.join(ROOT, "metacache-mpi_b200", "metacache_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "metacache_mpi_ref")
SHIM = os.path.join(ROOT, "oracle", "mpirun_shim.py")
QFLAGS = ["-pairfiles", "-lowest", "species", "-maxcand", "4", "-hitmin", "4", "-hitdiff", "80", "-tophits", "-queryids"]
G = int(sys.argv[1]) if len(sys.argv) > 1 else 10
L = 5_000_000
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000


def run(cmd, cwd):
    t0 = time.time()
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    dt = time.time() - t0
    if r.returncode != 0:
        print(r.stdout.decode()[-3000:])
        raise SystemExit("failed: %s" % " ".join(cmd))
    return dt, r.stdout.decode()


def stable(path):
    return [l for l in open(path).read().splitlines() if not (l.startswith("# time:") or l.startswith("# speed:"))]


d = tempfile.mkdtemp(prefix="c1_")
genomes = synth.make_genomes(G, L)
ids, par, rk, species = synth.make_taxonomy(G)
synth.write_genome_files(os.path.join(d, "genomes"), genomes, species)
synth.write_taxonomy_files(os.path.join(d, "tax"), ids, par, rk)
b1, o1, b2, o2, t = synth.make_read_pairs(genomes, NP, seed=7)
synth.write_fastq_pair(os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq"), b1, o1, b2, o2, t)
print("files in", d, flush=True)
res = {}
res["gpu build"], out_b = run([EXE, "build", "gdb", "genomes", "-taxonomy", "tax", "-gpus", "2"], d)
res["gpu query"], out_q = run([EXE, "query", "gdb", "r1.fq", "r2.fq"] + QFLAGS + ["-out", "res_gpu.txt"], d)
for l in (out_b + out_q).splitlines():
    if l.startswith("[metacache_b200]"):
        print("  ", l)
for k in ("gpu build", "gpu query"):
    print("%-14s %7.2f s wall" % (k, res[k]), flush=True)
if os.path.exists(REF):
    cores = os.cpu_count() or 1
    T = max(1, min(8, cores // 2))
    res["ref build"], out = run([sys.executable, SHIM, "-np", "2", "--", REF, "build", "rdb", "genomes", "-taxonomy", "tax"], d)
    for l in out.splitlines():
        if "Added" in l or "Total build time" in l:
            print("   reference:", l.strip())
    res["ref query"], out = run([sys.executable, SHIM, "-np", "2", "--", REF, "query", "rdb", "r1.fq", "r2.fq"] + QFLAGS +
                                ["-threads", str(T), "-query-limit", str((NP + T - 1) // T), "-out", "res_ref.txt"], d)
    for k in ("ref build", "ref query"):
        print("%-14s %7.2f s wall (np=2%s)" % (k, res[k], ", %d threads per rank" % T if "query" in k else ""), flush=True)
    for l in open(os.path.join(d, "res_ref.txt")).read().splitlines():
        if l.startswith("# time:") or l.startswith("# speed:"):
            print("   reference:", l)
    # the reference prints reads in thread-completion order: compare as sets of per-read lines
    a = sorted(l for l in stable(os.path.join(d, "res_gpu.txt")) if not l.startswith("#"))
    b = sorted(l for l in stable(os.path.join(d, "res_ref.txt")) if not l.startswith("#"))
    print("per-read output lines identical:", a == b, "(%d reads)" % len(a))
    sa = [l for l in stable(os.path.join(d, "res_gpu.txt")) if l.startswith("# ") and ("classified" in l or "%" in l)]
    sb = [l for l in stable(os.path.join(d, "res_ref.txt")) if l.startswith("# ") and ("classified" in l or "%" in l)]
    print("classification summary identical:", sa == sb)
    print("speed-up build %.1fx, query %.1fx (wall, whole command lines)" % (res["ref build"] / res["gpu build"], res["ref query"] / res["gpu query"]))
