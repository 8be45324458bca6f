"""BASELINE.json configs[0] at its stated size, end to end through the two command lines:
10 synthetic genomes x 5 Mbp (50 Mbp), 50 000 read pairs 2 x 150 bp (100 k reads), the authors' query flags.
  * `metacache_b200 build/query -gpus 2`  vs  the REFERENCE BINARY under `mpirun -np 2` (oracle/_ref, MPI shim):
    shard files (header, taxa, every bucket), per-read `-tophits -queryids` lines, classification summary;
    each side loads the other's shard files and prints the same output;
  * `metacache_b200 -gpus 1`  vs  the oracle at P = 1 (the reference classifies nothing at -np 1).
Without oracle/_ref (it travels with the repo to the GPU box) the reference half is skipped loudly."""
import os
import subprocess
import sys

import numpy as np
import pytest

from metacache_mpi_b200 import synth, taxonomy
from oracle import pyoracle as po

import clifmt

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "metacache-mpi_b200", "metacache_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "metacache_mpi_ref")
SHIM = os.path.join(ROOT, "oracle", "mpirun_shim.py")
QFLAGS = ["-pairfiles", "-lowest", "species", "-maxcand", "4", "-hitmin", "4", "-hitdiff", "80", "-tophits", "-queryids"]
G, L, NPAIRS = 10, 5_000_000, 50_000


ENV = dict(os.environ, MCG_NCCL_MIN_READS="0")      # one GPU per shard: through the NCCL path also at this size


def run(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=1200, env=ENV)
    assert r.returncode == 0, r.stdout.decode()[-3000:]
    return r.stdout.decode()


def read_lines(path):
    return sorted(l for l in open(path).read().splitlines() if not l.startswith("#"))


def summary(path):
    return [l for l in open(path).read().splitlines() if l.startswith("# ") and ("classified" in l or "% (" in l or l.startswith("# queries"))]


@pytest.fixture(scope="module")
def world(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("c0"))
    genomes = synth.make_genomes(G, L)
    ids, par, rk, species = synth.make_taxonomy(G)
    synth.write_genome_files(os.path.join(d, "genomes"), genomes, species)
    synth.write_taxonomy_files(os.path.join(d, "tax"), ids, par, rk)
    b1, o1, b2, o2, truth = synth.make_read_pairs(genomes, NPAIRS, seed=7)
    synth.write_fastq_pair(os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq"), b1, o1, b2, o2, truth)
    return dict(d=d, genomes=genomes, tabs=taxonomy.TaxonomyTables(ids, par, rk, species), reads=(b1, o1, b2, o2), truth=truth)


def test_config0_two_shards_vs_reference_binary(world):
    d = world["d"]
    assert os.path.exists(EXE), "host driver not built (python __graft_entry__.py)"
    run([EXE, "build", "gdb", "genomes", "-taxonomy", "tax", "-gpus", "2"], d)
    run([EXE, "query", "gdb", "r1.fq", "r2.fq"] + QFLAGS + ["-out", "res_gpu.txt"], d)
    mine = read_lines(os.path.join(d, "res_gpu.txt"))
    assert len(mine) == NPAIRS
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/metacache_mpi_ref not present: reference half not run")
    run([sys.executable, SHIM, "-np", "2", "--", REF, "build", "rdb", "genomes", "-taxonomy", "tax"], d)
    T = max(1, min(8, (os.cpu_count() or 2) // 2))
    ref_q = ["-threads", str(T), "-query-limit", str((NPAIRS + T - 1) // T)]
    run([sys.executable, SHIM, "-np", "2", "--", REF, "query", "rdb", "r1.fq", "r2.fq"] + QFLAGS + ref_q + ["-out", "res_ref.txt"], d)
    # the reference prints reads in thread-completion order: compare the sorted per-read lines
    ref = read_lines(os.path.join(d, "res_ref.txt"))
    assert mine == ref
    assert summary(os.path.join(d, "res_gpu.txt")) == summary(os.path.join(d, "res_ref.txt"))
    # shard files: same parameters, taxa and buckets (the reference writes buckets in hash-table order)
    for r in range(2):
        out = run([EXE, "shards-equal", "gdb.db_%d" % r, "rdb.db_%d" % r], d)
        assert "shards equal" in out, out
    # each side reads the other's shards
    run([sys.executable, SHIM, "-np", "2", "--", REF, "query", "gdb", "r1.fq", "r2.fq"] + QFLAGS + ref_q + ["-out", "res_ref_on_gpu_db.txt"], d)
    assert read_lines(os.path.join(d, "res_ref_on_gpu_db.txt")) == ref
    run([EXE, "query", "rdb", "r1.fq", "r2.fq"] + QFLAGS + ["-out", "res_gpu_on_ref_db.txt"], d)
    assert read_lines(os.path.join(d, "res_gpu_on_ref_db.txt")) == ref


def test_config0_one_shard_vs_oracle(world, oracle):
    d = world["d"]
    run([EXE, "build", "gdb1", "genomes", "-taxonomy", "tax", "-gpus", "1"], d)
    run([EXE, "query", "gdb1", "r1.fq", "r2.fq"] + QFLAGS + ["-out", "res_gpu1.txt"], d)
    got = [l for l in open(os.path.join(d, "res_gpu1.txt")).read().splitlines() if not l.startswith("#")]
    # the oracle needs only the lists of the reads' features: restricted build, all threads (exact, see mc_oracle.h)
    b1, o1, b2, o2 = world["reads"]
    f1, _ = oracle.sketch_sequences(b1, o1)
    f2, _ = oracle.sketch_sequences(b2, o2)
    allowed = np.unique(np.concatenate([f1.ravel(), f2.ravel()]))
    allowed = allowed[allowed != 0xFFFFFFFF]
    buf, offs = po.concat(world["genomes"])
    odb = oracle.db_create(shard=0, nshards=1)
    odb.add_targets_restricted(buf, offs, np.arange(G, dtype=np.uint32), allowed, threads=os.cpu_count() or 1)
    odb.finalize()
    tabs = world["tabs"]
    tax = po.TaxTables(tabs.lineage, tabs.rank, G)
    ot, oh, on, best, brank, st = oracle.query_batch([odb], tax, po.make_qopt(maxcand=4, lowest=4, hitmin=4, hitdiff=0.8), b1, o1, b2, o2)
    want = clifmt.format_lines(tabs, world["truth"], ot, oh, on, best, lowest=4)
    assert got == want
    assert clifmt.golden_summary(os.path.join(d, "res_gpu1.txt")) == clifmt.stats_summary(st, NPAIRS)
