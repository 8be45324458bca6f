"""Drop-in check of the C++ host driver (metacache-mpi_b200/metacache_b200) against the REAL
reference binary (oracle/_ref/metacache_mpi_ref under the MPI shim) on the same files:
  * shard files `<db>.db_<r>`: same header, taxonomy records and bucket contents; each side can
    load the other's shards and produces the same query output;
  * query output (`-tophits -queryids`): identical per-read lines, TABLE_LAYOUT, parameter header
    and classification summary (only the time / speed lines differ).
Without oracle/_ref the committed golden output of the reference is used."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

from metacache_mpi_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "metacache-mpi_b200", "metacache_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "metacache_mpi_ref")
SHIM = os.path.join(ROOT, "oracle", "mpirun_shim.py")
GOLDEN = os.path.join(ROOT, "tests", "golden", "cli_np2_species_maxcand4.txt")
QFLAGS = ["-pairfiles", "-lowest", "species", "-maxcand", "4", "-hitmin", "4", "-hitdiff", "80", "-tophits", "-queryids"]


def make_files(d, no_taxid=()):
    G, L = 6, 30000
    genomes = synth.make_genomes(G, L)
    ids, par, rk, species = synth.make_taxonomy(G)
    species = species.copy()
    for g in no_taxid:                 # header without a usable taxid: the target hangs below no species
        species[g] = 0
    synth.write_genome_files(os.path.join(d, "genomes"), genomes, species)
    synth.write_taxonomy_files(os.path.join(d, "tax"), ids, par, rk)
    b1, o1, b2, o2, t = synth.make_read_pairs(genomes, 300, seed=5)
    synth.write_fastq_pair(os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq"), b1, o1, b2, o2, t)


def parse_shard(path):
    d = open(path, "rb").read()
    o = 0

    def rd(fmt):
        nonlocal o
        v = struct.unpack_from("<" + fmt, d, o)
        o += struct.calcsize("<" + fmt)
        return v

    def rds():
        nonlocal o
        n, = rd("Q")
        s = d[o:o + n]
        o += n
        return s

    head = rd("Q") + rd("6B") + rd("QQQQ") + rd("QQQQ") + rd("Q")
    nt, = rd("Q")
    taxa = []
    for _ in range(nt):
        tid, par, rank = rd("qqB")
        name, fn = rds(), rds()
        idx, win = rd("QQ")
        taxa.append((tid, par, rank, name, os.path.basename(fn), idx, win))
    ntgt, = rd("I")
    nk, nv = rd("QQ")
    buckets = {}
    for _ in range(nk):
        key, n = rd("IB")
        n1, = rd("Q")
        tg = rd("%dI" % n1)
        n2, = rd("Q")
        wn = rd("%dI" % n2)
        assert n == n1 == n2
        buckets[key] = (tg, wn)
    assert o == len(d), "trailing bytes in shard file"
    return head, taxa, ntgt, nv, buckets


def stable_lines(path):
    out = []
    for l in open(path).read().splitlines():
        if l.startswith("# time:") or l.startswith("# speed:"):
            continue
        out.append(l)
    return out


# on a box with one GPU per shard the driver runs the NCCL path of libmcgpu (mcg_query_batch_sharded per GPU thread);
# its small-input shortcut (no NCCL below 256 k reads) is switched off so that these small cases go through it
ENV = dict(os.environ, MCG_NCCL_MIN_READS="0")


def run(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600, env=ENV)
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    return r.stdout.decode()


def test_cli_build_query_dropin(tmp_path):
    d = str(tmp_path)
    make_files(d)
    assert os.path.exists(EXE), "host driver not built (python __graft_entry__.py)"
    run([EXE, "build", "gdb", "genomes", "-taxonomy", "tax", "-gpus", "2"], d)
    run([EXE, "query", "gdb", "r1.fq", "r2.fq"] + QFLAGS + ["-out", "res_gpu.txt"], d)
    mine = stable_lines(os.path.join(d, "res_gpu.txt"))
    if os.path.exists(GOLDEN):
        assert mine == open(GOLDEN).read().splitlines()
    if not os.path.exists(REF):
        assert os.path.exists(GOLDEN), "neither oracle/_ref nor the golden output is available"
        return
    # the reference itself, np = 2
    run([sys.executable, SHIM, "-np", "2", "--", REF, "build", "rdb", "genomes", "-taxonomy", "tax"], d)
    run([sys.executable, SHIM, "-np", "2", "--", REF, "query", "rdb", "r1.fq", "r2.fq"] + QFLAGS +
        ["-threads", "1", "-query-limit", "300", "-out", "res_ref.txt"], d)
    ref = stable_lines(os.path.join(d, "res_ref.txt"))
    assert mine == ref
    # shard files: same content
    for r in range(2):
        a = parse_shard(os.path.join(d, "gdb.db_%d" % r))
        b = parse_shard(os.path.join(d, "rdb.db_%d" % r))
        assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2] and a[3] == b[3]
        assert a[4] == b[4]
    # cross-loading: the reference reads GPU-built shards, the GPU driver reads reference-built shards
    run([sys.executable, SHIM, "-np", "2", "--", REF, "query", "gdb", "r1.fq", "r2.fq"] + QFLAGS +
        ["-threads", "1", "-query-limit", "300", "-out", "res_ref_on_gpu_db.txt"], d)
    assert stable_lines(os.path.join(d, "res_ref_on_gpu_db.txt")) == ref
    run([EXE, "query", "rdb", "r1.fq", "r2.fq"] + QFLAGS + ["-out", "res_gpu_on_ref_db.txt"], d)
    assert stable_lines(os.path.join(d, "res_gpu_on_ref_db.txt")) == ref


def test_cli_defaults_and_overpopulated(tmp_path):
    """reference defaults (-maxcand 2, hitmin 5, hitdiff 1.0) at -lowest genus with lineage + taxids,
    and a capped database with -remove-overpopulated-features"""
    d = str(tmp_path)
    make_files(d)
    run([EXE, "build", "gdb", "genomes", "-taxonomy", "tax", "-gpus", "2", "-max-locations-per-feature", "2",
         "-remove-overpopulated-features"], d)
    flags = ["-pairfiles", "-lowest", "genus", "-tophits", "-queryids", "-lineage", "-taxids"]
    run([EXE, "query", "gdb", "r1.fq", "r2.fq"] + flags + ["-out", "res_gpu.txt"], d)
    if not os.path.exists(REF):
        return
    run([sys.executable, SHIM, "-np", "2", "--", REF, "build", "rdb", "genomes", "-taxonomy", "tax",
         "-max-locations-per-feature", "2", "-remove-overpopulated-features"], d)
    run([sys.executable, SHIM, "-np", "2", "--", REF, "query", "rdb", "r1.fq", "r2.fq"] + flags +
        ["-threads", "1", "-query-limit", "300", "-out", "res_ref.txt"], d)
    assert stable_lines(os.path.join(d, "res_gpu.txt")) == stable_lines(os.path.join(d, "res_ref.txt"))
    for r in range(2):
        a = parse_shard(os.path.join(d, "gdb.db_%d" % r))
        b = parse_shard(os.path.join(d, "rdb.db_%d" % r))
        assert a[4] == b[4] and a[3] == b[3]


def test_cli_abundances(tmp_path):
    """-abundances / -abundance-per (classification.cpp:362-428, printing.cpp:474-516): per-taxon read
    counts and the estimate at species level, in the mapping output and in a file of their own.
    -hitdiff 30 leaves most reads at genus level, so the estimate has counts to distribute."""
    d = str(tmp_path)
    make_files(d)
    run([EXE, "build", "gdb", "genomes", "-taxonomy", "tax", "-gpus", "2"], d)
    base = ["-pairfiles", "-lowest", "species", "-maxcand", "4", "-hitmin", "4", "-hitdiff", "30"]
    inline = base + ["-abundances", "-abundance-per", "species"]
    tofile = base + ["-abundance-per", "species", "-abundances"]
    run([EXE, "query", "gdb", "r1.fq", "r2.fq"] + inline + ["-out", "res_gpu.txt"], d)
    run([EXE, "query", "gdb", "r1.fq", "r2.fq"] + tofile + ["ab_gpu.txt", "-out", "res_gpu2.txt"], d)
    mine = stable_lines(os.path.join(d, "res_gpu.txt"))
    golden = os.path.join(ROOT, "tests", "golden", "cli_np2_abundances.txt")
    if os.path.exists(golden):
        assert mine == open(golden).read().splitlines()
    assert any(l.startswith("# estimated abundance") for l in mine)
    tail = mine[[i for i, l in enumerate(mine) if l.startswith("# query summary")][0]:]
    tail = tail[:[i for i, l in enumerate(tail) if l.startswith("# queries:")][0]]
    assert open(os.path.join(d, "ab_gpu.txt")).read().splitlines() == tail
    if not os.path.exists(REF):
        assert os.path.exists(golden), "neither oracle/_ref nor the golden output is available"
        return
    run([sys.executable, SHIM, "-np", "2", "--", REF, "build", "rdb", "genomes", "-taxonomy", "tax"], d)
    ref_tail = ["-threads", "1", "-query-limit", "300"]
    run([sys.executable, SHIM, "-np", "2", "--", REF, "query", "rdb", "r1.fq", "r2.fq"] + inline + ref_tail + ["-out", "res_ref.txt"], d)
    run([sys.executable, SHIM, "-np", "2", "--", REF, "query", "rdb", "r1.fq", "r2.fq"] + tofile + ["ab_ref.txt"] + ref_tail +
        ["-out", "res_ref2.txt"], d)
    assert mine == stable_lines(os.path.join(d, "res_ref.txt"))
    assert stable_lines(os.path.join(d, "res_gpu2.txt")) == stable_lines(os.path.join(d, "res_ref2.txt"))
    assert open(os.path.join(d, "ab_gpu.txt")).read() == open(os.path.join(d, "ab_ref.txt")).read()


def test_cli_hits_per_seq(tmp_path):
    """-hits-per-seq: the per-target lists are dead code in this fork of the reference:
    query ids are forced on, two parameter lines appear and the table header is printed without rows --
    in the mapping output or in a file of its own"""
    d = str(tmp_path)
    make_files(d)
    run([EXE, "build", "gdb", "genomes", "-taxonomy", "tax", "-gpus", "2"], d)
    flags = ["-pairfiles", "-lowest", "species", "-maxcand", "4", "-hitmin", "4", "-hitdiff", "80", "-hits-per-seq"]
    run([EXE, "query", "gdb", "r1.fq", "r2.fq"] + flags + ["-out", "res_gpu.txt"], d)
    run([EXE, "query", "gdb", "r1.fq", "r2.fq"] + flags + ["hps_gpu.txt", "-out", "res_gpu2.txt"], d)
    mine = stable_lines(os.path.join(d, "res_gpu.txt"))
    golden = os.path.join(ROOT, "tests", "golden", "cli_np2_hits_per_seq.txt")
    assert mine == open(golden).read().splitlines()
    hdr = [l for l in mine if "list of hits for each reference sequence" in l or "window start position" in l or
           l.startswith("# TABLE_LAYOUT:  sequence")]
    assert len(hdr) == 3 and open(os.path.join(d, "hps_gpu.txt")).read().splitlines() == hdr
    assert [l for l in stable_lines(os.path.join(d, "res_gpu2.txt")) if l not in hdr] == [l for l in mine if l not in hdr]
    if not os.path.exists(REF):
        return
    run([sys.executable, SHIM, "-np", "2", "--", REF, "build", "rdb", "genomes", "-taxonomy", "tax"], d)
    run([sys.executable, SHIM, "-np", "2", "--", REF, "query", "rdb", "r1.fq", "r2.fq"] + flags +
        ["-threads", "1", "-query-limit", "300", "-out", "res_ref.txt"], d)
    assert mine == stable_lines(os.path.join(d, "res_ref.txt"))


@pytest.mark.parametrize("golden,P,no_taxid,flags", [
    ("cli_np2_default_flags.txt", 2, (), ["-pairfiles", "-tophits", "-queryids"]),
    ("cli_np4_default_flags.txt", 4, (), ["-pairfiles", "-tophits", "-queryids"]),
    ("cli_np2_mixed_lowest_species.txt", 2, (1, 4), QFLAGS),
    ("cli_np4_mixed_lowest_species.txt", 4, (1, 4), QFLAGS),
])
def test_cli_reference_compat(tmp_path, golden, P, no_taxid, flags):
    """`-reference-compat`: byte-identical output under the reference's DEFAULT flags (-lowest sequence), where
    its uint32 wire format drops the sequence-level candidates of all ranks but the receiver (quirk Q1), at
    np = 2 and along the reduction tree at np = 4; goldens printed by the reference binary"""
    d = str(tmp_path)
    make_files(d, no_taxid=no_taxid)
    run([EXE, "build", "gdb", "genomes", "-taxonomy", "tax", "-gpus", str(P)], d)
    run([EXE, "query", "gdb", "r1.fq", "r2.fq"] + flags + ["-reference-compat", "-out", "res_gpu.txt"], d)
    mine = stable_lines(os.path.join(d, "res_gpu.txt"))
    assert mine == open(os.path.join(ROOT, "tests", "golden", golden)).read().splitlines()
    if os.path.exists(REF):
        run([sys.executable, SHIM, "-np", str(P), "--", REF, "build", "rdb", "genomes", "-taxonomy", "tax"], d)
        run([sys.executable, SHIM, "-np", str(P), "--", REF, "query", "rdb", "r1.fq", "r2.fq"] + flags +
            ["-threads", "1", "-query-limit", "300", "-out", "res_ref.txt"], d)
        assert mine == stable_lines(os.path.join(d, "res_ref.txt"))
    # without the switch the restated semantics classify the reads of the other shards' genomes as well
    run([EXE, "query", "gdb", "r1.fq", "r2.fq"] + flags + ["-out", "res_gpu_plain.txt"], d)
    assert stable_lines(os.path.join(d, "res_gpu_plain.txt")) != mine


def test_cli_unlimited_maxcand(tmp_path):
    """-maxcand 0 = unlimited (query_options.cpp:176-178): accepted, lists up to 32 entries on the device; with 6
    genomes no list is longer, so the output equals the reference's"""
    d = str(tmp_path)
    make_files(d)
    run([EXE, "build", "gdb", "genomes", "-taxonomy", "tax", "-gpus", "2"], d)
    flags = ["-pairfiles", "-lowest", "species", "-maxcand", "0", "-hitmin", "4", "-hitdiff", "80", "-tophits", "-queryids"]
    run([EXE, "query", "gdb", "r1.fq", "r2.fq"] + flags + ["-out", "res_gpu.txt"], d)
    mine = stable_lines(os.path.join(d, "res_gpu.txt"))
    assert any("18446744073709551615 classification candidates" in l for l in mine)
    if os.path.exists(REF):
        run([sys.executable, SHIM, "-np", "2", "--", REF, "build", "rdb", "genomes", "-taxonomy", "tax"], d)
        run([sys.executable, SHIM, "-np", "2", "--", REF, "query", "rdb", "r1.fq", "r2.fq"] + flags +
            ["-threads", "1", "-query-limit", "300", "-out", "res_ref.txt"], d)
        assert mine == stable_lines(os.path.join(d, "res_ref.txt"))


def make_unranked_files(d):
    """no genome carries a taxid in its header; NCBI-style accession -> taxon tables in the taxonomy directory name
    four of them by accession.version, one by accession only (another version), one not at all"""
    make_files(d, no_taxid=range(6))
    with open(os.path.join(d, "tax", "nucl_gb.accession2taxid"), "w") as f:
        f.write("accession\taccession.version\ttaxid\tgi\n")
        f.write("XX000009\tXX000009.1\t100005\t77\n")
        for g in range(4):
            f.write("SYN%06d\tSYN%06d.1\t%d\t%d\n" % (g, g, 100000 + g, 1000 + g))
    with open(os.path.join(d, "tax", "zz_more.accession2taxid"), "w") as f:
        f.write("accession\taccession.version\ttaxid\tgi\n")
        f.write("SYN000004\tSYN000004.7\t100004\t1004\n")


def test_cli_rank_unranked_targets(tmp_path):
    """targets without a taxon in their header get their parent from *.accession2taxid files of the taxonomy directory
    (mode_build.cpp:493-522): same shard files (taxon records) and same classification as the reference"""
    d = str(tmp_path)
    make_unranked_files(d)
    out = run([EXE, "build", "gdb", "genomes", "-taxonomy", "tax", "-gpus", "2"], d)
    assert "6 targets are unranked." in out and "1 targets remain unranked." in out
    # (genome 5 stays unranked, i.e. a sequence-level candidate of shard 1: the reference loses it on its uint32 wire,
    # SURVEY 3.4 Q1 -- the comparison therefore runs with the reference-compatible merge)
    run([EXE, "query", "gdb", "r1.fq", "r2.fq"] + QFLAGS + ["-reference-compat", "-out", "res_gpu.txt"], d)
    mine = stable_lines(os.path.join(d, "res_gpu.txt"))
    golden = os.path.join(ROOT, "tests", "golden", "cli_np2_rank_unranked.txt")
    if os.path.exists(golden):
        assert mine == open(golden).read().splitlines()
    if not os.path.exists(REF):
        assert os.path.exists(golden), "neither oracle/_ref nor the golden output is available"
        return
    run([sys.executable, SHIM, "-np", "2", "--", REF, "build", "rdb", "genomes", "-taxonomy", "tax"], d)
    run([sys.executable, SHIM, "-np", "2", "--", REF, "query", "rdb", "r1.fq", "r2.fq"] + QFLAGS +
        ["-threads", "1", "-query-limit", "300", "-out", "res_ref.txt"], d)
    assert mine == stable_lines(os.path.join(d, "res_ref.txt"))
    for r in range(2):
        assert "shards equal" in run([EXE, "shards-equal", "gdb.db_%d" % r, "rdb.db_%d" % r], d)


def make_assembly_summary_files(d):
    """an NCBI RefSeq style download: no taxid in any header; `assembly_summary.txt` next to the sequence files maps
    assembly accessions (= the start of the file names) to taxa; genome 1 is named by its sequence accession, genome 4
    has a row with a longer key (prefix match), genome 5 has no row and keeps the taxid of its header"""
    G, L = 6, 30000
    genomes = synth.make_genomes(G, L)
    ids, par, rk, species = synth.make_taxonomy(G)
    gd = os.path.join(d, "genomes")
    os.makedirs(gd, exist_ok=True)
    for g in range(G):
        head = "NC_%06d.1 synthetic genome %d" % (900 + g, g)
        if g == 5:
            head = "NC_%06d.1 taxid=%d synthetic genome %d" % (900 + g, species[g], g)
        name = "GCF_%09d.%d_ASM%dv1_genomic.fna" % (100 + g, 2 if g == 4 else 1, g)
        synth.write_fasta(os.path.join(gd, name), head, genomes[g])
    with open(os.path.join(gd, "assembly_summary.txt"), "w") as f:
        f.write("#   See ftp://ftp.ncbi.nlm.nih.gov/genomes/README_assembly_summary.txt for a description of the columns in this file.\n")
        f.write("# assembly_accession\tbioproject\tbiosample\twgs_master\trefseq_category\ttaxid\tspecies_taxid\torganism_name\n")
        for g in (0, 2, 3):
            f.write("GCF_%09d.1\tPRJNA%d\tSAMN%d\t\treference genome\t%d\t%d\tSynthetic organism %d\n" % (100 + g, g, g, species[g], species[g], g))
        f.write("NC_%06d.1\tPRJNA1\tSAMN1\t\tna\t%d\t%d\tSynthetic organism 1\n" % (901, species[1], species[1]))
        f.write("GCF_%09d.2x\tPRJNA4\tSAMN4\t\tna\t%d\t%d\tSynthetic organism 4\n" % (104, species[4], species[4]))
    synth.write_taxonomy_files(os.path.join(d, "tax"), ids, par, rk)
    b1, o1, b2, o2, t = synth.make_read_pairs(genomes, 300, seed=5)
    synth.write_fastq_pair(os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq"), b1, o1, b2, o2, t)


def test_cli_assembly_summary_mapping(tmp_path):
    """`assembly_summary.txt` in a directory of sequence files maps sequence / file accessions to taxa before the headers
    are consulted (args_handling.cpp:120, taxonomy_io.cpp:190-311, mode_build.cpp:531-604): same shard files (taxon
    records) and same classification as the reference"""
    d = str(tmp_path)
    make_assembly_summary_files(d)
    run([EXE, "build", "gdb", "genomes", "-taxonomy", "tax", "-gpus", "2"], d)
    run([EXE, "query", "gdb", "r1.fq", "r2.fq"] + QFLAGS + ["-out", "res_gpu.txt"], d)
    mine = stable_lines(os.path.join(d, "res_gpu.txt"))
    golden = os.path.join(ROOT, "tests", "golden", "cli_np2_assembly_summary.txt")
    if os.path.exists(golden):
        assert mine == open(golden).read().splitlines()
    if not os.path.exists(REF):
        assert os.path.exists(golden), "neither oracle/_ref nor the golden output is available"
        return
    run([sys.executable, SHIM, "-np", "2", "--", REF, "build", "rdb", "genomes", "-taxonomy", "tax"], d)
    run([sys.executable, SHIM, "-np", "2", "--", REF, "query", "rdb", "r1.fq", "r2.fq"] + QFLAGS +
        ["-threads", "1", "-query-limit", "300", "-out", "res_ref.txt"], d)
    assert mine == stable_lines(os.path.join(d, "res_ref.txt"))
    for r in range(2):
        assert "shards equal" in run([EXE, "shards-equal", "gdb.db_%d" % r, "rdb.db_%d" % r], d)


MERGE_FLAGS = ["-lowest", "species", "-maxcand", "4", "-hitmin", "4", "-hitdiff", "80", "-tophits", "-queryids"]


def make_split_files(d):
    """the six genomes as two independent databases (A: genomes 0-2, B: genomes 3-5), one read set"""
    G, L = 6, 30000
    genomes = synth.make_genomes(G, L)
    ids, par, rk, species = synth.make_taxonomy(G)
    for name, gs in (("gA", range(0, 3)), ("gB", range(3, 6))):
        os.makedirs(os.path.join(d, name), exist_ok=True)
        for g in gs:
            synth.write_fasta(os.path.join(d, name, "g%05d.fna" % g), "SYN%06d.1 taxid=%d synthetic genome %d" % (g, species[g], g), genomes[g])
    synth.write_taxonomy_files(os.path.join(d, "tax"), ids, par, rk)
    b1, o1, b2, o2, t = synth.make_read_pairs(genomes, 300, seed=5)
    synth.write_fastq_pair(os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq"), b1, o1, b2, o2, t)


def test_cli_merge_mode(tmp_path):
    """`merge` (mode_merge.cpp, docs/merge.txt): the top-hit lists of two independent query runs (two databases holding
    half of the genomes each) are merged per read and classified; output identical to the reference's merge of the same
    result files"""
    d = str(tmp_path)
    make_split_files(d)
    for nm in ("A", "B"):
        run([EXE, "build", "db" + nm, "g" + nm, "-taxonomy", "tax", "-gpus", "2"], d)
        run([EXE, "query", "db" + nm, "r1.fq", "r2.fq", "-pairfiles"] + MERGE_FLAGS + ["-out", "res%s.txt" % nm], d)
    run([EXE, "merge", "resA.txt", "resB.txt", "-taxonomy", "tax"] + MERGE_FLAGS + ["-out", "merged_gpu.txt"], d)
    mine = stable_lines(os.path.join(d, "merged_gpu.txt"))
    golden = os.path.join(ROOT, "tests", "golden", "cli_merge_two_databases.txt")
    if os.path.exists(golden):
        assert mine == open(golden).read().splitlines()
    # merging finds the genomes of both halves: more reads classified at species level than either run alone
    def species_count(path):
        return [int(l[l.rindex("(") + 1:l.rindex(")")]) for l in open(path).read().splitlines() if l.startswith("#   species")][0]
    assert species_count(os.path.join(d, "merged_gpu.txt")) > max(species_count(os.path.join(d, "resA.txt")), species_count(os.path.join(d, "resB.txt")))
    if os.path.exists(REF):
        run([sys.executable, SHIM, "-np", "1", "--", REF, "merge", "resA.txt", "resB.txt", "-taxonomy", "tax"] + MERGE_FLAGS +
            ["-out", "merged_ref.txt"], d)
        assert mine == stable_lines(os.path.join(d, "merged_ref.txt"))
        # plain output (no ids, no top hits), reference defaults for everything but the rank
        run([EXE, "merge", "resA.txt", "resB.txt", "-taxonomy", "tax", "-out", "m2_gpu.txt"], d)
        run([sys.executable, SHIM, "-np", "1", "--", REF, "merge", "resA.txt", "resB.txt", "-taxonomy", "tax", "-out", "m2_ref.txt"], d)
        assert stable_lines(os.path.join(d, "m2_gpu.txt")) == stable_lines(os.path.join(d, "m2_ref.txt"))
