"""Host driver logic that needs no GPU (`metacache_b200 selftest-*` never touches CUDA):
  * shard files: write (all host threads) -> read (mapped, chunk-parallel parse) round trip, and the
    written bytes parse as the reference's format (sketch_database.h:959-999, hash_multimap.h:975-1029);
  * -abundances / -abundance-per: the tables of classification.cpp:362-428 / printing.cpp:474-516 from
    given per-taxon counts equal what the REFERENCE printed for those counts (tests/golden/cli_np2_*.txt,
    written by tests/golden/make_cli_golden.py from oracle/_ref)."""
import os
import subprocess

import pytest

from metacache_mpi_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "metacache-mpi_b200", "metacache_b200")


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        import __graft_entry__ as g
        g.build()
    assert os.path.exists(EXE), "host driver not built"
    return EXE


def test_shard_file_round_trip_and_format(exe, tmp_path):
    f = str(tmp_path / "x.db_0")
    r = subprocess.run([exe, "selftest-shard", f, "300000"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    assert r.returncode == 0 and b"selftest-shard ok" in r.stdout, r.stdout.decode()
    from test_gpu_cli_dropin import parse_shard
    head, taxa, ntgt, nv, buckets = parse_shard(f)                 # asserts exact EOF
    assert head[0] == 20181001 and head[1:7] == (4, 4, 4, 1, 8, 21)
    assert ntgt == 3 and len(taxa) == 4 and len(buckets) == 300000
    assert sum(len(t) for t, _ in buckets.values()) == nv
    assert max(len(t) for t, _ in buckets.values()) == 254
    # a few keys only: the tiny file goes through the same code with a single chunk
    r = subprocess.run([exe, "selftest-shard", f, "5"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=60)
    assert r.returncode == 0, r.stdout.decode()
    assert len(parse_shard(f)[4]) == 5


@pytest.mark.parametrize("golden,rank", [("cli_np2_abundances.txt", "species"), ("cli_np2_abundance_per_genus.txt", "genus"),
                                          ("cli_np2_abundance_two_levels.txt", "species")])
def test_abundance_tables_equal_reference_output(exe, tmp_path, golden, rank):
    lines = open(os.path.join(ROOT, "tests", "golden", golden)).read().splitlines()
    a = [i for i, l in enumerate(lines) if l.startswith("# query summary")][0]
    b = [i for i, l in enumerate(lines) if l.startswith("# queries:")][0]
    tables = lines[a:b]
    est = [i for i, l in enumerate(tables) if l.startswith("# estimated abundance")][0]
    counts = []
    for l in tables[1:est]:
        name, cnt, _ = l.split("\t|\t")
        counts.append("%s:%s" % (name.split(":taxon")[1], cnt))
    ids, par, rk, _ = synth.make_taxonomy(6)                        # the taxonomy of make_files() in test_gpu_cli_dropin
    synth.write_taxonomy_files(str(tmp_path / "tax"), ids, par, rk)
    r = subprocess.run([exe, "selftest-abundance", str(tmp_path / "tax"), rank] + counts + ["-total", "300"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=60)
    assert r.returncode == 0, r.stdout.decode()
    assert r.stdout.decode().splitlines() == tables


# (table text, expected taxon of the sequences 0..3); the sequences: file GCF_00000010<g>.1_x.fna, header
# ">NC_00090<g>.1 taxid=5<g> ..." -- the header taxid 5<g> shows that no table row named the sequence
TAXMAP_TABLES = {
    "ncbi": ("#   See README\n# assembly_accession\tbioproject\ttaxid\tspecies_taxid\n"
             "GCF_000000100.1\tP0\t70\t7\nNC_000901.1\tP1\t71\t7\nGCF_000000102.1x\tP2\t72\t7\n", [70, 71, 72, 53]),
    # (a key that is a prefix of the sequence name does not count, only the other way round: row 3 names nothing)
    "two_columns_no_header": ("GCF_000000100.1\t80\nGCF_000000101.1 \t81\n\nNC_000902\t82\n", [80, 81, 52, 53]),
    # no '#': the first word of the header is not a column, so "taxid" is column 0, which means "no taxid column":
    # the header line itself is read as a row, does not hold a number, and ends the table
    "header_without_hash": ("accession.version\ttaxid\nGCF_000000100.1\t90\n", [50, 51, 52, 53]),
    # "#key" is the dropped word: columns are counted from the second name on, rows from their first field
    "shifted_header": ("# a comment\n#key\tother\ttaxid\nGCF_000000100.1\t60\tzzz\nGCF_000000103.1\t63\tzzz\n", [60, 51, 52, 63]),
    # the first row of a key wins; a row without a number ends the table (later rows are not read)
    "duplicates_and_bad_row": ("# assembly_accession\ttaxid\nGCF_000000100.1\t40\nGCF_000000100.1\t41\nGCF_000000101.1\tnone\n"
                               "GCF_000000102.1\t42\n", [40, 51, 52, 53]),
}


def write_taxmap_world(d, table):
    gd = os.path.join(d, "genomes")
    os.makedirs(gd, exist_ok=True)
    genomes = synth.make_genomes(4, 2000)
    for g in range(4):
        synth.write_fasta(os.path.join(gd, "GCF_00000010%d.1_x.fna" % g), "NC_00090%d.1 taxid=5%d synthetic" % (g, g), genomes[g])
    with open(os.path.join(gd, "assembly_summary.txt"), "w") as f:
        f.write(table)


@pytest.mark.parametrize("case", sorted(TAXMAP_TABLES))
def test_sequence_to_taxon_tables(exe, tmp_path, case):
    """`assembly_summary.txt` next to the sequence files (taxonomy_io.cpp:190-311, mode_build.cpp:531-604): the parent taxon
    of every sequence equals the one the REFERENCE build stores in its shard file (when oracle/_ref is present) and the
    expectation written down here"""
    import sys
    from test_gpu_cli_dropin import REF, SHIM, parse_shard
    table, expect = TAXMAP_TABLES[case]
    d = str(tmp_path)
    write_taxmap_world(d, table)
    r = subprocess.run([exe, "selftest-taxmap", "genomes"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=60)
    assert r.returncode == 0, r.stdout.decode()
    rows = [l.split("\t") for l in r.stdout.decode().splitlines() if "\t" in l]
    assert [int(x[2]) for x in rows] == expect
    assert [x[1] for x in rows] == ["NC_00090%d.1" % g for g in range(4)]
    if os.path.exists(REF):
        r = subprocess.run([sys.executable, SHIM, "-np", "1", "--", REF, "build", "rdb", "genomes"], cwd=d,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
        assert r.returncode == 0, r.stdout.decode()[-2000:]
        taxa = parse_shard(os.path.join(d, "rdb.db_0"))[1]
        parents = {name.decode(): par for tid, par, rank, name, fn, idx, win in taxa if tid < 0}
        assert [parents["NC_00090%d.1" % g] for g in range(4)] == expect
