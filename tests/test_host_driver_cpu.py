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
