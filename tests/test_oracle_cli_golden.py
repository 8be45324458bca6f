"""The oracle's whole query path (per-shard lists, cross-rank merge, decision, statistics) against what the
REFERENCE BINARY printed for the same files (tests/golden/cli_*.txt, written by tests/golden/make_cli_golden.py
from oracle/_ref under the MPI shim).  Pins in particular the reference-compatible merge (SURVEY 3.4 quirk Q1:
uint32 wire format, reduction tree of querying.h:892-1066) at np = 2 and np = 4.  No GPU, no /root/reference."""
import os

import numpy as np
import pytest

from metacache_mpi_b200 import synth, taxonomy
from oracle import pyoracle as po

import clifmt

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# golden file, ranks, genomes without taxid, oracle options
CASES = [
    ("cli_np2_species_maxcand4.txt", 2, (), dict(maxcand=4, lowest=4, hitmin=4, hitdiff=0.8, reference_compat=0)),
    ("cli_np2_species_maxcand4.txt", 2, (), dict(maxcand=4, lowest=4, hitmin=4, hitdiff=0.8, reference_compat=1)),
    ("cli_np2_default_flags.txt", 2, (), dict(maxcand=2, lowest=0, hitmin=5, hitdiff=1.0, reference_compat=1)),
    ("cli_np4_default_flags.txt", 4, (), dict(maxcand=2, lowest=0, hitmin=5, hitdiff=1.0, reference_compat=1)),
    ("cli_np2_mixed_lowest_species.txt", 2, (1, 4), dict(maxcand=4, lowest=4, hitmin=4, hitdiff=0.8, reference_compat=1)),
    ("cli_np4_mixed_lowest_species.txt", 4, (1, 4), dict(maxcand=4, lowest=4, hitmin=4, hitdiff=0.8, reference_compat=1)),
]


def cli_world(no_taxid=()):
    """the files of test_gpu_cli_dropin.make_files() as arrays"""
    G, L = 6, 30000
    genomes = synth.make_genomes(G, L)
    ids, par, rk, species = synth.make_taxonomy(G)
    species = species.copy()
    for g in no_taxid:
        species[g] = 0
    tabs = taxonomy.TaxonomyTables(ids, par, rk, species)
    b1, o1, b2, o2, truth = synth.make_read_pairs(genomes, 300, seed=5)
    return genomes, tabs, (b1, o1, b2, o2), truth


@pytest.mark.parametrize("golden,P,no_taxid,opts", CASES)
def test_oracle_equals_reference_binary_output(oracle, golden, P, no_taxid, opts):
    genomes, tabs, reads, truth = cli_world(no_taxid)
    dbs = [oracle.db_create(shard=r, nshards=P) for r in range(P)]
    for r in range(P):
        for g, s in enumerate(genomes):
            dbs[r].add_target(s, g)
        dbs[r].finalize()
    tax = po.TaxTables(tabs.lineage, tabs.rank, len(genomes))
    q = po.make_qopt(**opts)
    ot, oh, on, best, brank, st = oracle.query_batch(dbs, tax, q, *reads)
    mine = clifmt.format_lines(tabs, truth, ot, oh, on, best, opts["lowest"])
    want = clifmt.golden_read_lines(os.path.join(GOLD, golden))
    assert len(want) == 300
    assert mine == want
    assert clifmt.stats_summary(st, 300) == clifmt.golden_summary(os.path.join(GOLD, golden))


def test_compat_changes_the_result():
    """the two merges differ where the reference loses candidates: the goldens above are not vacuous"""
    O = po.Oracle()
    genomes, tabs, reads, truth = cli_world()
    dbs = [O.db_create(shard=r, nshards=2) for r in range(2)]
    for r in range(2):
        for g, s in enumerate(genomes):
            dbs[r].add_target(s, g)
        dbs[r].finalize()
    tax = po.TaxTables(tabs.lineage, tabs.rank, len(genomes))
    a = O.query_batch(dbs, tax, po.make_qopt(maxcand=2, lowest=0, hitmin=5, reference_compat=0), *reads)
    b = O.query_batch(dbs, tax, po.make_qopt(maxcand=2, lowest=0, hitmin=5, reference_compat=1), *reads)
    # reads of the odd genomes (shard 1) are assigned to their even sibling by the reference
    odd = truth % 2 == 1
    assert (a[3][odd] != b[3][odd]).mean() > 0.9 and (a[3][~odd] == b[3][~odd]).mean() > 0.9
