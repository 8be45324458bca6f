"""Live pin of the plain-C oracle against the reference compiled here (oracle/_ref);
skipped where oracle/_ref does not exist.  Larger / randomised than the golden fixtures."""
import numpy as np
import pytest

from metacache_mpi_b200 import synth
from oracle import pyoracle as po
from util import World


def test_l1_random(oracle, reference):
    rng = np.random.default_rng(3)
    for x in rng.integers(0, 2 ** 32, size=300, dtype=np.uint64):
        assert oracle.tmh(int(x)) == reference.tmh(int(x))
    for L in list(range(0, 300)) + [1000, 12345]:
        for wl, st in ((128, 113), (128, 128), (100, 30), (16, 1), (50, 60)):
            assert oracle.windows(L, wl, st) == reference.windows(L, wl, st)
    g = synth.genome(3, 40000)
    for k, s, wl, st in ((16, 16, 128, 113), (13, 20, 200, 150), (16, 48, 127, 100), (3, 5, 40, 7)):
        fo, co = oracle.sketch_sequences(*po.concat([g]), k=k, sketchlen=s, winlen=wl, stride=st)
        fr, cr = reference.sketch_sequence(g, k=k, sketchlen=s, winlen=wl, stride=st)
        assert fo.shape == fr.shape and (fo == fr).all() and (co == cr).all()


@pytest.mark.parametrize("P", [2, 8])
def test_pipeline_random(oracle, reference, P):
    w = World(G=10, L=25000, npairs=150, seed=P)
    rdbs = [reference.db_create() for _ in range(P)]
    for r in range(P):
        rdbs[r].set_taxonomy(w.ids, w.par, w.rk)
        for g in range(w.G):
            assert rdbs[r].add_target(w.genomes[g], "SYN%06d.1" % g, int(w.species[g]), P, r) == 1
    odbs = w.oracle_shards(oracle, P)
    for r in range(P):
        for x, y in zip(rdbs[r].dump(), odbs[r].export()):
            assert x.shape == y.shape and (x == y).all()
    for maxcand, lowest, hitmin, hitdiff in ((4, 4, 4, 0.8), (2, 0, 5, 1.0)):
        q = po.make_qopt(maxcand=maxcand, lowest=lowest, hitmin=hitmin, hitdiff=hitdiff)
        for i in range(len(w.o1) - 1):
            s1, s2 = w.b1[w.o1[i]:w.o1[i + 1]], w.b2[w.o2[i]:w.o2[i + 1]]
            lt, lh, rt, rh = [], [], [], []
            for r in range(P):
                t, h = odbs[r].query_local(w.tax, q, s1, s2)
                t2, h2 = rdbs[r].query_local(s1, s2, maxcand, lowest)
                assert (w.tabs.ids_of(t) == t2).all() and (h == h2).all()
                lt += list(t); lh += list(h); rt += list(t2); rh += list(h2)
            ot, oh, best, br = oracle.merge_classify(w.tax, q, lt, lh)
            ot2, oh2, best2, br2 = rdbs[0].merge_classify(rt, rh, maxcand, lowest, 19, hitmin, hitdiff)
            assert (w.tabs.ids_of(ot) == ot2).all() and (oh == oh2).all()
            assert (int(w.tabs.taxid[best]) if best >= 0 else 0) == best2 and br == br2
