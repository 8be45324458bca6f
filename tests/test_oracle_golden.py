"""The plain-C oracle against the committed golden vectors, which were produced by the REAL
reference (tests/golden/make_golden.py via oracle/_ref).  Runs everywhere (no GPU, no
/root/reference)."""
import os
import sys

import numpy as np
import pytest

from oracle import pyoracle as po

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden as mg  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))


def test_hash_revcomp_canonical(oracle):
    for x, y in zip(G["tmh_in"], G["tmh_out"]):
        assert oracle.tmh(int(x)) == int(y)
    # SURVEY.md 8(c) known answers
    assert [oracle.tmh(x) for x in (0, 1, 0xFFFFFFFF, 0x12345678, 0xDEADBEEF)] == \
        [0, 0x31251ba7, 0x2028884f, 0x055ccace, 0x5353e2e9]
    for x, k, r, c in zip(G["rc_in"], G["rc_k"], G["rc_out"], G["canon_out"]):
        assert oracle.revcomp(int(x), int(k)) == int(r)
        assert oracle.canonical(int(x), int(k)) == int(c)
    assert oracle.revcomp(0x1b1b1b1b, 16) == 0x1b1b1b1b and oracle.revcomp(1, 16) == 0xbfffffff


def test_kmers_and_ambiguity(oracle):
    km, am = oracle.kmers("ACGTNacgtRYACGTTGCAACGT", 4)
    assert (km == G["kmers4"]).all() and (am == G["ambig4"]).all()
    km, am = oracle.kmers("ACGTTGCA", 4)
    assert list(km) == [0x1b, 0x6f, 0xbe, 0xf9, 0xe4] and not am.any()


def test_windows(oracle):
    for row in G["windows"]:
        L, w, s, n = (int(x) for x in row[:4])
        ws = oracle.windows(L, w, s)
        assert len(ws) == n
        first = [x for be in ws[:3] for x in be] + [0] * (6 - 2 * min(len(ws), 3))
        assert first == [int(x) for x in row[4:10]] and list(ws[-1]) == [int(x) for x in row[10:12]]
    assert oracle.windows(300, 128, 113) == [(0, 128), (113, 241), (226, 300)]
    assert oracle.windows(241, 128, 113) == [(0, 128), (113, 241), (226, 241)]


def test_sketches(oracle):
    g = G["sk_genome"]
    for i, (k, s, w, st) in enumerate(mg.SKETCH_PARAMS):
        f, c = oracle.sketch_sequences(*po.concat([g]), k=k, sketchlen=s, winlen=w, stride=st)
        assert f.shape == G["sk_feat_%d" % i].shape
        assert (f == G["sk_feat_%d" % i]).all() and (c == G["sk_cnt_%d" % i]).all()
    i = 0
    while "kat_seq_%d" % i in G:
        f, c = oracle.sketch_sequences(*po.concat([G["kat_seq_%d" % i]]))
        assert (f == G["kat_feat_%d" % i]).all() and (c == G["kat_cnt_%d" % i]).all()
        i += 1
    assert i >= 9
    # strand symmetry and case-insensitivity (entries 0, 3, 4 are W, lower(W), revcomp(W))
    assert (G["kat_feat_0"] == G["kat_feat_3"]).all() and (G["kat_feat_0"] == G["kat_feat_4"]).all()


@pytest.mark.parametrize("P", [1, 2, 4])
def test_database_and_query(oracle, P):
    w = mg.golden_world()
    for maxlocs, remove in ((254, False), (3, True)):
        tag = "P%d_m%d" % (P, maxlocs)
        dbs = w.oracle_shards(oracle, P, maxlocs=maxlocs)
        if remove:
            oracle.remove_overpopulated(dbs)
        for r in range(P):
            k, s, t, wn = dbs[r].export()
            assert (k == G["%s_r%d_keys" % (tag, r)]).all() and (s == G["%s_r%d_sizes" % (tag, r)]).all()
            assert (t == G["%s_r%d_tgts" % (tag, r)]).all() and (wn == G["%s_r%d_wins" % (tag, r)]).all()
        if maxlocs != 254:
            continue
        n = len(w.o1) - 1
        pos = 0
        q0 = po.make_qopt()
        for i in range(n):
            a, b = dbs[0].matches(q0, w.b1[w.o1[i]:w.o1[i + 1]], w.b2[w.o2[i]:w.o2[i + 1]])
            c = int(G[tag + "_match_cnt"][i])
            assert len(a) == c
            assert (a == G[tag + "_match_tgt"][pos:pos + c]).all() and (b == G[tag + "_match_win"][pos:pos + c]).all()
            pos += c
        for ci, (maxcand, lowest, hitmin, hitdiff) in enumerate(mg.QUERY_CONFIGS):
            q = "%s_q%d" % (tag, ci)
            oq = po.make_qopt(maxcand=maxcand, lowest=lowest, hitmin=hitmin, hitdiff=hitdiff)
            for r in range(P):
                for i in range(n):
                    t, h = dbs[r].query_local(w.tax, oq, w.b1[w.o1[i]:w.o1[i + 1]], w.b2[w.o2[i]:w.o2[i + 1]])
                    c = int(G[q + "_ln"][r, i])
                    assert len(t) == c
                    assert (w.tabs.ids_of(t) == G[q + "_ltax"][r, i, :c]).all() and (h == G[q + "_lhits"][r, i, :c]).all()
            mt, mh, mn, best, brank, stats = oracle.query_batch(dbs, w.tax, oq, w.b1, w.o1, w.b2, w.o2)
            assert (mn == G[q + "_mn"]).all()
            for i in range(n):
                c = int(mn[i])
                assert (w.tabs.ids_of(mt[i, :c]) == G[q + "_mtax"][i, :c]).all() and (mh[i, :c] == G[q + "_mhits"][i, :c]).all()
            bid = np.where(best >= 0, w.tabs.taxid[np.maximum(best, 0)], 0)
            assert (bid == G[q + "_best"]).all() and (brank == G[q + "_brank"]).all()
            assert (stats == G[q + "_stats"]).all()


def test_lineage_tables():
    w = mg.golden_world()
    assert (w.tabs.taxid == G["lineage_ids"]).all()
    assert (w.tabs.ids_of(w.tabs.lineage) == G["lineage"]).all()


def test_statistics(oracle):
    ranks = np.array([0, 4, 4, 6, 21, 20, 10, 21, 19], np.int32)
    st = oracle.statistics(ranks)
    assert st[21] == 2 and st[20] == 7 and st[0] == 1 and st[4] == 3 and st[6] == 4 and st[10] == 5 and st[19] == 6
