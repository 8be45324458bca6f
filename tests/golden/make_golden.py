#!/usr/bin/env python3
"""tests/golden/make_golden.py -- generates tests/golden/*.npz from the REAL reference
(oracle/_ref/libref_l1.so = unmodified /root/reference sources behind ref_harness.cpp).
Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The fixtures pin the plain-C oracle (and through it the CUDA path) to the reference's own
outputs on machines that have neither /root/reference nor oracle/_ref.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from __graft_entry__ import load_package  # noqa: E402

load_package()
from metacache_mpi_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from util import World, lcg_window  # noqa: E402

SKETCH_PARAMS = [(16, 16, 128, 113), (12, 8, 64, 50), (16, 32, 256, 241), (5, 16, 128, 113), (16, 16, 20, 5),
                 (1, 4, 10, 10), (16, 24, 127, 100), (8, 16, 128, 200)]
QUERY_CONFIGS = [(4, 4, 4, 0.8), (2, 0, 5, 1.0), (3, 6, 2, 0.5), (16, 10, 1, 0.1)]


def golden_world():
    return World(G=6, L=113 * 90 + 37, npairs=60, seed=11)


def main():
    R = po.Reference()
    out = {}
    # ---- L1 known answers
    xs = np.array([0, 1, 0xFFFFFFFF, 0x12345678, 0xDEADBEEF, 0x1b1b1b1b, 0x80000000, 12345], np.uint64)
    out["tmh_in"] = xs
    out["tmh_out"] = np.array([R.tmh(int(x)) for x in xs], np.uint64)
    rng = np.random.default_rng(1)
    rc_in, rc_k, rc_out, cn_out = [], [], [], []
    for k in range(1, 17):
        for x in rng.integers(0, 2 ** 32, size=12, dtype=np.uint64):
            x = int(x) & ((1 << (2 * k)) - 1)
            rc_in.append(x); rc_k.append(k); rc_out.append(R.revcomp(x, k)); cn_out.append(R.canonical(x, k))
    out["rc_in"], out["rc_k"] = np.array(rc_in, np.uint64), np.array(rc_k, np.int32)
    out["rc_out"], out["canon_out"] = np.array(rc_out, np.uint64), np.array(cn_out, np.uint64)
    W = lcg_window()
    out["W"] = np.frombuffer(W.encode(), np.uint8)
    km, am = R.kmers("ACGTNacgtRYACGTTGCAACGT", 4)
    out["kmers4"], out["ambig4"] = km, am
    wl = []
    for L in (0, 1, 15, 16, 127, 128, 129, 150, 240, 241, 242, 300, 1000):
        for (w, s) in ((128, 113), (128, 128), (100, 30), (16, 1), (50, 60)):
            ws = R.windows(L, w, s)
            wl.append([L, w, s, len(ws)] + [x for be in ws[:3] for x in be] + [0] * (6 - 2 * min(len(ws), 3)) + list(ws[-1]))
    out["windows"] = np.array(wl, np.int64)
    Wn = list(W)
    Wn[40], Wn[41], Wn[100] = "N", "n", "R"
    seqs = [W, "".join(Wn), W[:20], W.lower(), synth.revcomp_ascii(np.frombuffer(W.encode(), np.uint8)).tobytes().decode(),
            "A" * 300, "AC" * 200, "N" * 140, "ACGT" * 3]
    g = synth.genome(1, 3000)
    for i, (k, s, w, st) in enumerate(SKETCH_PARAMS):
        f, c = R.sketch_sequence(g, k=k, sketchlen=s, winlen=w, stride=st)
        out["sk_feat_%d" % i], out["sk_cnt_%d" % i] = f, c
    out["sk_genome"] = g
    for i, sq in enumerate(seqs):
        f, c = R.sketch_sequence(sq, 16, 16, 128, 113)
        out["kat_seq_%d" % i] = np.frombuffer(sq.encode() if isinstance(sq, str) else sq, np.uint8)
        out["kat_feat_%d" % i], out["kat_cnt_%d" % i] = f, c

    # ---- L2..L4 on a small world
    w = golden_world()
    for P in (1, 2, 4):
        for (maxlocs, remove) in ((254, False), (3, True)):
            tag = "P%d_m%d" % (P, maxlocs)
            rdbs = [R.db_create(maxlocs=maxlocs) for _ in range(P)]
            for r in range(P):
                rdbs[r].set_taxonomy(w.ids, w.par, w.rk)
                for gi in range(w.G):
                    assert rdbs[r].add_target(w.genomes[gi], "SYN%06d.1" % gi, int(w.species[gi]), P, r) == 1
            if remove:
                tot = {}
                for r in range(P):
                    k, s, _, _ = rdbs[r].dump()
                    for kk, ss in zip(k.tolist(), s.tolist()):
                        tot[kk] = tot.get(kk, 0) + ss
                lim = rdbs[0].maxlocs() - 1
                rem = np.array(sorted(k for k, v in tot.items() if v > lim), np.uint32)
                out[tag + "_removed"] = rem
                for r in range(P):
                    rdbs[r].remove_keys(rem)
            for r in range(P):
                k, s, t, wn = rdbs[r].dump()
                out["%s_r%d_keys" % (tag, r)], out["%s_r%d_sizes" % (tag, r)] = k, s
                out["%s_r%d_tgts" % (tag, r)], out["%s_r%d_wins" % (tag, r)] = t, wn
            if maxlocs != 254:
                continue
            n = len(w.o1) - 1
            # sorted match lists of shard 0
            cnt, tg, wn = [], [], []
            for i in range(n):
                a, b = rdbs[0].matches(w.b1[w.o1[i]:w.o1[i + 1]], w.b2[w.o2[i]:w.o2[i + 1]])
                cnt.append(len(a)); tg.append(a); wn.append(b)
            out[tag + "_match_cnt"] = np.array(cnt, np.int64)
            out[tag + "_match_tgt"], out[tag + "_match_win"] = np.concatenate(tg), np.concatenate(wn)
            for ci, (maxcand, lowest, hitmin, hitdiff) in enumerate(QUERY_CONFIGS):
                lt = np.zeros((P, n, maxcand), np.int64); lh = np.zeros((P, n, maxcand), np.uint64); ln = np.zeros((P, n), np.int32)
                mt = np.zeros((n, maxcand), np.int64); mh = np.zeros((n, maxcand), np.uint64); mn = np.zeros(n, np.int32)
                best = np.zeros(n, np.int64); brank = np.zeros(n, np.int32)
                for i in range(n):
                    s1, s2 = w.b1[w.o1[i]:w.o1[i + 1]], w.b2[w.o2[i]:w.o2[i + 1]]
                    at, ah = [], []
                    for r in range(P):
                        t, h = rdbs[r].query_local(s1, s2, maxcand, lowest)
                        ln[r, i] = len(t); lt[r, i, :len(t)] = t; lh[r, i, :len(t)] = h
                        at += list(t); ah += list(h)
                    t, h, b, br = rdbs[0].merge_classify(at, ah, maxcand, lowest, 19, hitmin, hitdiff)
                    mn[i] = len(t); mt[i, :len(t)] = t; mh[i, :len(t)] = h; best[i] = b; brank[i] = br
                q = "%s_q%d" % (tag, ci)
                out[q + "_ltax"], out[q + "_lhits"], out[q + "_ln"] = lt, lh, ln
                out[q + "_mtax"], out[q + "_mhits"], out[q + "_mn"] = mt, mh, mn
                out[q + "_best"], out[q + "_brank"] = best, brank
                out[q + "_stats"] = R.statistics(brank)
    # lineages of every taxon (targets + higher taxa) from the reference's ranked_lineages_cache
    rdb = R.db_create()
    rdb.set_taxonomy(w.ids, w.par, w.rk)
    for gi in range(w.G):
        rdb.add_target(w.genomes[gi], "SYN%06d.1" % gi, int(w.species[gi]), 1, 0)
    out["lineage_ids"] = w.tabs.taxid
    out["lineage"] = np.stack([rdb.lineage(int(t)) for t in w.tabs.taxid])
    path = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
