"""one small invocation of the whole hot path on cuda:0, checked against the oracle"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from metacache_mpi_b200 import api, synth, taxonomy
    from oracle import pyoracle as po

    G, L, P = 4, 20000, 2
    genomes = synth.make_genomes(G, L)
    ids, par, rk, species = synth.make_taxonomy(G)
    tabs = taxonomy.TaxonomyTables(ids, par, rk, species)
    b1, o1, b2, o2, _ = synth.make_read_pairs(genomes, 200, seed=5)
    buf, offs = po.concat(genomes)

    ctx = api.Context(0)                      # raises when libmcgpu.so or the GPU is missing
    O = po.Oracle()
    # sketch
    fg, cg = ctx.sketch_windows(buf, offs)
    fo, co = O.sketch_sequences(buf, offs)
    assert (fg == fo).all() and (cg == co).all(), "sketch mismatch"
    # build
    gd = [api.Database(ctx, shard=r, nshards=P) for r in range(P)]
    od = [O.db_create(shard=r, nshards=P) for r in range(P)]
    for r in range(P):
        gd[r].add_targets(buf, offs, np.arange(G, dtype=np.uint32))
        gd[r].finalize()
        for g in range(G):
            od[r].add_target(genomes[g], g)
        od[r].finalize()
        for a, b in zip(gd[r].export(), od[r].export()):
            assert a.shape == b.shape and (a == b).all(), "shard mismatch"
    # query + merge + classify
    ctx.upload_taxonomy(tabs.lineage, tabs.rank, G)
    q = api.make_query_opts(maxcand=4, lowest=4, hitmin=4, hitdiff=0.8)
    oq = po.make_qopt(maxcand=4, lowest=4, hitmin=4, hitdiff=0.8)
    res = [d.query_batch(q, b1, o1, b2, o2) for d in gd]
    ct, ch, cn = (np.stack([x[i] for x in res]) for i in range(3))
    mt, mh, mn, best, brank, stats = ctx.merge_classify(q, ct, ch, cn)
    tax = po.TaxTables(tabs.lineage, tabs.rank, G)
    et, eh, en, eb, er, es = O.query_batch(od, tax, oq, b1, o1, b2, o2)
    assert (mt == et).all() and (mh == eh).all() and (mn == en).all(), "top-hit mismatch"
    assert (best == eb).all() and (brank == er).all() and (stats == es).all(), "classification mismatch"
    # the one-call form (mcg_query_batch_sharded; a single shard needs no communicator): query + merge + decision
    g1 = api.Database(ctx)
    g1.add_targets(buf, offs, np.arange(G, dtype=np.uint32))
    g1.finalize()
    o1db = O.db_create()
    for g in range(G):
        o1db.add_target(genomes[g], g)
    o1db.finalize()
    st, sh, sn, sb, sr, ss = g1.query_batch_sharded(None, q, b1, o1, b2, o2)
    et1, eh1, en1, eb1, er1, es1 = O.query_batch([o1db], tax, oq, b1, o1, b2, o2)
    assert (st == et1).all() and (sh == eh1).all() and (sn == en1).all() and (sb == eb1).all() and (sr == er1).all(), "sharded call mismatch"
    assert (ss == es1).all(), "sharded call statistics mismatch"
    print("smoke ok: %d windows, %d reads, %d classified, %d kernel launches"
          % (len(fg), len(o1) - 1, int(stats[20]), ctx.launches()))
    ctx.close()


if __name__ == "__main__":
    run()
