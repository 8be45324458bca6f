"""K2 parity: shard contents built on the GPU == oracle (order-insensitive per key is not even
needed: lists are compared in (tgt,win) order, which both sides produce)"""
import numpy as np
import pytest

from metacache_mpi_b200 import api
from util import World

pytestmark = pytest.mark.gpu


def _same(a, b):
    for x, y in zip(a, b):
        assert x.shape == y.shape, (x.shape, y.shape)
        assert (x == y).all()


@pytest.mark.parametrize("P", [1, 2, 4])
def test_shards_equal_oracle(ctx, oracle, P):
    w = World(G=8, L=30000)
    od = w.oracle_shards(oracle, P)
    gd = w.gpu_shards(api, ctx, P)
    for r in range(P):
        _same(gd[r].export(), od[r].export())
        nk, ne, nl = gd[r].counts()
        assert (nk, ne, nl) == od[r].counts()


@pytest.mark.parametrize("maxlocs", [1, 3, 254, 300])
def test_location_list_capping(ctx, oracle, maxlocs):
    # many identical genomes: every feature has G locations, so the cap (and its uint8
    # narrowing, 300 -> 44) decides which survive: the first maxlocs in (tgt,win) order
    w = World(G=6, L=12000, tie_block=False)
    w.genomes = [w.genomes[0].copy() for _ in range(6)]
    from oracle import pyoracle as po
    w.buf, w.offs = po.concat(w.genomes)
    od = w.oracle_shards(oracle, 2, maxlocs=maxlocs)
    gd = w.gpu_shards(api, ctx, 2, maxlocs=maxlocs)
    for r in range(2):
        _same(gd[r].export(), od[r].export())


def test_remove_overpopulated(ctx, oracle):
    w = World(G=8, L=30000)
    for P in (1, 2, 4):
        od = w.oracle_shards(oracle, P, maxlocs=3)
        gd = w.gpu_shards(api, ctx, P, maxlocs=3)
        assert api.remove_overpopulated(gd) == oracle.remove_overpopulated(od)
        for r in range(P):
            _same(gd[r].export(), od[r].export())
            assert gd[r].counts() == od[r].counts()


def test_incremental_add_and_import_roundtrip(ctx, oracle):
    w = World(G=8, L=20000)
    od = w.oracle_shards(oracle, 2)
    for r in range(2):
        db = api.Database(ctx, shard=r, nshards=2)
        # targets arrive in three calls
        for lo, hi in ((0, 3), (3, 4), (4, 8)):
            from oracle import pyoracle as po
            buf, offs = po.concat(w.genomes[lo:hi])
            wins = db.add_targets(buf, offs, np.arange(lo, hi, dtype=np.uint32))
            for g in range(lo, hi):
                want = oracle.L.mco_num_windows(len(w.genomes[g]), 128, 113) if g % 2 == r else 0
                assert wins[g - lo] == want
        db.finalize()
        exp = db.export()
        _same(exp, od[r].export())
        # export -> import (shuffled bucket order, like a reference-written file) -> export
        keys, sizes, tg, wn = exp
        perm = np.random.default_rng(1).permutation(len(keys))
        starts = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        idx = np.concatenate([np.arange(starts[i], starts[i] + sizes[i]) for i in perm])
        db2 = api.Database(ctx, shard=r, nshards=2)
        db2.import_(keys[perm], sizes[perm], tg[idx], wn[idx])
        _same(db2.export(), exp)


def test_empty_and_tiny_targets(ctx, oracle):
    from oracle import pyoracle as po
    seqs = [b"ACGT", b"", b"ACGTACGTACGTACGTA", b"N" * 200, b"ACGT" * 64]
    db = api.Database(ctx)
    od = oracle.db_create()
    buf, offs = po.concat(seqs)
    db.add_targets(buf, offs, np.arange(len(seqs), dtype=np.uint32))
    for i, s in enumerate(seqs):
        od.add_target(s, i)
    db.finalize()
    od.finalize()
    _same(db.export(), od.export())
