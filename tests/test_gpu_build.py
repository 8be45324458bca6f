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


def test_medium_size_database_against_numpy(ctx):
    """40 Mbp + a low-complexity target: thousands of sort / run / table tiles, a run of one feature far longer
    than a tile (poly-A, cut to maxlocs), N runs.  Expected shard = stable sort of the (already verified) GPU
    sketches by feature, first maxlocs locations per feature; expected matches of a read = the lists of its
    features, merged by (target, window) (sketch_database.h:1079-1097, hash_multimap.h:706-713,
    sketch_database.h:804-823).  Checks the radix sort, the run / location emission, the bucket-centric
    table and the probe against plain numpy at a size where tile boundaries are everywhere."""
    from metacache_mpi_b200 import synth
    G, L = 8, 5_000_000
    genomes = [synth.genome(g, L) for g in range(G)]
    genomes.append(np.frombuffer(b"A" * 700_000, np.uint8).copy())                  # one feature, ~6200 windows
    genomes.append(np.frombuffer((b"ACGT" * 40 + b"N" * 60) * 2000, np.uint8).copy())   # few features, many locations
    buf = np.concatenate(genomes)
    offs = np.concatenate([[0], np.cumsum([len(g) for g in genomes])]).astype(np.int64)
    ids = np.arange(len(genomes), dtype=np.uint32)
    feat, cnt = ctx.sketch_windows(buf, offs)
    db = api.Database(ctx)
    db.add_targets(buf, offs, ids)
    db.finalize()
    keys, sizes, tg, wn = db.export()
    # expectation
    cfg = api.SketchCfg(16, 16, 128, 113)
    nwin_per = [ctx.count_windows(cfg, np.array([0, len(g)], np.int64)) for g in genomes]
    assert sum(nwin_per) == feat.shape[0]
    wtgt = np.repeat(np.arange(len(genomes), dtype=np.uint32), nwin_per)
    wwin = np.concatenate([np.arange(n, dtype=np.uint32) for n in nwin_per])
    valid = (np.arange(16)[None, :] < cnt[:, None]).ravel()
    f = feat.ravel()[valid]
    ft = np.repeat(wtgt, 16)[valid]
    fw = np.repeat(wwin, 16)[valid]
    order = np.argsort(f, kind="stable")
    f, ft, fw = f[order], ft[order], fw[order]
    head = np.concatenate([[True], f[1:] != f[:-1]])
    start = np.flatnonzero(head)
    runlen = np.diff(np.concatenate([start, [len(f)]]))
    pos = np.arange(len(f)) - np.repeat(start, runlen)
    keep = pos < 254
    assert (keys == f[head]).all()
    assert (sizes == np.minimum(runlen, 254)).all()
    assert (tg == ft[keep]).all() and (wn == fw[keep]).all()
    assert runlen.max() > 4096 and (sizes == 254).any()
    # probes: reads from the genomes, the poly-A target and absent features
    rng = np.random.default_rng(5)
    reads = []
    for _ in range(3000):
        g = int(rng.integers(0, len(genomes)))
        p = int(rng.integers(0, len(genomes[g]) - 150))
        reads.append(genomes[g][p:p + 150].copy())
    reads += [rng.choice(np.frombuffer(b"ACGT", np.uint8), 150) for _ in range(200)]
    b1 = np.concatenate(reads)
    o1 = (np.arange(len(reads) + 1) * 150).astype(np.int64)
    from metacache_mpi_b200 import taxonomy
    tabs = taxonomy.TaxonomyTables(*synth.make_taxonomy(len(genomes)))
    ctx.upload_taxonomy(tabs.lineage, tabs.rank, len(genomes))
    q = api.make_query_opts()
    mcnt, mtg, mwn = db.query_matches(q, b1, o1)
    rf, rc = ctx.sketch_windows(b1, o1)
    koffs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    p0 = 0
    for i in range(len(reads)):
        ff = np.concatenate([rf[2 * i][:rc[2 * i]], rf[2 * i + 1][:rc[2 * i + 1]]])
        j = np.searchsorted(keys, ff)
        hit = (j < len(keys)) & (keys[np.minimum(j, len(keys) - 1)] == ff)
        exp = np.concatenate([np.arange(koffs[x], koffs[x + 1]) for x in j[hit]]) if hit.any() else np.zeros(0, np.int64)
        e = np.sort((tg[exp].astype(np.uint64) << np.uint64(32)) | wn[exp].astype(np.uint64), kind="stable")
        got = (mtg[p0:p0 + mcnt[i]].astype(np.uint64) << np.uint64(32)) | mwn[p0:p0 + mcnt[i]].astype(np.uint64)
        assert mcnt[i] == len(e) and (got == e).all(), i
        p0 += int(mcnt[i])
    db.close()
