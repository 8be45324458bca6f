"""K3..K6 parity: sorted match lists, per-shard candidate lists, merged lists, classification
and statistics from the CUDA path == oracle, bit-exact, at P = 1, 2, 4 shards"""
import numpy as np
import pytest

from metacache_mpi_b200 import api, synth
from oracle import pyoracle as po
from util import World

pytestmark = pytest.mark.gpu

import os as _os
ROOT_DIR = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))

CONFIGS = [  # maxcand, lowest, hitmin, hitdiff
    (4, 4, 4, 0.8),      # the authors' script settings: -lowest species -maxcand 4 -hitmin 4 -hitdiff 80
    (2, 0, 5, 1.0),      # reference defaults
    (3, 6, 2, 0.5),      # -lowest genus
    (16, 10, 1, 0.1),    # wide list, -lowest family
]


def _gpu_full(ctx, dbs, q, b1, o1, b2, o2):
    ct, ch, cn = [], [], []
    for d in dbs:
        t, h, n = d.query_batch(q, b1, o1, b2, o2)
        ct.append(t); ch.append(h); cn.append(n)
    ct, ch, cn = np.stack(ct), np.stack(ch), np.stack(cn)
    return (ct, ch, cn) + ctx.merge_classify(q, ct, ch, cn)


def _oracle_local(dbs, tax, oq, b1, o1, b2, o2, K):
    n = len(o1) - 1
    ct = np.full((len(dbs), n, K), -1, np.int32)
    ch = np.zeros((len(dbs), n, K), np.uint32)
    cn = np.zeros((len(dbs), n), np.int32)
    for r, d in enumerate(dbs):
        for i in range(n):
            s2 = b2[o2[i]:o2[i + 1]] if b2 is not None else b""
            t, h = d.query_local(tax, oq, b1[o1[i]:o1[i + 1]], s2)
            cn[r, i] = len(t)
            ct[r, i, :len(t)] = t
            ch[r, i, :len(t)] = h
    return ct, ch, cn


@pytest.mark.parametrize("P", [1, 2, 4])
def test_pairs_all_stages(ctx, oracle, P):
    w = World(G=8, L=30000, npairs=400)
    ctx.upload_taxonomy(w.tabs.lineage, w.tabs.rank, w.G)
    od = w.oracle_shards(oracle, P)
    gd = w.gpu_shards(api, ctx, P)
    # sorted match lists (querying.h:798-811)
    q0 = api.make_query_opts()
    oq0 = po.make_qopt()
    for r in range(P):
        cnt, tg, wn = gd[r].query_matches(q0, w.b1, w.o1, w.b2, w.o2)
        pos = 0
        for i in range(len(w.o1) - 1):
            mt, mw = od[r].matches(oq0, w.b1[w.o1[i]:w.o1[i + 1]], w.b2[w.o2[i]:w.o2[i + 1]])
            assert cnt[i] == len(mt)
            assert (tg[pos:pos + cnt[i]] == mt).all() and (wn[pos:pos + cnt[i]] == mw).all()
            pos += int(cnt[i])
    for maxcand, lowest, hitmin, hitdiff in CONFIGS:
        q = api.make_query_opts(maxcand=maxcand, lowest=lowest, hitmin=hitmin, hitdiff=hitdiff)
        oq = po.make_qopt(maxcand=maxcand, lowest=lowest, hitmin=hitmin, hitdiff=hitdiff)
        ct, ch, cn, mt, mh, mn, best, brank, stats = _gpu_full(ctx, gd, q, w.b1, w.o1, w.b2, w.o2)
        ot, oh, on = _oracle_local(od, w.tax, oq, w.b1, w.o1, w.b2, w.o2, maxcand)
        assert (cn == on).all() and (ct == ot).all() and (ch == oh).all()
        et, eh, en, eb, er, es = oracle.query_batch(od, w.tax, oq, w.b1, w.o1, w.b2, w.o2)
        assert (mn == en).all() and (mt == et).all() and (mh == eh).all()
        assert (best == eb).all() and (brank == er).all()
        assert (stats == es).all()
        assert stats[20] + stats[21] == len(w.o1) - 1


def test_many_targets_per_read(ctx, oracle):
    """a block shared by 38 of 40 genomes: location lists of every gather class (inline, per-lane,
    whole-warp), reads with 65..128 matches (register sort of 128), reads with more than 32
    candidates (chunked, filtered top-list insertion) and reads left over to the staged path"""
    w = World(G=40, L=12000, npairs=1500, seed=11)
    ctx.upload_taxonomy(w.tabs.lineage, w.tabs.rank, w.G)
    for P in (1, 2):
        od = w.oracle_shards(oracle, P)
        gd = w.gpu_shards(api, ctx, P)
        for maxcand, lowest, hitmin, hitdiff in CONFIGS:
            q = api.make_query_opts(maxcand=maxcand, lowest=lowest, hitmin=hitmin, hitdiff=hitdiff)
            oq = po.make_qopt(maxcand=maxcand, lowest=lowest, hitmin=hitmin, hitdiff=hitdiff)
            ct, ch, cn, mt, mh, mn, best, brank, stats = _gpu_full(ctx, gd, q, w.b1, w.o1, w.b2, w.o2)
            ot, oh, on = _oracle_local(od, w.tax, oq, w.b1, w.o1, w.b2, w.o2, maxcand)
            assert (cn == on).all() and (ct == ot).all() and (ch == oh).all()
            et, eh, en, eb, er, es = oracle.query_batch(od, w.tax, oq, w.b1, w.o1, w.b2, w.o2)
            assert (mn == en).all() and (mt == et).all() and (mh == eh).all()
            assert (best == eb).all() and (brank == er).all() and (stats == es).all()
        q0 = api.make_query_opts()
        cnt, _, _ = gd[0].query_matches(q0, w.b1, w.o1, w.b2, w.o2)
        assert ((cnt > 64) & (cnt <= 128)).any() and (cnt > 128).any() and (cnt <= 64).any()


def test_single_end_and_ragged_reads(ctx, oracle):
    w = World(G=6, L=20000, npairs=10)
    ctx.upload_taxonomy(w.tabs.lineage, w.tabs.rank, w.G)
    od = w.oracle_shards(oracle, 2)
    gd = w.gpu_shards(api, ctx, 2)
    rng = np.random.default_rng(9)
    reads = [b"", b"ACGT", b"N" * 150]
    for n in (16, 50, 128, 129, 241, 500, 2000):
        g = w.genomes[int(rng.integers(0, w.G))]
        p = int(rng.integers(0, len(g) - n))
        reads.append(g[p:p + n].copy())
    b1, o1 = po.concat(reads)
    q = api.make_query_opts(maxcand=4, lowest=4, hitmin=2, hitdiff=0.8)
    oq = po.make_qopt(maxcand=4, lowest=4, hitmin=2, hitdiff=0.8)
    ct, ch, cn, mt, mh, mn, best, brank, stats = _gpu_full(ctx, gd, q, b1, o1, None, None)
    et, eh, en, eb, er, es = oracle.query_batch(od, w.tax, oq, b1, o1, None, None)
    assert (mn == en).all() and (mt == et).all() and (mh == eh).all() and (best == eb).all() and (stats == es).all()


def test_long_reads_many_windows(ctx, oracle):
    """config 5 shape at test size: 10 kbp reads, ~89 windows each, all size classes of the sorter"""
    w = World(G=6, L=100000, npairs=4)
    ctx.upload_taxonomy(w.tabs.lineage, w.tabs.rank, w.G)
    # identical genomes make every feature hit 3 locations per shard -> thousands of matches per read
    w.genomes = [w.genomes[g % 2].copy() for g in range(6)]
    w.buf, w.offs = po.concat(w.genomes)
    od = w.oracle_shards(oracle, 2)
    gd = w.gpu_shards(api, ctx, 2)
    buf, offs, _ = synth.make_long_reads(w.genomes, 24, readlen=10000, sub=0.02)
    extra = [w.genomes[0][:90000].copy(), w.genomes[1][100:300].copy()]     # > 16384 matches: global-memory class
    b_all, o_all = po.concat([buf[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)] + extra)
    q = api.make_query_opts(maxcand=4, lowest=4, hitmin=4, hitdiff=0.8)
    oq = po.make_qopt(maxcand=4, lowest=4, hitmin=4, hitdiff=0.8)
    for r in range(2):
        cnt, tg, wn = gd[r].query_matches(q, b_all, o_all)
        pos = 0
        for i in range(len(o_all) - 1):
            mt, mw = od[r].matches(oq, b_all[o_all[i]:o_all[i + 1]])
            assert cnt[i] == len(mt)
            assert (tg[pos:pos + cnt[i]] == mt).all() and (wn[pos:pos + cnt[i]] == mw).all()
            pos += int(cnt[i])
        assert cnt.max() > 16384 and cnt.min() < 512
    ct, ch, cn, mt, mh, mn, best, brank, stats = _gpu_full(ctx, gd, q, b_all, o_all, None, None)
    et, eh, en, eb, er, es = oracle.query_batch(od, w.tax, oq, b_all, o_all, None, None)
    assert (mn == en).all() and (mt == et).all() and (mh == eh).all() and (best == eb).all() and (stats == es).all()


def test_query_window_overrides(ctx, oracle):
    """-winlen / -winstride / -sketchlen given at query time (query_options.cpp:40-75)"""
    w = World(G=6, L=20000, npairs=200)
    ctx.upload_taxonomy(w.tabs.lineage, w.tabs.rank, w.G)
    od = w.oracle_shards(oracle, 2)
    gd = w.gpu_shards(api, ctx, 2)
    for sk, wl, st in ((8, 128, 113), (16, 100, 100), (24, 150, 150)):
        q = api.make_query_opts(sketchlen=sk, winlen=wl, winstride=st, maxcand=4, lowest=4, hitmin=2, hitdiff=0.8, insertsize=600)
        oq = po.make_qopt(sketchlen=sk, winlen=wl, stride=st, maxcand=4, lowest=4, hitmin=2, hitdiff=0.8, insertsize=600)
        ct, ch, cn, mt, mh, mn, best, brank, stats = _gpu_full(ctx, gd, q, w.b1, w.o1, w.b2, w.o2)
        et, eh, en, eb, er, es = oracle.query_batch(od, w.tax, oq, w.b1, w.o1, w.b2, w.o2)
        assert (mn == en).all() and (mt == et).all() and (mh == eh).all() and (best == eb).all() and (stats == es).all()


@pytest.mark.parametrize("k,s,wl,st,maxlocs", [(12, 32, 256, 241, 10), (16, 8, 64, 49, 254), (9, 16, 128, 120, 3)])
def test_nondefault_database_parameters(ctx, oracle, k, s, wl, st, maxlocs):
    """databases built with other k / sketchlen / window parameters and a small location cap, queried with the
    database's own parameters (mode_build.cpp:93-138): wide-sketch K1, 2 and 4 feature slots per lane in the
    fused kernel, short k-mers (many chance hits, long lists, capping)"""
    w = World(G=8, L=30000, npairs=300, seed=13)
    ctx.upload_taxonomy(w.tabs.lineage, w.tabs.rank, w.G)
    kw = dict(k=k, sketchlen=s, winlen=wl, maxlocs=maxlocs)
    od = w.oracle_shards(oracle, 2, stride=st, **kw)
    gd = w.gpu_shards(api, ctx, 2, winstride=st, **kw)
    from test_gpu_build import _same
    for r in range(2):
        _same(gd[r].export(), od[r].export())
    q = api.make_query_opts(sketchlen=s, winlen=wl, winstride=st, maxcand=4, lowest=4, hitmin=2, hitdiff=0.8)
    oq = po.make_qopt(sketchlen=s, winlen=wl, stride=st, maxcand=4, lowest=4, hitmin=2, hitdiff=0.8)
    ct, ch, cn, mt, mh, mn, best, brank, stats = _gpu_full(ctx, gd, q, w.b1, w.o1, w.b2, w.o2)
    ot, oh, on = _oracle_local(od, w.tax, oq, w.b1, w.o1, w.b2, w.o2, 4)
    assert (cn == on).all() and (ct == ot).all() and (ch == oh).all()
    et, eh, en, eb, er, es = oracle.query_batch(od, w.tax, oq, w.b1, w.o1, w.b2, w.o2)
    assert (mn == en).all() and (mt == et).all() and (mh == eh).all() and (best == eb).all() and (stats == es).all()


def test_host_path_sub_batches(tmp_path):
    """mcg_query_batch cuts a batch into pipelined sub-batches (MCG_QUERY_SUB reads each): the
    results must not depend on the cut.  Own processes: the knob is read once per process."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from __graft_entry__ import load_package; load_package()\n"
        "from metacache_mpi_b200 import api, synth, taxonomy\n"
        "ctx = api.Context(0)\n"
        "G = 6\n"
        "genomes = [synth.genome(g, 60000) for g in range(G)]\n"
        "ids, par, rk, species = synth.make_taxonomy(G)\n"
        "tabs = taxonomy.TaxonomyTables(ids, par, rk, species)\n"
        "ctx.upload_taxonomy(tabs.lineage, tabs.rank, G)\n"
        "db = api.Database(ctx)\n"
        "db.add_targets(np.concatenate(genomes), np.arange(G + 1, dtype=np.int64) * 60000, np.arange(G, dtype=np.uint32))\n"
        "db.finalize()\n"
        "b1, o1, b2, o2, truth = synth.make_read_pairs(genomes, 1000, seed=3)\n"
        "q = api.make_query_opts(maxcand=4, lowest=4, hitmin=4, hitdiff=0.8)\n"
        "ct, ch, cn = db.query_batch(q, b1, o1, b2, o2)\n"
        "st, sh, sn = db.query_batch(q, b1, o1, None, None)\n"
        "np.savez(sys.argv[1], ct=ct, ch=ch, cn=cn, st=st, sh=sh, sn=sn)\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),)
    outs = []
    for sub in ("37", "1000000"):
        f = str(tmp_path / ("q%s.npz" % sub))
        r = subprocess.run([sys.executable, "-c", code, f], env=dict(os.environ, MCG_QUERY_SUB=sub), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(np.load(f))
    for k in ("ct", "ch", "cn", "st", "sh", "sn"):
        assert (outs[0][k] == outs[1][k]).all(), k
    assert (outs[0]["cn"] > 0).sum() > 900


def test_fused_kernel_variants_agree(tmp_path):
    """the per-read top lists must not depend on which path a read takes: fused kernel holding 128 matches,
    holding 256 (databases with long lists), no fusion at all (staged path for every read), left-over reads
    not compacted.  40 genomes sharing a block put reads into every class.  Own processes: the knobs are
    read once per process."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from __graft_entry__ import load_package; load_package()\n"
        "from metacache_mpi_b200 import api\n"
        "from util import World\n"
        "ctx = api.Context(0)\n"
        "w = World(G=40, L=12000, npairs=1500, seed=11)\n"
        "ctx.upload_taxonomy(w.tabs.lineage, w.tabs.rank, w.G)\n"
        "db = w.gpu_shards(api, ctx, 1)[0]\n"
        "q = api.make_query_opts(maxcand=4, lowest=4, hitmin=4, hitdiff=0.8)\n"
        "ct, ch, cn = db.query_batch(q, w.b1, w.o1, w.b2, w.o2)\n"
        "np.savez(sys.argv[1], ct=ct, ch=ch, cn=cn)\n"
    ) % (ROOT_DIR, os.path.join(ROOT_DIR, "tests"))
    outs = []
    for i, env in enumerate(({"MCG_FUSED_CAP": "128"}, {"MCG_FUSED_CAP": "256"}, {"MCG_QUERY_NO_FUSION": "1"},
                             {"MCG_FUSED_CAP": "128", "MCG_QUERY_NO_COMPACT": "1"})):
        f = str(tmp_path / ("v%d.npz" % i))
        r = subprocess.run([sys.executable, "-c", code, f], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(np.load(f))
    for o in outs[1:]:
        for k in ("ct", "ch", "cn"):
            assert (outs[0][k] == o[k]).all(), k
    assert (outs[0]["cn"] > 0).sum() > 1400


@pytest.mark.parametrize("P", [2, 4, 8, 3])
def test_reference_compat_merge(ctx, oracle, P):
    """quirk Q1 (sequence-level taxa of other ranks lost on the uint32 wire, querying.h:958,984,1001) along the
    reference's reduction tree: merge kernel == oracle, whose compat mode is pinned against the reference binary
    (tests/test_oracle_cli_golden.py).  Targets 1 and 4 hang below no species, so at -lowest species lifted and
    sequence-level candidates mix."""
    w = World(G=8, L=30000, npairs=400, seed=17)
    species = w.species.copy()
    species[[1, 4]] = 0
    from metacache_mpi_b200 import taxonomy
    tabs = taxonomy.TaxonomyTables(w.ids, w.par, w.rk, species)
    tax = po.TaxTables(tabs.lineage, tabs.rank, w.G)
    ctx.upload_taxonomy(tabs.lineage, tabs.rank, w.G)
    od = w.oracle_shards(oracle, P)
    gd = w.gpu_shards(api, ctx, P)
    differs = False
    for maxcand, lowest, hitmin, hitdiff in CONFIGS + [(32, 4, 2, 0.3)]:
        res = {}
        for compat in (0, 1):
            q = api.make_query_opts(maxcand=maxcand, lowest=lowest, hitmin=hitmin, hitdiff=hitdiff, reference_compat=compat)
            oq = po.make_qopt(maxcand=maxcand, lowest=lowest, hitmin=hitmin, hitdiff=hitdiff, reference_compat=compat)
            ct, ch, cn, mt, mh, mn, best, brank, stats = _gpu_full(ctx, gd, q, w.b1, w.o1, w.b2, w.o2)
            et, eh, en, eb, er, es = oracle.query_batch(od, tax, oq, w.b1, w.o1, w.b2, w.o2)
            assert (mn == en).all() and (mt == et).all() and (mh == eh).all()
            assert (best == eb).all() and (brank == er).all() and (stats == es).all()
            res[compat] = best.copy()
        differs |= bool((res[0] != res[1]).any())
    assert differs


def test_query_argument_checks(ctx, oracle):
    """a shard whose target ids exceed the uploaded taxonomy, candidate taxa outside the tables and negative
    thresholds are refused instead of read out of bounds"""
    w = World(G=6, L=20000, npairs=20)
    gd = w.gpu_shards(api, ctx, 1)
    ctx.upload_taxonomy(w.tabs.lineage[:], w.tabs.rank[:], 3)          # taxonomy knows 3 of the 6 targets
    q = api.make_query_opts(maxcand=4, lowest=4, hitmin=4, hitdiff=0.8)
    with pytest.raises(api.McgError, match="target id"):
        gd[0].query_batch(q, w.b1, w.o1, w.b2, w.o2)
    ctx.upload_taxonomy(w.tabs.lineage, w.tabs.rank, w.G)
    with pytest.raises(api.McgError, match="hitmin"):
        gd[0].query_batch(api.make_query_opts(hitmin=-1), w.b1, w.o1, w.b2, w.o2)
    ct = np.full((1, 4, 4), -1, np.int32); ch = np.zeros((1, 4, 4), np.uint32); cn = np.ones((1, 4), np.int32)
    ct[0, :, 0] = 10 ** 6
    ch[0, :, 0] = 9
    with pytest.raises(api.McgError, match="taxon"):
        ctx.merge_classify(q, ct, ch, cn)
    ct[0, :, 0] = 0
    ctx.merge_classify(q, ct, ch, cn)                                   # the flag does not stick
