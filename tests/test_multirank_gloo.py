"""N > 1 host logic on CPU: two processes over gloo run the sharded query plumbing
(metacache_mpi_b200.distributed): target partitioning, rank-ordered candidate gather,
overpopulated-feature reduction.  The per-shard compute is stood in by the oracle (this is a
test of the plumbing, not of the kernels); the expectation is the single-process oracle."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    load_package()
    from metacache_mpi_b200 import distributed as D
    from oracle import pyoracle as po
    from util import World
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        O = po.Oracle()
        w = World(G=7, L=20000, npairs=80, seed=21)
        own = D.owned_targets(w.G, rank, world)
        assert list(own) == [g for g in range(w.G) if g % world == rank]
        db = O.db_create(shard=rank, nshards=world, maxlocs=3)
        for g in range(w.G):
            db.add_target(w.genomes[g], g)
        db.finalize()
        # -remove-overpopulated-features across ranks
        k, s, _, _ = db.export()
        rem = D.reduce_key_sizes(k, s, db.maxlocs())
        # reads are broadcast from rank 0
        b1 = torch.from_numpy(w.b1.copy()) if rank == 0 else torch.zeros(len(w.b1), dtype=torch.uint8)
        b2 = torch.from_numpy(w.b2.copy()) if rank == 0 else torch.zeros(len(w.b2), dtype=torch.uint8)
        D.broadcast_reads([b1, b2])
        assert (b1.numpy() == w.b1).all() and (b2.numpy() == w.b2).all()
        K = 4
        oq = po.make_qopt(maxcand=K, lowest=4, hitmin=2, hitdiff=0.8)
        n = len(w.o1) - 1
        lt = np.full((n, K), -1, np.int32); lh = np.zeros((n, K), np.int32); ln = np.zeros(n, np.int32)
        for i in range(n):
            t, h = db.query_local(w.tax, oq, w.b1[w.o1[i]:w.o1[i + 1]], w.b2[w.o2[i]:w.o2[i + 1]])
            ln[i] = len(t); lt[i, :len(t)] = t; lh[i, :len(t)] = h
        gt, gh, gn = D.gather_candidates(torch.from_numpy(lt), torch.from_numpy(lh), torch.from_numpy(ln))
        assert gt.shape == (world, n, K) and (gt[rank].numpy() == lt).all()
        if rank == 0:
            dbs = [O.db_create(shard=r, nshards=world, maxlocs=3) for r in range(world)]
            for r in range(world):
                for g in range(w.G):
                    dbs[r].add_target(w.genomes[g], g)
                dbs[r].finalize()
            # expected removal list = what the single-process restatement removes
            before = [set(d.export()[0].tolist()) for d in dbs]
            O.remove_overpopulated(dbs)
            after = [set(d.export()[0].tolist()) for d in dbs]
            want_rem = sorted(set().union(*[b - a for a, b in zip(after, before)]))
            assert sorted(rem.tolist()) == want_rem and len(want_rem) > 0
            dbs = [O.db_create(shard=r, nshards=world, maxlocs=3) for r in range(world)]
            for r in range(world):
                for g in range(w.G):
                    dbs[r].add_target(w.genomes[g], g)
                dbs[r].finalize()
            et, eh, en, eb, er, es = O.query_batch(dbs, w.tax, oq, w.b1, w.o1, w.b2, w.o2)
            for i in range(n):
                it, ih = [], []
                for r in range(world):
                    c = int(gn[r, i])
                    it += gt[r, i, :c].tolist(); ih += gh[r, i, :c].tolist()
                mt, mh, best, br = O.merge_classify(w.tax, oq, it, ih)
                assert (mt == et[i, :en[i]]).all() and (mh == eh[i, :en[i]]).all() and best == eb[i] and br == er[i]
        q.put((rank, "ok"))
    except Exception as e:      # pragma: no cover
        import traceback
        q.put((rank, "fail: %r\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
