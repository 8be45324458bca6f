"""N > 1 host logic on CPU: two processes over gloo walk through the launcher side of the sharded query
(metacache_mpi_b200.distributed): target partitioning, batch slicing, the rendezvous that carries the
communicator id, and the order in which per-shard lists meet in the merge.  The per-shard compute is stood in
by the oracle (this is a test of the plumbing and of the slice-wise merge schedule, not of the kernels; the GPU
data plane itself is tested by tests/test_gpu_nccl.py); the expectation is the single-process oracle."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
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
        w = World(G=7, L=20000, npairs=81, seed=21)
        own = D.owned_targets(w.G, rank, world)
        assert list(own) == [g for g in range(w.G) if g % world == rank]
        # the communicator id reaches every rank unchanged (what make_comm hands to mcg_comm_create)
        box = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        assert box[0] == bytes(range(128))
        db = O.db_create(shard=rank, nshards=world)
        for g in range(w.G):
            db.add_target(w.genomes[g], g)
        db.finalize()
        K = 4
        oq = po.make_qopt(maxcand=K, lowest=4, hitmin=2, hitdiff=0.8)
        n = len(w.o1) - 1
        bnd = D.slice_bounds(n, world)
        assert bnd[0] == 0 and bnd[-1] == n and all(b1 >= b0 for b0, b1 in zip(bnd, bnd[1:]))
        # every shard answers every read (reads are replicated by the all-gather of their sketches) ...
        lists = []
        for i in range(n):
            t, h = db.query_local(w.tax, oq, w.b1[w.o1[i]:w.o1[i + 1]], w.b2[w.o2[i]:w.o2[i + 1]])
            lists.append((list(t), list(h)))
        # ... and the lists of slice r travel to rank r (the all-to-all), which merges them in rank order
        send = [lists[bnd[r]:bnd[r + 1]] for r in range(world)]
        recv = [None] * world
        dist.all_to_all_object_list(recv, send) if hasattr(dist, "all_to_all_object_list") else None
        if recv[0] is None:                                   # older torch: all-gather everything, pick the own slice
            allsend = [None] * world
            dist.all_gather_object(allsend, send)
            recv = [allsend[r][rank] for r in range(world)]
        merged = []
        for i in range(bnd[rank + 1] - bnd[rank]):
            it, ih = [], []
            for r in range(world):
                it += recv[r][i][0]; ih += recv[r][i][1]
            merged.append(O.merge_classify(w.tax, oq, it, ih))
        allm = [None] * world
        dist.all_gather_object(allm, [(m[0].tolist(), m[1].tolist(), m[2], m[3]) for m in merged])
        if rank == 0:
            dbs = w.oracle_shards(O, world)
            et, eh, en, eb, er, es = O.query_batch(dbs, w.tax, oq, w.b1, w.o1, w.b2, w.o2)
            flat = [m for part in allm for m in part]
            assert len(flat) == n
            for i, (mt, mh, best, br) in enumerate(flat):
                assert mt == et[i, :en[i]].tolist() and mh == eh[i, :en[i]].tolist() and best == eb[i] and br == er[i]
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
