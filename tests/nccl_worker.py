"""One rank of the multi-GPU parity test (launched by tests/test_gpu_nccl.py through torch.distributed.run,
one process per GPU, NCCL).  Runs the PRODUCT pipeline -- libmcgpu's own communicator, mcg_query_batch_sharded
(slice ingest -> K1 on the slice -> all-gather of sketches -> shard query -> all-to-all of top lists -> merge +
decision of the slice) and mcg_db_remove_overpopulated_dist -- and compares, on rank 0, with the single-process
oracle at the same number of shards: merged lists, taxa, decision ranks and the 22 counters, bit-exact."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

from __graft_entry__ import load_package

load_package()
from metacache_mpi_b200 import api, distributed as D, synth, taxonomy  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from util import World  # noqa: E402


def gather0(obj):
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(obj, out, dst=0)
    return out


def check_query(tag, w, tax, tabs, db, comm, od, O, opts, reads, bounds=None, dev=False, ctx=None):
    rank, world = dist.get_rank(), dist.get_world_size()
    b1, o1, b2, o2 = reads
    n = len(o1) - 1
    bnd = bounds if bounds is not None else D.slice_bounds(n, world)
    lo, hi = bnd[rank], bnd[rank + 1]
    s1, so1 = b1[o1[lo]:o1[hi]], (o1[lo:hi + 1] - o1[lo]).astype(np.int64)
    s2 = so2 = None
    if b2 is not None:
        s2, so2 = b2[o2[lo]:o2[hi]], (o2[lo:hi + 1] - o2[lo]).astype(np.int64)
    q = api.make_query_opts(**opts)
    if not dev:
        res = db.query_batch_sharded(comm, q, s1, so1, s2, so2)
    else:
        import ctypes as C
        dv = torch.device("cuda", torch.cuda.current_device())
        m = hi - lo
        K = q.maxcand

        def up(a):
            return torch.from_numpy(np.concatenate([np.ascontiguousarray(a).view(np.uint8).ravel(), np.zeros(64, np.uint8)])).to(dv)
        d1, do1 = up(s1 if m else np.zeros(1, np.uint8)), up(so1)
        d2 = do2 = None
        if s2 is not None:
            d2, do2 = up(s2 if m else np.zeros(1, np.uint8)), up(so2)
        ot = torch.empty((max(m, 1), K), dtype=torch.int32, device=dv); oh = torch.empty((max(m, 1), K), dtype=torch.int32, device=dv)
        on = torch.empty(max(m, 1), dtype=torch.int32, device=dv); ob = torch.empty(max(m, 1), dtype=torch.int32, device=dv)
        obr = torch.empty(max(m, 1), dtype=torch.int32, device=dv); st = torch.zeros(22, dtype=torch.int64, device=dv)
        torch.cuda.synchronize()
        rc = db.L.mcg_query_batch_sharded_dev(db.h, comm.h, C.byref(q), d1.data_ptr(), do1.data_ptr(), int(so1[-1]),
                                              d2.data_ptr() if d2 is not None else None, do2.data_ptr() if do2 is not None else None,
                                              int(so2[-1]) if so2 is not None else 0, m, 0,
                                              ot.data_ptr(), oh.data_ptr(), on.data_ptr(), ob.data_ptr(), obr.data_ptr(), st.data_ptr())
        assert rc == 0, db.L.mcg_last_error().decode()
        ctx.sync()
        res = (ot.cpu().numpy()[:m], oh.cpu().numpy().view(np.uint32)[:m], on.cpu().numpy()[:m], ob.cpu().numpy()[:m],
               obr.cpu().numpy()[:m], st.cpu().numpy().astype(np.uint64))
    allres = gather0(res)
    if rank != 0:
        return
    oq = po.make_qopt(**opts)
    et, eh, en, eb, er, es = O.query_batch(od, tax, oq, b1, o1, b2, o2)
    mt = np.concatenate([r[0] for r in allres]); mh = np.concatenate([r[1] for r in allres]); mn = np.concatenate([r[2] for r in allres])
    best = np.concatenate([r[3] for r in allres]); brank = np.concatenate([r[4] for r in allres])
    stats = np.sum([r[5] for r in allres], axis=0)
    assert len(mn) == n, (tag, len(mn), n)
    assert (mn == en).all() and (mt == et).all() and (mh == eh).all(), tag
    assert (best == eb).all() and (brank == er).all(), tag
    assert (stats == es).all(), (tag, stats, es)
    print("  ok %-34s %d reads, classified %d" % (tag, n, int(es[20])), flush=True)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = api.Context(local)
    comm = D.make_comm(ctx)
    assert comm.rank == rank and comm.nranks == world
    O = po.Oracle()
    # tie-heavy world (a stride-aligned block shared by most genomes: equal hit counts across shards);
    # genomes 1 and 4 hang below no species (sequence-level and lifted candidates mix)
    w = World(G=13, L=30000, npairs=3000, seed=23)
    species = w.species.copy()
    species[[1, 4]] = 0
    tabs = taxonomy.TaxonomyTables(w.ids, w.par, w.rk, species)
    tax = po.TaxTables(tabs.lineage, tabs.rank, w.G)
    ctx.upload_taxonomy(tabs.lineage, tabs.rank, w.G)
    ids = np.arange(w.G, dtype=np.uint32)

    # ---- -remove-overpopulated-features across ranks == oracle
    db = api.Database(ctx, shard=rank, nshards=world, maxlocs=3)
    db.add_targets(w.buf, w.offs, ids)
    db.finalize()
    removed_here, list_len = db.remove_overpopulated_dist(comm)
    mine = db.export()
    allx = gather0((mine, removed_here, list_len))
    if rank == 0:
        od = w.oracle_shards(O, world, maxlocs=3)
        before = [set(d.export()[0].tolist()) for d in od]
        O.remove_overpopulated(od)
        want_rem = set().union(*[b - set(d.export()[0].tolist()) for b, d in zip(before, od)])
        assert len(want_rem) > 0
        for r in range(world):
            for a, b in zip(allx[r][0], od[r].export()):
                assert a.shape == b.shape and (a == b).all(), "shard %d after overpopulated-feature removal" % r
            assert allx[r][2] == len(want_rem), (allx[r][2], len(want_rem))
        print("  ok remove_overpopulated_dist: %d keys removed, shards identical to the oracle" % len(want_rem), flush=True)
    db.close()

    # ---- query
    db = api.Database(ctx, shard=rank, nshards=world)
    db.add_targets(w.buf, w.offs, ids)
    db.finalize()
    od = w.oracle_shards(O, world) if rank == 0 else None
    reads = (w.b1, w.o1, w.b2, w.o2)
    n = len(w.o1) - 1
    base = dict(maxcand=4, lowest=4, hitmin=4, hitdiff=0.8)
    check_query("authors' flags", w, tax, tabs, db, comm, od, O, base, reads)
    check_query("reference defaults", w, tax, tabs, db, comm, od, O, dict(maxcand=2, lowest=0, hitmin=5, hitdiff=1.0), reads)
    check_query("reference-compat defaults", w, tax, tabs, db, comm, od, O,
                dict(maxcand=2, lowest=0, hitmin=5, hitdiff=1.0, reference_compat=1), reads)
    check_query("reference-compat species", w, tax, tabs, db, comm, od, O, dict(base, reference_compat=1), reads)
    check_query("wide list, family", w, tax, tabs, db, comm, od, O, dict(maxcand=16, lowest=10, hitmin=1, hitdiff=0.1), reads)
    # ragged slices: one rank without reads, one with most of them
    bnd = [0] + [n // 7] * (world - 1) + [n]
    bnd[1] = 0 if world > 2 else n // 7
    check_query("ragged slices", w, tax, tabs, db, comm, od, O, base, reads, bounds=sorted(bnd))
    check_query("device-resident slice", w, tax, tabs, db, comm, od, O, base, reads, dev=True, ctx=ctx)
    # several parts per call (pipelined stage A / stage B, double-buffered part state)
    os.environ["MCG_SHARD_MIN_PART"] = "50"
    check_query("4 parts per call", w, tax, tabs, db, comm, od, O, base, reads)
    # single-end, ragged read lengths, long reads (staged path: left-over reads on every rank)
    rng = np.random.default_rng(5)
    rl = []
    for L in [0, 4, 16, 50, 128, 129, 241, 500, 2000, 10000] * 6:
        g = w.genomes[int(rng.integers(0, w.G))]
        p = int(rng.integers(0, len(g) - L))
        rl.append(g[p:p + L].copy())
    lb, lo_ = po.concat(rl)
    check_query("single-end ragged + long reads", w, tax, tabs, db, comm, od, O, dict(maxcand=4, lowest=4, hitmin=2, hitdiff=0.8),
                (lb, lo_, None, None))
    db.close()
    comm.close()
    ctx.close()
    dist.barrier()
    if rank == 0:
        print("NCCL_WORKER_OK world=%d" % world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
