"""Timing of the distributed -remove-overpopulated-features reduction (mcg_db_remove_overpopulated_dist, comm.cu) on a
shard of BASELINE.json configs[2] size: 625 genomes x 4 Mbp = 2.5 Gbp per GPU, ~2 x 10^8 (key, size) pairs per rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ovp_bench.py [Mbp per GPU]

The synthetic genomes hold no feature with more than 253 locations, so the removal list is empty; every (key, size) pair
still travels to its owner (key % N), is sorted and summed there -- the whole data movement of mode_build.cpp:847-1074."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

load_package()
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from metacache_mpi_b200 import api, distributed as D, synth  # noqa: E402

GLEN = 4_000_000
mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = api.Context(local)
comm = D.make_comm(ctx)
Gper = max(mbp * 1_000_000 // GLEN, 1)
own = np.arange(rank, Gper * world, world, dtype=np.uint32)
genomes = torch.empty(Gper * GLEN + 64, dtype=torch.uint8, device=dev)
for i, g in enumerate(own):
    synth.genome_fast(int(g), GLEN, dev, out=genomes[i * GLEN:(i + 1) * GLEN])
goffs = np.arange(Gper + 1, dtype=np.int64) * GLEN
torch.cuda.synchronize()
db = api.Database(ctx, shard=rank, nshards=world)
for a in range(0, Gper, 125):
    b = min(Gper, a + 125)
    rc = ctx.L.mcg_db_add_targets_dev(db.h, genomes.data_ptr(), goffs[a:b + 1].ctypes.data, own[a:b].ctypes.data, b - a, None)
    assert rc == 0, ctx.L.mcg_last_error().decode()
db.finalize()
ctx.sync()
nk, ne, nl = db.counts()
for it in range(3):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    removed, listlen = db.remove_overpopulated_dist(comm)
    ctx.sync()
    ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print("remove_overpopulated_dist, %d rank(s) x %.1f Gbp: call %d: %.1f ms (max over ranks); rank 0: %d keys with locations, "
              "removal list %d, removed here %d" % (world, Gper * GLEN / 1e9, it, ms, ne, listlen, removed), flush=True)
db.close()
comm.close()
if world > 1:
    dist.destroy_process_group()
