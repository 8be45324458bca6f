"""device-timed query batch (development aid)"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
load_package()
import torch
from metacache_mpi_b200 import api, synth, taxonomy

QG = int(sys.argv[1]) if len(sys.argv) > 1 else 64
npairs = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
readlen = int(sys.argv[3]) if len(sys.argv) > 3 else 150
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = api.Context(0, st.cuda_stream); L = ctx.L
genomes = [synth.genome(g, 5_000_000) for g in range(QG)]
ids, par, rk, species = synth.make_taxonomy(QG)
tabs = taxonomy.TaxonomyTables(ids, par, rk, species)
ctx.upload_taxonomy(tabs.lineage, tabs.rank, QG)
db = api.Database(ctx)
db.add_targets(np.concatenate(genomes), np.arange(QG + 1, dtype=np.int64) * 5_000_000, np.arange(QG, dtype=np.uint32))
db.finalize()
print("db", db.counts(), flush=True)
if readlen == 150:
    b1, o1, b2, o2, truth = synth.make_read_pairs(genomes, npairs, seed=7)
else:
    b1, o1, truth = synth.make_long_reads(genomes, npairs, readlen=readlen)
    b2 = o2 = None
q = api.make_query_opts(maxcand=4, lowest=4, hitmin=4, hitdiff=0.8)
dev = torch.device("cuda")
d1 = torch.from_numpy(np.concatenate([b1, np.zeros(64, np.uint8)])).to(dev); do1 = torch.from_numpy(o1).to(dev)
if b2 is not None:
    d2 = torch.from_numpy(np.concatenate([b2, np.zeros(64, np.uint8)])).to(dev); do2 = torch.from_numpy(o2).to(dev)
lt = torch.empty((npairs, 4), dtype=torch.int32, device=dev); lh = torch.empty_like(lt); ln = torch.empty(npairs, dtype=torch.int32, device=dev)
def step():
    rc = L.mcg_query_batch_dev(db.h, C.byref(q), d1.data_ptr(), do1.data_ptr(), len(b1),
                               d2.data_ptr() if b2 is not None else None, do2.data_ptr() if b2 is not None else None,
                               len(b2) if b2 is not None else 0, npairs, readlen, lt.data_ptr(), lh.data_ptr(), ln.data_ptr())
    assert rc == 0, L.mcg_last_error()
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("query batch %d reads(pairs) x %d bp: %.3f ms  %.2f M reads/s  ncand mean %.2f" % (npairs, readlen, ms, (2 if b2 is not None else 1) * npairs / ms / 1e3, float(ln.float().mean())))
