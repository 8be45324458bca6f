"""one-GPU scale check (development aid): build a shard of BASELINE.json configs[2] size
(20 Gbp / 8 GPUs = 2.5 Gbp, 625 genomes x 4 Mbp), query it, print timings and counts"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
load_package()
import torch
from metacache_mpi_b200 import api, synth, taxonomy

G = int(sys.argv[1]) if len(sys.argv) > 1 else 625
glen = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
npairs = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = api.Context(0, st.cuda_stream); L = ctx.L
t0 = time.time()
hbuf_t = torch.empty(G * glen, dtype=torch.uint8, pin_memory=True)
hbuf = hbuf_t.numpy()
for g in range(G):
    hbuf[g * glen:(g + 1) * glen] = synth.genome(g, glen)
offs = np.arange(G + 1, dtype=np.int64) * glen
print("generated %.2f Gbp in %.1f s" % (G * glen / 1e9, time.time() - t0), flush=True)
ids, par, rk, species = synth.make_taxonomy(G)
tabs = taxonomy.TaxonomyTables(ids, par, rk, species)
ctx.upload_taxonomy(tabs.lineage, tabs.rank, G)
for it in range(2):
    db = api.Database(ctx)
    t0 = time.time()
    db.add_targets(hbuf, offs, np.arange(G, dtype=np.uint32))
    t1 = time.time()
    db.finalize(); ctx.sync()
    t2 = time.time()
    print("build %.2f Gbp: add_targets %.3f s finalize %.3f s -> %.1f Gbp/s; (keys, nonempty, locations) = %s; free HBM %.1f GB"
          % (G * glen / 1e9, t1 - t0, t2 - t1, G * glen / 1e9 / (t2 - t0), db.counts(), torch.cuda.mem_get_info()[0] / 1e9), flush=True)
    if it == 0:
        db.close()
genomes = [hbuf[g * glen:(g + 1) * glen] for g in range(G)]
b1, o1, b2, o2, truth = synth.make_read_pairs(genomes, npairs, seed=7)
q = api.make_query_opts(maxcand=4, lowest=4, hitmin=4, hitdiff=0.8)
dev = torch.device("cuda")
d1 = torch.from_numpy(np.concatenate([b1, np.zeros(64, np.uint8)])).to(dev); do1 = torch.from_numpy(o1).to(dev)
d2 = torch.from_numpy(np.concatenate([b2, np.zeros(64, np.uint8)])).to(dev); do2 = torch.from_numpy(o2).to(dev)
lt = torch.empty((npairs, 4), dtype=torch.int32, device=dev); lh = torch.empty_like(lt); ln = torch.empty(npairs, dtype=torch.int32, device=dev)
def step():
    rc = L.mcg_query_batch_dev(db.h, C.byref(q), d1.data_ptr(), do1.data_ptr(), len(b1), d2.data_ptr(), do2.data_ptr(), len(b2),
                               npairs, 150, lt.data_ptr(), lh.data_ptr(), ln.data_ptr())
    assert rc == 0, L.mcg_last_error()
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
mt, mh, mn, best, brank, stats = ctx.merge_classify(q, lt.cpu().numpy()[None], lh.cpu().numpy()[None], ln.cpu().numpy()[None])
sp_truth = np.array([tabs.index_of[100000 + int(g)] for g in truth])
print("query %d pairs vs %.2f Gbp shard: %.3f ms per batch, %.1f M reads/s; classified %d, correct species %.4f"
      % (npairs, G * glen / 1e9, ms, 2 * npairs / ms / 1e3, int(stats[20]), float((best == sp_truth).mean())), flush=True)
