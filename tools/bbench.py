"""device-side build breakdown (development aid)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
load_package()
import torch
from metacache_mpi_b200 import api, synth
mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
G = mbp // 5
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = api.Context(0, st.cuda_stream)
hb = torch.empty(G * 5_000_000, dtype=torch.uint8, pin_memory=True); buf = hb.numpy()
for g in range(G): buf[g * 5_000_000:(g + 1) * 5_000_000] = synth.genome(g, 5_000_000)
offs = np.arange(G + 1, dtype=np.int64) * 5_000_000
ids = np.arange(G, dtype=np.uint32)
for it in range(3):
    db = api.Database(ctx)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    db.add_targets(buf, offs, ids)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    db.finalize()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("build %d Mbp (pinned): add_targets %.1f ms finalize %.1f ms -> %.2f Gbp/s; counts %s" % (mbp, 1e3 * (t1 - t0), 1e3 * (t2 - t1), mbp / 1e3 / (t2 - t0), db.counts()), flush=True)
    db.close()
