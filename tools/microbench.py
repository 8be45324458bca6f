"""quick device-timed numbers for the individual stages (development aid; bench.py is the contract)"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

load_package()
import torch  # noqa: E402
from metacache_mpi_b200 import api, synth, taxonomy  # noqa: E402


def main():
    mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    stages = sys.argv[2].split(",") if len(sys.argv) > 2 else ["sketch", "build", "query"]
    torch.cuda.init()
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    ctx = api.Context(0, st.cuda_stream)
    L = ctx.L
    G = max(mbp // 5, 1)
    glen = mbp * 1_000_000 // G
    t0 = time.time()
    genomes = [synth.genome(g, glen) for g in range(G)]
    buf = np.concatenate(genomes)
    offs = np.arange(G + 1, dtype=np.int64) * glen
    print("generated %d Mbp in %.1fs" % (mbp, time.time() - t0), flush=True)
    cfg = api.SketchCfg(16, 16, 128, 113)
    nwin = ctx.count_windows(cfg, offs)
    d_seq = torch.empty(buf.size + 64, dtype=torch.uint8, device="cuda")
    d_seq[:buf.size].copy_(torch.from_numpy(buf))
    d_off = torch.from_numpy(offs).cuda()
    d_feat = torch.empty(nwin * 16, dtype=torch.int32, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(5):
        e0.record()
        rc = L.mcg_sketch_windows_dev(ctx.h, C.byref(cfg), d_seq.data_ptr(), buf.size, d_off.data_ptr(), G, nwin,
                                      d_feat.data_ptr(), None)
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0
        ms = e0.elapsed_time(e1)
        print("sketch %d Mbp: %.3f ms  %.1f Gbp/s  alg %.1f GB/s" % (mbp, ms, buf.size / ms / 1e6, buf.size * 1.566 / ms / 1e6), flush=True)
    if stages == ["sketch"]:
        return
    # build
    ids, par, rk, species = synth.make_taxonomy(G)
    tabs = taxonomy.TaxonomyTables(ids, par, rk, species)
    ctx.upload_taxonomy(tabs.lineage, tabs.rank, G)
    for it in range(2):
        db = api.Database(ctx)
        t0 = time.time()
        db.add_targets(buf, offs, np.arange(G, dtype=np.uint32))
        t1 = time.time()
        db.finalize()
        ctx.sync()
        t2 = time.time()
        print("build: add_targets %.3fs finalize %.3fs  counts %s" % (t1 - t0, t2 - t1, db.counts()), flush=True)
        if it == 0:
            db.close()
    # query
    npairs = 200_000
    b1, o1, b2, o2, truth = synth.make_read_pairs(genomes, npairs, seed=7)
    q = api.make_query_opts(maxcand=4, lowest=4, hitmin=4, hitdiff=0.8)
    for it in range(3):
        t0 = time.time()
        ct, ch, cn = db.query_batch(q, b1, o1, b2, o2)
        t1 = time.time()
        mt, mh, mn, best, brank, stats = ctx.merge_classify(q, ct[None], ch[None], cn[None])
        t2 = time.time()
        print("query %d pairs: local %.3fs (%.2f Mreads/s) merge %.3fs classified %d" % (npairs, t1 - t0, 2 * npairs / (t1 - t0) / 1e6, t2 - t1, int(stats[20])), flush=True)
    sp = tabs.index_of
    ok = sum(1 for i in range(npairs) if best[i] >= 0)
    print("launches", ctx.launches())


if __name__ == "__main__":
    main()
