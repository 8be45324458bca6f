#!/usr/bin/env python3
"""bench.py -- the measurement contract of this repository.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl mcgpu|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline (BASELINE.json configs[1]): sketching-only throughput of 1 Gbp of synthetic
reference genomes at k=16 / sketchlen=16 / winlen=128 (stride 113), per GPU.  A "step" is one
pass of the sketch path (window bookkeeping + K1) over that batch.

  value : bp/s, whole job, inputs resident in HBM, CUDA events on the launching stream,
          max over ranks.  With N > 1 every rank sketches its own 1 Gbp of genomes (the
          reference partitions targets by target_id % ranks: no data-path collective) => weak.
  e2e   : the same metric through the C-ABI call mcg_sketch_windows() with pinned HOST
          buffers: H2D of the sequence bytes and D2H of the sketches inside the timed region.
  roofline      : the sketch kernel against the measured HBM copy bandwidth
                  (MEASURED_PEAKS.json) with 1.566 algorithmic bytes per bp.
  cpu_baseline  : the reference's own sketcher (oracle/_ref, all host threads) on a bounded
                  sample; a reported baseline, not the target.
  build / query : extra objects: device-timed database build (bp/s) and query (reads/s) of
                  the same path at a smaller synthetic size, with their own byte accounting;
                  at N > 1 the query shards the database by target_id % N, every GPU queries
                  all reads and the per-rank top lists are gathered over NCCL before the merge.

--impl reference times the reference's CPU implementation of the headline path (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
print_json = None

K, S, WINLEN, STRIDE = 16, 16, 128, 113
GENOME_LEN = 5_000_000
ALG_BYTES_PER_BP = 1.0 + 4.0 * S / STRIDE        # 1 B in + 64 B per 113 bp out
FALLBACK_HBM_GBS = 6650.0                         # B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="mcgpu", choices=["mcgpu", "reference"])
    ap.add_argument("--mbp", type=int, default=1000, help="Mbp of reference genomes per GPU (default = configs[1])")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-extras", action="store_true", help="skip the build/query extra objects")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    return ap.parse_args()


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def load_traffic():
    """dram bytes per bp of the sketch kernel from the committed ncu capture (profiles/traffic.json)"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return {}
    return {}


class ClockSampler:
    """samples SM clock / throttle reasons of one GPU through NVML while the timed region runs"""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_genome_block(first_genome, count, glen):
    from __graft_entry__ import load_package
    load_package()
    from metacache_mpi_b200 import synth
    buf = np.empty(count * glen, np.uint8)
    for i in range(count):
        buf[i * glen:(i + 1) * glen] = synth.genome(first_genome + i, glen)
    offs = np.arange(count + 1, dtype=np.int64) * glen
    return buf, offs


# ----------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    if po.have_reference():
        impl, kind = po.Reference(), "reference"
    else:
        impl, kind = po.Oracle(), "port"
    # bounded sample of the same workload: ~22 Mbp/s/core => a few seconds per step
    sample_mbp = int(min(args.mbp, max(10, 16 * cores)))
    G = max(sample_mbp // 5, 1)
    buf, offs = make_genome_block(0, G, GENOME_LEN)
    bp = int(offs[-1])
    times = []
    for it in range(args.warmup + args.steps):
        t, cs, nw = impl.sketch_throughput(buf, offs, K, S, WINLEN, STRIDE, threads=cores)
        if it >= args.warmup:
            times.append(t)
    tot = float(np.sum(times))
    val = bp * len(times) / tot
    line = {
        "impl": "reference", "metric": "sketch_bp_per_s", "value": val, "unit": "bp/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "sketching-only: %d Mbp of synthetic reference genomes per GPU (%d x 5 Mbp), k=16 sketchlen=16 "
                               "winlen=128 winstride=113 (BASELINE.json configs[1])" % (args.mbp, max(args.mbp // 5, 1)),
                   "sample": "%d Mbp per step: the first %d genomes of that workload (bounded sample, same generator and seeds)"
                             % (bp // 1_000_000, G)},
        "cpu_baseline": {"value": val, "unit": "bp/s", "cores": cores, "kind": kind,
                         "sample": "%d genomes x 5 Mbp, %d threads over sequences, hash_dna.h sketcher" % (G, cores)},
        "e2e": {"value": val, "unit": "bp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_json(line)


# ----------------------------------------------------------------------------- our arm
def run_mcgpu(args):
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    load_package()
    from metacache_mpi_b200 import api, synth, taxonomy

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout at INFO/VERSION level: keep stdout to the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE"):
            os.environ["NCCL_DEBUG_FILE"] = os.environ.get("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = api.Context(local, stream.cuda_stream)        # raises if libmcgpu.so / the GPU is missing
    L = ctx.L

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- workload: G genomes of 5 Mbp per rank (targets g with g % world == rank)
    G = max(args.mbp // 5, 1)
    hbuf_t = torch.empty(G * GENOME_LEN, dtype=torch.uint8, pin_memory=True)
    hbuf = hbuf_t.numpy()
    for i in range(G):
        hbuf[i * GENOME_LEN:(i + 1) * GENOME_LEN] = synth.genome(rank + i * world, GENOME_LEN)
    offs = np.arange(G + 1, dtype=np.int64) * GENOME_LEN
    bp = int(offs[-1])
    cfg = api.SketchCfg(K, S, WINLEN, STRIDE)
    nwin = ctx.count_windows(cfg, offs)
    d_seq = torch.empty(bp + 64, dtype=torch.uint8, device=dev)
    d_seq[:bp].copy_(hbuf_t, non_blocking=True)
    d_off = torch.from_numpy(offs).to(dev)
    d_feat = torch.empty(nwin * S, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    def step_dev():
        rc = L.mcg_sketch_windows_dev(ctx.h, C.byref(cfg), d_seq.data_ptr(), bp, d_off.data_ptr(), G, nwin,
                                      d_feat.data_ptr(), None)
        if rc != 0:
            raise RuntimeError(L.mcg_last_error().decode())

    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        step_dev()
    barrier()
    l0 = ctx.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launches() - l0
    value = bp * world * args.steps / (dev_ms * 1e-3)

    # ---- e2e: pinned host buffers through the C-ABI host entry point
    hfeat_t = torch.empty(nwin * S, dtype=torch.int32, pin_memory=True)
    hfeat = hfeat_t.numpy().view(np.uint32)
    nw_out = C.c_int64()

    def step_e2e():
        rc = L.mcg_sketch_windows(ctx.h, C.byref(cfg), hbuf.ctypes.data, offs.ctypes.data, G, hfeat.ctypes.data, None,
                                  C.byref(nw_out))
        if rc != 0:
            raise RuntimeError(L.mcg_last_error().decode())

    for _ in range(min(args.warmup, 3)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        step_e2e()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    sampler.stop()
    e2e_value = bp * world * args.e2e_steps / e2e_s
    # parity spot check of what was just timed: device-resident result == host-path result
    same = bool((d_feat.cpu().numpy().view(np.uint32) == hfeat).all())

    peak, peak_src = hbm_peak()
    kern_ms = dev_ms / args.steps
    achieved = ALG_BYTES_PER_BP * bp / (kern_ms * 1e-3) / 1e9
    tr = load_traffic().get("sketch_kernel")
    traffic = None
    if tr:
        traffic = tr["dram_bytes_per_bp"] * bp
    line = {
        "metric": "sketch_bp_per_s", "value": value, "unit": "bp/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": kern_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "sketching-only: %d Mbp of synthetic reference genomes per GPU (%d x 5 Mbp), k=16 sketchlen=16 "
                               "winlen=128 winstride=113 (BASELINE.json configs[1])" % (bp // 1_000_000, G),
                   "l2": "inputs (%.2f GB per step) larger than L2, no flush needed" % (bp / 1e9),
                   "parallelism": "targets sharded by target_id %% %d, no data-path collective" % world,
                   "self_check": "device-resident and host-path sketches identical: %s" % same},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": "bp/s", "h2d_bytes_per_step": bp + 8 * (G + 1), "d2h_bytes_per_step": nwin * S * 4,
                "steps": args.e2e_steps, "api": "mcg_sketch_windows (pinned host buffers)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "mcg::sketch_tpw_kernel",
                     "algorithmic_bytes_per_bp": ALG_BYTES_PER_BP,
                     "note": "instruction-throughput bound kernel (ncu: ALU pipe 86 %, issue slots 79 % busy, see profiles/): "
                             "the HBM fraction is reported as required, integer issue is the binding limit"},
    }

    # ---- extras: build + query of the same path
    if not args.no_extras:
        try:
            line.update(run_extras(args, ctx, api, synth, taxonomy, torch, dist, dev, stream, world, rank, barrier, max_over_ranks, peak))
        except Exception as e:       # extras must never hide the headline
            line["extras_error"] = repr(e)

    # ---- CPU baseline on the host cores (rank 0, N = 1 only)
    if world == 1 and not args.no_cpu:
        try:
            from oracle import pyoracle as po
            cores = os.cpu_count() or 1
            impl, kind = (po.Reference(), "reference") if po.have_reference() else (po.Oracle(), "port")
            sg = int(max(2, min(G, 40 * cores)))          # ~20 core-seconds of the reference's sketcher
            t, cs, nw = impl.sketch_throughput(hbuf[:sg * GENOME_LEN], offs[:sg + 1].copy(), K, S, WINLEN, STRIDE, threads=cores)
            line["cpu_baseline"] = {"value": sg * GENOME_LEN / t, "unit": "bp/s", "cores": cores, "kind": kind,
                                    "sample": "%d of the %d genomes (%d Mbp), %d threads, %.1f s" % (sg, G, sg * 5, cores, t)}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": "bp/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}

    if rank == 0:
        print_json(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_extras(args, ctx, api, synth, taxonomy, torch, dist, dev, stream, world, rank, barrier, max_over_ranks, peak):
    """device-timed build (bp/s) and query (reads/s): 64 genomes x 5 Mbp sharded by target_id % world,
    1 M read pairs (2 x 150 bp) per batch, the authors' query flags"""
    L = ctx.L
    out = {}
    QG = 64
    ids, par, rk, species = synth.make_taxonomy(QG)
    tabs = taxonomy.TaxonomyTables(ids, par, rk, species)
    ctx.upload_taxonomy(tabs.lineage, tabs.rank, QG)
    own = [g for g in range(QG) if g % world == rank]
    genomes = {g: synth.genome(g, GENOME_LEN) for g in own}
    buf_t = torch.empty(len(own) * GENOME_LEN, dtype=torch.uint8, pin_memory=True)
    buf = buf_t.numpy()
    for i, g in enumerate(own):
        buf[i * GENOME_LEN:(i + 1) * GENOME_LEN] = genomes[g]
    offs = np.arange(len(own) + 1, dtype=np.int64) * GENOME_LEN
    tids = np.array(own, np.uint32)
    # build: add_targets (H2D + K1) + finalize (sort, CSR, table), timed end to end on the host clock
    # after a warm-up build (allocations), device work dominates
    times = []
    db = None
    for it in range(4):
        if db is not None:
            db.close()
        db = api.Database(ctx, shard=rank, nshards=world)
        barrier()
        t0 = time.perf_counter()
        db.add_targets(buf, offs, tids)
        db.finalize()
        ctx.sync()
        times.append(time.perf_counter() - t0)
    bt = max_over_ranks(float(np.median(times[1:])))
    nk, ne, nl = db.counts()
    bbp = QG * GENOME_LEN
    alg_build = 1.0 * len(buf) + 8.0 * nl + 12.0 * nk
    out["build"] = {"metric": "build_bp_per_s", "value": bbp / bt, "unit": "bp/s", "seconds": bt,
                    "workload": "%d genomes x 5 Mbp sharded over %d GPU(s), host buffers -> finalized HBM multimap" % (QG, world),
                    "shard0_keys": int(nk), "shard0_locations": int(nl),
                    "roofline": {"bound": "hbm", "achieved": alg_build / bt / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": alg_build / bt / 1e9 / peak,
                                 "traffic": (load_traffic().get("build_kernels_1gbp", {}).get("dram_bytes_per_bp") or 0) * len(buf) or None,
                                 "note": "wall time of mcg_db_add_targets + mcg_db_finalize incl. the H2D copy of the sequences from pinned host memory"}}
    # query
    npairs = 1_000_000
    allg = [synth.genome(g, GENOME_LEN) for g in range(QG)] if world > 1 else [genomes[g] for g in range(QG)]
    b1, o1, b2, o2, truth = synth.make_read_pairs(allg, npairs, seed=7)
    del allg
    q = api.make_query_opts(maxcand=4, lowest=4, hitmin=4, hitdiff=0.8)
    Kc = 4
    d1 = torch.from_numpy(np.concatenate([b1, np.zeros(64, np.uint8)])).to(dev)
    d2 = torch.from_numpy(np.concatenate([b2, np.zeros(64, np.uint8)])).to(dev)
    do1 = torch.from_numpy(o1).to(dev)
    do2 = torch.from_numpy(o2).to(dev)
    lt = torch.empty((npairs, Kc), dtype=torch.int32, device=dev)
    lh = torch.empty((npairs, Kc), dtype=torch.int32, device=dev)
    ln = torch.empty(npairs, dtype=torch.int32, device=dev)
    # cross-shard exchange of the per-read top lists.  Preferred: every rank's candidate kernel stores its
    # lists straight into rank 0's [P][n][K] buffers through NVLink peer memory (torch symmetric memory),
    # followed by one barrier -- compute and "gather" are the same kernel.  Fallback: NCCL all-gather.
    gather_mode = "local"
    hdl_t = hdl_h = hdl_n = None
    if world > 1:
        gather_mode = "nccl-all-gather"
        if os.environ.get("MCG_BENCH_NO_P2P", "") == "":
            try:
                import torch.distributed._symmetric_memory as symm
                gt = symm.empty((world, npairs, Kc), dtype=torch.int32, device=dev)
                gh = symm.empty((world, npairs, Kc), dtype=torch.int32, device=dev)
                gn = symm.empty((world, npairs), dtype=torch.int32, device=dev)
                hdl_t = symm.rendezvous(gt, dist.group.WORLD)
                hdl_h = symm.rendezvous(gh, dist.group.WORLD)
                hdl_n = symm.rendezvous(gn, dist.group.WORLD)
                gather_mode = "p2p-store into rank 0 (symmetric memory) + barrier"
            except Exception as e:          # pragma: no cover
                gather_mode = "nccl-all-gather (symmetric memory unavailable: %s)" % type(e).__name__
                hdl_t = None
    if hdl_t is None:
        gt = torch.empty((world, npairs, Kc), dtype=torch.int32, device=dev)
        gh = torch.empty((world, npairs, Kc), dtype=torch.int32, device=dev)
        gn = torch.empty((world, npairs), dtype=torch.int32, device=dev)
    mt = torch.empty((npairs, Kc), dtype=torch.int32, device=dev)
    mh = torch.empty((npairs, Kc), dtype=torch.int32, device=dev)
    mn = torch.empty(npairs, dtype=torch.int32, device=dev)
    best = torch.empty(npairs, dtype=torch.int32, device=dev)
    brank = torch.empty(npairs, dtype=torch.int32, device=dev)
    stats = torch.zeros(22, dtype=torch.int64, device=dev)
    if hdl_t is not None:
        out_t = hdl_t.buffer_ptrs[0] + rank * npairs * Kc * 4
        out_h = hdl_h.buffer_ptrs[0] + rank * npairs * Kc * 4
        out_n = hdl_n.buffer_ptrs[0] + rank * npairs * 4
    else:
        out_t, out_h, out_n = lt.data_ptr(), lh.data_ptr(), ln.data_ptr()

    def qstep():
        rc = L.mcg_query_batch_dev(db.h, C.byref(q), d1.data_ptr(), do1.data_ptr(), len(b1), d2.data_ptr(), do2.data_ptr(), len(b2),
                                   npairs, 150, out_t, out_h, out_n)
        if rc != 0:
            raise RuntimeError(L.mcg_last_error().decode())
        if hdl_t is not None:
            hdl_t.barrier(channel=0)            # all shards' lists have landed in rank 0's memory
            ct, ch, cn = gt, gh, gn
        elif world > 1:
            dist.all_gather_into_tensor(gt.view(world * npairs, Kc), lt)
            dist.all_gather_into_tensor(gh.view(world * npairs, Kc), lh)
            dist.all_gather_into_tensor(gn.view(world * npairs), ln)
            ct, ch, cn = gt, gh, gn
        else:
            ct, ch, cn = lt, lh, ln
        if rank == 0:
            stats.zero_()
            rc = L.mcg_merge_classify_dev(ctx.h, C.byref(q), ct.data_ptr(), ch.data_ptr(), cn.data_ptr(), world, npairs,
                                          mt.data_ptr(), mh.data_ptr(), mn.data_ptr(), best.data_ptr(), brank.data_ptr(),
                                          stats.data_ptr())
            if rc != 0:
                raise RuntimeError(L.mcg_last_error().decode())
        if hdl_t is not None:
            hdl_t.barrier(channel=1)            # rank 0 is done reading before the next batch overwrites

    for _ in range(3):
        qstep()
    barrier()
    l0 = ctx.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    qsteps = 10
    e0.record(stream)
    for _ in range(qsteps):
        qstep()
    e1.record(stream)
    barrier()
    qms = max_over_ranks(e0.elapsed_time(e1)) / qsteps
    launches_q = int((ctx.launches() - l0) / qsteps)
    st = stats.cpu().numpy()
    # the same batch through the host entry point: pinned host reads in, top lists out (H2D / query / D2H pipelined)
    qe2e = None
    try:
        hb1 = torch.from_numpy(b1).pin_memory(); hb2 = torch.from_numpy(b2).pin_memory()
        hot = torch.empty((npairs, Kc), dtype=torch.int32, pin_memory=True)
        hoh = torch.empty((npairs, Kc), dtype=torch.int32, pin_memory=True)
        hon = torch.empty(npairs, dtype=torch.int32, pin_memory=True)

        def qhost():
            rc = L.mcg_query_batch(db.h, C.byref(q), hb1.data_ptr(), o1.ctypes.data, hb2.data_ptr(), o2.ctypes.data, npairs,
                                   hot.data_ptr(), hoh.data_ptr(), hon.data_ptr())
            if rc != 0:
                raise RuntimeError(L.mcg_last_error().decode())
        qhost()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            qhost()
        barrier()
        qe = max_over_ranks(time.perf_counter() - t0) / 3
        same_q = bool((hon.numpy() == ln.cpu().numpy()).all()) if world == 1 else None
        qe2e = {"value": 2 * npairs / qe, "unit": "reads/s", "ms_per_batch": 1e3 * qe,
                "h2d_bytes_per_step": int(len(b1) + len(b2) + 16 * (npairs + 1)), "d2h_bytes_per_step": int(npairs * (8 * Kc + 4)),
                "api": "mcg_query_batch (pinned host buffers, per-shard top lists)", "same_as_device_path": same_q}
    except Exception as e:
        qe2e = {"error": repr(e)}
    # N > 1: every rank copying all reads over PCIe itself (above) is bounded by the host.  The reads of a batch
    # cross PCIe ONCE (rank 0), travel to the other GPUs by an NCCL broadcast over NVLink, every GPU queries its
    # shard, the per-shard top lists land in rank 0 (P2P stores / all-gather), rank 0 merges, classifies and
    # copies the result to the host.
    if world > 1 and isinstance(qe2e, dict) and "error" not in qe2e:
        try:
            hdo1 = torch.from_numpy(o1).pin_memory(); hdo2 = torch.from_numpy(o2).pin_memory()
            hbest = torch.empty(npairs, dtype=torch.int32, pin_memory=True)
            hmt = torch.empty((npairs, Kc), dtype=torch.int32, pin_memory=True)
            hmh = torch.empty((npairs, Kc), dtype=torch.int32, pin_memory=True)
            n1, n2 = len(b1), len(b2)

            # sub-batches: while sub-batch j is broadcast and queried, rank 0 copies sub-batch j + 1 over PCIe
            NSUB = 4
            assert npairs % NSUB == 0
            ns = npairs // NSUB
            rel = torch.from_numpy(np.arange(ns + 1, dtype=np.int64) * 150).pin_memory()     # offsets of a sub-batch
            drel = torch.empty(ns + 1, dtype=torch.int64, device=dev)
            copy_stream = torch.cuda.Stream(device=dev)
            ev_in = [torch.cuda.Event() for _ in range(NSUB)]
            ev_free = torch.cuda.Event()

            def qbcast():
                cur = torch.cuda.current_stream()
                if rank == 0:
                    ev_free.record(cur)                       # the previous batch no longer reads d1 / d2
                    with torch.cuda.stream(copy_stream):
                        copy_stream.wait_event(ev_free)
                        drel.copy_(rel, non_blocking=True)
                        for j in range(NSUB):
                            d1[j * ns * 150:(j + 1) * ns * 150].copy_(hb1[j * ns * 150:(j + 1) * ns * 150], non_blocking=True)
                            d2[j * ns * 150:(j + 1) * ns * 150].copy_(hb2[j * ns * 150:(j + 1) * ns * 150], non_blocking=True)
                            ev_in[j].record(copy_stream)
                for j in range(NSUB):
                    if rank == 0:
                        cur.wait_event(ev_in[j])
                    if j == 0:
                        dist.broadcast(drel, src=0)
                    s1 = d1[j * ns * 150:(j + 1) * ns * 150 + (64 if j == NSUB - 1 else 0)]
                    s2 = d2[j * ns * 150:(j + 1) * ns * 150 + (64 if j == NSUB - 1 else 0)]
                    dist.broadcast(s1, src=0)
                    dist.broadcast(s2, src=0)
                    rc = L.mcg_query_batch_dev(db.h, C.byref(q), d1.data_ptr() + j * ns * 150, drel.data_ptr(), ns * 150,
                                               d2.data_ptr() + j * ns * 150, drel.data_ptr(), ns * 150, ns, 150,
                                               out_t + j * ns * Kc * 4, out_h + j * ns * Kc * 4, out_n + j * ns * 4)
                    if rc != 0:
                        raise RuntimeError(L.mcg_last_error().decode())
                if hdl_t is not None:
                    hdl_t.barrier(channel=0)
                    ct, ch, cn = gt, gh, gn
                else:
                    dist.all_gather_into_tensor(gt.view(world * npairs, Kc), lt)
                    dist.all_gather_into_tensor(gh.view(world * npairs, Kc), lh)
                    dist.all_gather_into_tensor(gn.view(world * npairs), ln)
                    ct, ch, cn = gt, gh, gn
                if rank == 0:
                    stats.zero_()
                    rc = L.mcg_merge_classify_dev(ctx.h, C.byref(q), ct.data_ptr(), ch.data_ptr(), cn.data_ptr(), world, npairs,
                                                  mt.data_ptr(), mh.data_ptr(), mn.data_ptr(), best.data_ptr(), brank.data_ptr(),
                                                  stats.data_ptr())
                    if rc != 0:
                        raise RuntimeError(L.mcg_last_error().decode())
                    hbest.copy_(best, non_blocking=True); hmt.copy_(mt, non_blocking=True); hmh.copy_(mh, non_blocking=True)
                if hdl_t is not None:
                    hdl_t.barrier(channel=1)
                torch.cuda.synchronize()
            st_before = stats.clone()
            qbcast()
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                qbcast()
            barrier()
            qb = max_over_ranks(time.perf_counter() - t0) / 3
            same_b = bool((stats == st_before).all().item()) if rank == 0 else None
            qe2e = {"value": 2 * npairs / qb, "unit": "reads/s", "ms_per_batch": 1e3 * qb,
                    "h2d_bytes_per_step": int(n1 + n2 + 16 * (npairs + 1)), "d2h_bytes_per_step": int(npairs * (8 * Kc + 4)),
                    "api": "rank 0: pinned host reads -> H2D once -> NCCL broadcast over NVLink -> mcg_query_batch_dev on every "
                           "shard -> top lists into rank 0 -> mcg_merge_classify_dev -> D2H of taxa and merged lists; "
                           "%d sub-batches, the PCIe copy of one overlaps the broadcast and query of the previous" % NSUB,
                    "same_statistics_as_device_path": same_b,
                    "every_rank_copies_all_reads_itself": qe2e}
        except Exception as e:      # pragma: no cover
            qe2e["broadcast_error"] = repr(e)
    # byte accounting per pair on this shard: 300 B reads + 32 B per probed feature + 8 B per location + 8K
    nfeat = 4 * 16
    mloc = None
    try:
        cnt = np.zeros(2000, np.int64)
        _, tgx, _ = db.query_matches(q, b1[:2000 * 150], o1[:2001].copy(), b2[:2000 * 150], o2[:2001].copy())
        mloc = len(tgx) / 2000.0
    except Exception:
        pass
    algq = (300 + 32 * nfeat + 8 * (mloc or 0) + 8 * Kc) * npairs
    out["query"] = {"metric": "query_reads_per_s", "value": 2 * npairs / (qms * 1e-3), "unit": "reads/s",
                    "pairs_per_s": npairs / (qms * 1e-3), "ms_per_batch": qms,
                    "workload": "%d pairs 2x150 bp per batch vs %d x 5 Mbp database sharded over %d GPU(s); -lowest species "
                                "-maxcand 4 -hitmin 4 -hitdiff 80; reads resident, per-rank top lists: %s, merge+classify on rank 0"
                                % (npairs, QG, world, gather_mode),
                    "classified": int(st[20]), "unclassified": int(st[21]),
                    "mean_locations_per_pair_shard0": mloc, "gpu_launches_per_batch": launches_q, "e2e": qe2e,
                    "roofline": {"bound": "hbm", "achieved": algq / (qms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": algq / (qms * 1e-3) / 1e9 / peak,
                                 "traffic": (load_traffic().get("query_fused_kernel", {}).get("dram_bytes_per_pair") or 0) * npairs or None,
                                 "note": "traffic: DRAM bytes of the fused kernel per batch from the committed ncu capture (1-GPU database); "
                                         "a probe fills a 128-byte line for 64 useful bytes"}}
    # long reads (BASELINE.json configs[4] shape): 10 kbp reads, ~89 windows / ~1400 features per read, staged path
    try:
        nlong = 20_000
        src = [synth.genome(g, GENOME_LEN) for g in range(0, QG, max(QG // 8, 1))]
        lb, lo, _ = synth.make_long_reads(src, nlong, readlen=10_000)
        del src
        dl = torch.from_numpy(np.concatenate([lb, np.zeros(64, np.uint8)])).to(dev)
        dlo = torch.from_numpy(lo).to(dev)
        llt = torch.empty((nlong, Kc), dtype=torch.int32, device=dev)
        llh = torch.empty((nlong, Kc), dtype=torch.int32, device=dev)
        lln = torch.empty(nlong, dtype=torch.int32, device=dev)

        def lstep():
            rc = L.mcg_query_batch_dev(db.h, C.byref(q), dl.data_ptr(), dlo.data_ptr(), len(lb), None, None, 0,
                                       nlong, 10_000, llt.data_ptr(), llh.data_ptr(), lln.data_ptr())
            if rc != 0:
                raise RuntimeError(L.mcg_last_error().decode())
        for _ in range(2):
            lstep()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(5):
            lstep()
        f1.record(stream)
        barrier()
        lms = max_over_ranks(f0.elapsed_time(f1)) / 5
        out["query_long"] = {"metric": "query_reads_per_s", "value": nlong / (lms * 1e-3), "unit": "reads/s",
                             "bp_per_s": nlong * 10_000 / (lms * 1e-3), "ms_per_batch": lms,
                             "workload": "%d single-end reads x 10 kbp (10 %% substitutions) per batch vs the shard of this rank "
                                         "(%d x 5 Mbp over %d GPU(s)); per-shard top lists only" % (nlong, QG, world),
                             "mean_candidates": float(lln.float().mean().item())}
    except Exception as e:
        out["query_long"] = {"error": repr(e)}
    db.close()
    # the reference's own build / query code (oracle/_ref, one MPI rank = one thread, as in the reference: build is
    # single-threaded per rank, querying.h runs T such workers per rank) on a bounded sample of the same data
    if world == 1 and not args.no_cpu:
        try:
            sys.path.insert(0, ROOT)
            from oracle import pyoracle as po
            if po.have_reference():
                ref = po.Reference()
                rdb = ref.db_create()
                rdb.set_taxonomy(ids, par, rk)
                ng = 4
                t0 = time.perf_counter()
                for g in range(ng):
                    rdb.add_target(genomes[g], "SYN%06d.1" % g, int(species[g]), 1, 0)
                tb = time.perf_counter() - t0
                out["build"]["cpu_baseline"] = {"value": ng * GENOME_LEN / tb, "unit": "bp/s", "cores": 1, "kind": "reference",
                                                "sample": "database::add_target_distributed on %d of the %d genomes (%d Mbp), one rank = "
                                                          "one thread, %.1f s; the reference scales with the number of MPI ranks"
                                                          % (ng, QG, ng * GENOME_LEN // 1_000_000, tb)}
                sub = [genomes[g] for g in range(ng)]
                sb1, so1, sb2, so2, _ = synth.make_read_pairs(sub, 20_000, seed=7)
                t0 = time.perf_counter()
                for i in range(20_000):
                    rdb.query_local(sb1[so1[i]:so1[i + 1]], sb2[so2[i]:so2[i + 1]], 4, 4)
                tq = time.perf_counter() - t0
                out["query"]["cpu_baseline"] = {"value": 2 * 20_000 / tq, "unit": "reads/s", "cores": 1, "kind": "reference",
                                                "sample": "accumulate_matches + candidate generation for 20000 pairs against the %d-genome "
                                                          "database above, one worker thread, %.1f s (ctypes call per read included)" % (ng, tq)}
        except Exception as e:      # pragma: no cover
            out["build"]["cpu_baseline"] = {"error": repr(e)}
    return out


def main():
    args = parse_args()
    # keep stdout for the ONE JSON line: anything libraries write to fd 1 meanwhile (e.g. NCCL's
    # version banner) goes to stderr
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(saved, "w")
    global print_json
    def print_json(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_mcgpu(args)


if __name__ == "__main__":
    main()
