#!/usr/bin/env python3
"""bench.py -- the measurement contract of this repository.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl mcgpu|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline = BASELINE.json's first metric, QUERY reads/s, on configs[3] (paired-end 2 x 150 bp reads against the
sharded database with the NCCL top-hit merge) over configs[2]'s database: 20 Gbp / 5000 genomes on 8 GPUs, i.e.
2.5 Gbp (625 genomes x 4 Mbp) per GPU.  With N GPUs the database is N x 2.5 Gbp (weak scaling in database size,
shard = target_id % N as in the reference); EVERY read is answered by EVERY shard, as in the reference.
A "step" is one batch of 1 M read pairs (2 M reads) through the whole path: K1 sketch of the reads -> [all-gather
of the sketches] -> table probe, gather, sort, window ranges, top list on every shard -> [all-to-all of the top
lists] -> merge in rank order + decision + statistics.  The default 10 steps are 20 % of configs[3]'s 50 M pairs.

  value : reads/s, whole job.  Every rank holds its 1/N slice of the batch resident in HBM; one step = one
          mcg_query_batch_sharded_dev call per rank; CUDA events on the library's stream, max over ranks.
  e2e   : the same through mcg_query_batch_sharded with pinned HOST slices: H2D of the reads and D2H of merged
          lists / taxa inside the timed region, every rank over its own PCIe link.
  roofline     : q_fused_warp_kernel (the dominant kernel) against the measured HBM copy bandwidth, its launch
                 durations taken live with CUDA events around every launch inside the timed region
                 (mcg_ctx_profile); algorithmic bytes per read pair = SURVEY 8(d): len1 + len2 + 32 F + 8 M + 8 K.
  cpu_baseline : the reference PROGRAM (oracle/_ref/metacache_mpi_ref, `mpirun -np 2 -threads cores/2` through the
                 FIFO shim) on a bounded sample: 32 genomes x 4 Mbp, 100 k pairs; its own `# time` line.
  parity       : after timing, 2000 read pairs of the timed batch are answered by the oracle over the SAME
                 database content (sketches of every genome of the shard, lists restricted to the sample's features)
                 and compared with what the GPUs returned: lists, taxa, decision.
  sketch / build / query_long : the other configs as extra objects with their own roofline / e2e / cpu_baseline.

--impl reference times the reference program on the host cores (rank 0 only) on a bounded sample of the same
workload and prints the same line with "impl": "reference".
"""
import argparse
import ctypes as C
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
print_json = None

K, S, WINLEN, STRIDE = 16, 16, 128, 113
GLEN = 4_000_000                                    # configs[2]: ~5 k genomes, ~20 Gbp
READLEN = 150
KC = 4                                              # -maxcand 4
QFLAGS = dict(maxcand=KC, lowest=4, hitmin=4, hitdiff=0.8)            # -lowest species -maxcand 4 -hitmin 4 -hitdiff 80
ALG_SKETCH = 1.0 + 4.0 * S / STRIDE                # 1 B in + 64 B per 113 bp out
FALLBACK_HBM_GBS = 6650.0                           # B200_PROFILING.md fallback
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "metacache_mpi_ref")
SHIM = os.path.join(ROOT, "oracle", "mpirun_shim.py")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mcgpu", choices=["mcgpu", "reference"])
    ap.add_argument("--db-mbp", type=int, default=2500, help="Mbp of reference genomes per GPU (default = configs[2] / 8)")
    ap.add_argument("--pairs", type=int, default=1_000_000, help="read pairs per step, whole job")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--no-extras", action="store_true", help="skip the sketch / build / long-read extra objects")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle sample check")
    ap.add_argument("--ref-genomes", type=int, default=32)
    ap.add_argument("--ref-pairs", type=int, default=100_000)
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
    """DRAM bytes of the dominant kernels from the committed ncu captures (profiles/traffic.json)"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))
    except Exception:
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


def workload_name(args, world):
    return ("query: batches of %d read pairs (2 x %d bp, %d reads) vs a %d x %.1f Gbp database (%d genomes x 4 Mbp, shard = "
            "target_id %% %d, k=16 sketchlen=16 winlen=128; BASELINE.json configs[3] reads over configs[2]'s database: "
            "20 Gbp / 8 per GPU); -lowest species -maxcand 4 -hitmin 4 -hitdiff 80"
            % (args.pairs, READLEN, 2 * args.pairs, world, args.db_mbp / 1000.0, world * (args.db_mbp * 1_000_000 // GLEN), world))


# ----------------------------------------------------------------------------- the reference program on the host cores
def reference_program_run(args, steps, warmup, ngenomes, npairs, log=None):
    """build (once) + `steps` query runs of the UNMODIFIED reference program under the MPI shim on a bounded sample
    of the workload; returns a dict with its own timers.  Needs oracle/_ref (prebuilt, travels with the repo)."""
    from __graft_entry__ import load_package
    load_package()
    from metacache_mpi_b200 import synth
    import torch
    if not os.path.exists(REF_EXE):
        raise RuntimeError("oracle/_ref/metacache_mpi_ref is not built")
    cores = os.cpu_count() or 1
    np_ = max(2, args.gpus)                     # the reference classifies nothing at -np 1
    T = max(1, cores // np_)
    gen_dev = "cuda" if torch.cuda.is_available() else "cpu"       # data generation only
    d = tempfile.mkdtemp(prefix="mcref_")
    try:
        genomes = torch.empty(ngenomes * GLEN, dtype=torch.uint8, device=gen_dev)
        for g in range(ngenomes):
            synth.genome_fast(g, GLEN, gen_dev, out=genomes[g * GLEN:(g + 1) * GLEN])
        ids, par, rk, species = synth.make_taxonomy(ngenomes)
        gh = genomes.cpu().numpy()
        synth.write_genome_files(os.path.join(d, "genomes"), [gh[g * GLEN:(g + 1) * GLEN] for g in range(ngenomes)], species)
        synth.write_taxonomy_files(os.path.join(d, "tax"), ids, par, rk)
        b1, b2, gsel = synth.read_pairs_fast(genomes, GLEN, npairs)
        offs = np.arange(npairs + 1, dtype=np.int64) * READLEN
        synth.write_fastq_pair(os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq"), b1.cpu().numpy(), offs, b2.cpu().numpy(), offs,
                               gsel.cpu().numpy())
        del genomes, b1, b2

        def run(cmd):
            r = subprocess.run(cmd, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
            if r.returncode != 0:
                raise RuntimeError("reference run failed: %s\n%s" % (" ".join(cmd), r.stdout.decode()[-2000:]))
            return r.stdout.decode()
        t0 = time.perf_counter()
        out = run([sys.executable, SHIM, "-np", str(np_), "--", REF_EXE, "build", "rdb", "genomes", "-taxonomy", "tax"])
        build_wall = time.perf_counter() - t0
        m = re.search(r"Added (\d+) reference sequences in ([0-9.eE+-]+) s", out)
        build_added_s = float(m.group(2)) if m else None
        qcmd = [sys.executable, SHIM, "-np", str(np_), "--", REF_EXE, "query", "rdb", "r1.fq", "r2.fq", "-pairfiles", "-lowest", "species",
                "-maxcand", "4", "-hitmin", "4", "-hitdiff", "80", "-threads", str(T), "-query-limit", str((npairs + T - 1) // T),
                "-out", "res.txt"]
        times, nq = [], 0
        for it in range(warmup + steps):
            run(qcmd)
            txt = open(os.path.join(d, "res.txt")).read()
            mt = re.search(r"# time:\s+(\d+) ms", txt)
            mq = re.search(r"# queries:\s+(\d+)", txt)
            if it >= warmup:
                times.append(float(mt.group(1)) * 1e-3)
                nq = int(mq.group(1))
            if log:
                log("reference query run %d: %s ms" % (it, mt.group(1)))
        cls = re.findall(r"#\s+(species|genus)\s+([0-9.]+)%", txt)
        return {"np": np_, "threads": T, "cores": cores, "query_s": times, "queries": nq, "build_added_s": build_added_s,
                "build_wall_s": build_wall, "build_bp": ngenomes * GLEN, "classified": dict(cls)}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def reference_sample_text(r, ngenomes, npairs):
    return ("the reference program (oracle/_ref/metacache_mpi_ref, unmodified sources, `mpirun -np %d` through the FIFO shim, "
            "-threads %d: %d of %d host threads) on a bounded sample of the workload: database %d genomes x 4 Mbp (%d Mbp), "
            "%d read pairs of the same generator per run; reads/s from its own '# time' line (rank 0, between barriers)"
            % (r["np"], r["threads"], r["np"] * r["threads"], r["cores"], ngenomes, ngenomes * GLEN // 1_000_000, npairs))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, min(args.warmup, 2)
    r = reference_program_run(args, steps, warmup, args.ref_genomes, args.ref_pairs)
    tot = float(np.sum(r["query_s"]))
    val = r["queries"] * len(r["query_s"]) / tot
    sample = reference_sample_text(r, args.ref_genomes, args.ref_pairs)
    line = {
        "impl": "reference", "metric": "query_reads_per_s", "value": val, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(r["query_s"]),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args, args.gpus), "sample": sample,
                   "note": "a smaller database means shorter location lists and FEWER chance hits per read: the sample favours the "
                           "reference; its sketching, table lookups and candidate generation per read are the same code"},
        "cpu_baseline": {"value": val, "unit": "reads/s", "cores": r["np"] * r["threads"], "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "build": {"metric": "build_bp_per_s", "value": r["build_bp"] / r["build_added_s"] if r["build_added_s"] else None, "unit": "bp/s",
                  "seconds": r["build_added_s"], "note": "the reference's own 'Added N reference sequences in X s' (sketch + insert), np=%d" % r["np"]},
        "gpu_launches": 0,
    }
    print_json(line)


# ----------------------------------------------------------------------------- our arm
def run_mcgpu(args):
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    load_package()
    from metacache_mpi_b200 import api, distributed as D, synth, taxonomy

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
    ctx = api.Context(local)                         # raises if libmcgpu.so / the GPU is missing; own (non-blocking) stream
    L = ctx.L
    comm = D.make_comm(ctx)                          # NCCL inside libmcgpu; torch.distributed only carries the id
    stream = torch.cuda.ExternalStream(ctx.L.mcg_ctx_stream(ctx.h), device=dev)
    t_start = time.perf_counter()

    def log(msg):
        if rank == 0 and os.environ.get("MCG_BENCH_VERBOSE"):
            sys.stderr.write("[bench %6.1f s] %s\n" % (time.perf_counter() - t_start, msg))
            sys.stderr.flush()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def check(rc):
        if rc != 0:
            raise RuntimeError(L.mcg_last_error().decode())

    peak, peak_src = hbm_peak()
    out_extra = {}

    # ---- database: this rank's shard of configs[2] (genomes g with g % world == rank), generated on the device
    Gper = max(args.db_mbp * 1_000_000 // GLEN, 1)
    Gtot = Gper * world
    own = np.arange(rank, Gtot, world, dtype=np.uint32)
    ids, par, rk, species = synth.make_taxonomy(Gtot)
    tabs = taxonomy.TaxonomyTables(ids, par, rk, species)
    ctx.upload_taxonomy(tabs.lineage, tabs.rank, Gtot)
    genomes = torch.empty(Gper * GLEN + 64, dtype=torch.uint8, device=dev)
    for i, g in enumerate(own):
        synth.genome_fast(int(g), GLEN, dev, out=genomes[i * GLEN:(i + 1) * GLEN])
    goffs = np.arange(Gper + 1, dtype=np.int64) * GLEN
    torch.cuda.synchronize()
    log("genomes generated on the device: %d x %d bp" % (Gper, GLEN))
    bp_shard = Gper * GLEN

    def build_shard(timed):
        db_ = api.Database(ctx, shard=rank, nshards=world)
        ctx.sync()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t0 = time.perf_counter()
        e0.record(stream)
        CH = 125                                    # genomes per add call (0.5 Gbp): the slot array grows in steps
        for a in range(0, Gper, CH):
            b = min(Gper, a + CH)
            check(L.mcg_db_add_targets_dev(db_.h, genomes.data_ptr(), goffs[a:b + 1].ctypes.data, own[a:b].ctypes.data, b - a, None))
        e1.record(stream)
        db_.finalize()
        e2.record(stream)
        ctx.sync()
        return db_, e0.elapsed_time(e1), e1.elapsed_time(e2), time.perf_counter() - t0

    db, _, _, _ = build_shard(False)                 # warm-up build (allocations)
    db.close()
    ctx.profile(True)
    ctx.kernel_times(0); ctx.kernel_times(2)
    barrier()
    db, ms_add, ms_fin, build_wall = build_shard(True)
    k1_ms, k1_n = ctx.kernel_times(0)
    rs_ms, rs_n = ctx.kernel_times(2)
    ctx.profile(False)
    nk, ne, nl = db.counts()
    build_ms = max_over_ranks(ms_add + ms_fin)
    log("shard built: %.1f ms sketch + %.1f ms finalize; %d keys, %d locations" % (ms_add, ms_fin, nk, nl))
    alg_build = 1.0 * bp_shard + 8.0 * nl + 12.0 * nk
    tr = load_traffic()
    out_extra["build"] = {
        "metric": "build_bp_per_s", "value": bp_shard * world / (build_ms * 1e-3), "unit": "bp/s", "ms": build_ms,
        "ms_sketch": ms_add, "ms_finalize": ms_fin,
        "workload": "database build, BASELINE.json configs[2] share: %d genomes x 4 Mbp (%.1f Gbp) per GPU x %d GPU(s), sequences resident "
                    "in HBM -> finalized HBM multimap (sketch, stable radix sort by feature, capped lists, bucket table); "
                    "no data-path collective" % (Gper, bp_shard / 1e9, world),
        "shard_keys": int(nk), "shard_locations": int(nl),
        "roofline": {"bound": "hbm", "achieved": alg_build / (ms_add + ms_fin) / 1e6, "peak": peak, "unit": "GB/s",
                     "frac": alg_build / (ms_add + ms_fin) / 1e6 / peak,
                     "algorithmic_bytes": "1 B/bp + 8 B/location + 12 B/key (SURVEY 8d) = %.2f B/bp" % (alg_build / bp_shard),
                     "traffic": (tr.get("build_kernels", {}).get("dram_bytes_per_bp") or 0) * bp_shard or None,
                     "kernels": {"sketch_tpw_kernel_ms": k1_ms, "rs_scatter_kernel_ms": rs_ms, "rs_scatter_launches": rs_n}},
    }

    # ---- reads: NB batches; this rank's slice of a batch = pairs sampled from its own genomes (every shard answers them)
    bnd = D.slice_bounds(args.pairs, world)
    n_loc = bnd[rank + 1] - bnd[rank]
    NB = 2
    gview = genomes[:Gper * GLEN]
    slices = []
    for b in range(NB):
        r1, r2, _ = synth.read_pairs_fast(gview, GLEN, n_loc, first_pair=b * args.pairs + bnd[rank], seed=7 + b)
        pad = torch.zeros(64, dtype=torch.uint8, device=dev)
        slices.append((torch.cat([r1, pad]), torch.cat([r2, pad])))
    offs_t = (torch.arange(n_loc + 1, dtype=torch.int64) * READLEN).pin_memory()       # pinned like the reads
    offs = offs_t.numpy()
    d_offs = offs_t.to(dev)
    nbytes = n_loc * READLEN
    q = api.make_query_opts(**QFLAGS)
    o_t = torch.empty((max(n_loc, 1), KC), dtype=torch.int32, device=dev)
    o_h = torch.empty((max(n_loc, 1), KC), dtype=torch.int32, device=dev)
    o_n = torch.empty(max(n_loc, 1), dtype=torch.int32, device=dev)
    o_b = torch.empty(max(n_loc, 1), dtype=torch.int32, device=dev)
    o_r = torch.empty(max(n_loc, 1), dtype=torch.int32, device=dev)
    d_stats = torch.zeros(22, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    log("read batches generated")

    def step_dev(i):
        s1, s2 = slices[i % NB]
        check(L.mcg_query_batch_sharded_dev(db.h, comm.h, C.byref(q), s1.data_ptr(), d_offs.data_ptr(), nbytes,
                                            s2.data_ptr(), d_offs.data_ptr(), nbytes, n_loc, READLEN,
                                            o_t.data_ptr(), o_h.data_ptr(), o_n.data_ptr(), o_b.data_ptr(), o_r.data_ptr(),
                                            d_stats.data_ptr()))

    sampler = ClockSampler(local)
    for i in range(args.warmup):
        step_dev(i)
    barrier()
    d_stats.zero_()
    ctx.profile(True)
    for w_ in range(8):
        ctx.kernel_times(w_)
    l0 = ctx.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    e0.record(stream)
    for i in range(args.steps):
        step_dev(i)
    e1.record(stream)
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launches() - l0
    fused_ms, fused_n = ctx.kernel_times(1)
    qk1_ms, qk1_n = ctx.kernel_times(0)
    phases = {name: ctx.kernel_times(w)[0] / max(args.steps, 1) for w, name in
              ((3, "bookkeeping_and_K1_of_the_slice"), (4, "exchange_of_sketches"), (5, "left_over_reads_staged_path"),
               (6, "exchange_of_top_lists"), (7, "merge_and_decision"))}
    phases["fused_kernel"] = fused_ms / max(args.steps, 1)
    ctx.profile(False)
    stats_dev = d_stats.cpu().numpy().astype(np.int64)
    value = 2.0 * args.pairs * args.steps / (dev_ms * 1e-3)
    log("device-resident: %.3f ms per step, fused kernel %.3f ms per launch" % (dev_ms / args.steps, fused_ms / max(fused_n, 1)))

    # ---- e2e: pinned host slices through the host entry point
    h1 = [s[0][:nbytes].cpu().pin_memory() for s in slices]
    h2 = [s[1][:nbytes].cpu().pin_memory() for s in slices]
    hot = torch.empty((max(n_loc, 1), KC), dtype=torch.int32).pin_memory()
    hoh = torch.empty((max(n_loc, 1), KC), dtype=torch.int32).pin_memory()
    hon = torch.empty(max(n_loc, 1), dtype=torch.int32).pin_memory()
    hob = torch.empty(max(n_loc, 1), dtype=torch.int32).pin_memory()
    hor = torch.empty(max(n_loc, 1), dtype=torch.int32).pin_memory()
    hstats = np.zeros(22, np.uint64)

    def step_e2e(i):
        check(L.mcg_query_batch_sharded(db.h, comm.h, C.byref(q), h1[i % NB].data_ptr(), offs.ctypes.data, h2[i % NB].data_ptr(),
                                        offs.ctypes.data, n_loc, hot.data_ptr(), hoh.data_ptr(), hon.data_ptr(), hob.data_ptr(),
                                        hor.data_ptr(), hstats.ctypes.data))

    for i in range(2):
        step_e2e(i)
    barrier()
    hstats[:] = 0
    e2e_prof = bool(os.environ.get("MCG_BENCH_E2E_PROFILE"))        # development aid: GPU time per phase inside the e2e calls
    if e2e_prof:
        ctx.profile(True)
        for w_ in range(8):
            ctx.kernel_times(w_)
    t0 = time.perf_counter()
    for i in range(args.e2e_steps):
        step_e2e(i)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    sampler.stop()
    if e2e_prof:
        log("e2e GPU ms per step by phase: %s" % {w_: round(ctx.kernel_times(w_)[0] / args.e2e_steps, 3) for w_ in range(8)})
        ctx.profile(False)
    e2e_value = 2.0 * args.pairs * args.e2e_steps / e2e_s
    log("e2e: %.3f ms per step" % (1e3 * e2e_s / args.e2e_steps))
    # what the e2e call returned for the batch it ran last ((e2e_steps - 1) % NB), kept for the parity sample
    last_b = (args.e2e_steps - 1) % NB
    e2e_res = (hot.numpy().copy(), hoh.numpy().view(np.uint32).copy(), hon.numpy().copy(), hob.numpy().copy(), hor.numpy().copy())
    # device-resident and host path agree on the statistics of the batches both ran
    same_stats = None
    if args.steps % NB == 0 and args.e2e_steps % NB == 0:
        same_stats = bool((stats_dev * args.e2e_steps == hstats.astype(np.int64) * args.steps).all())

    # ---- roofline of the dominant kernel: one launch handles all 2 M reads of a step against this shard
    mloc = None
    try:
        ns = min(2000, n_loc)
        _, tgx, _ = db.query_matches(q, h1[0].numpy()[:ns * READLEN], offs[:ns + 1].copy(), h2[0].numpy()[:ns * READLEN], offs[:ns + 1].copy())
        mloc = len(tgx) / float(ns)
    except Exception:
        pass
    F = 4 * S
    alg_pair = 2 * READLEN + 32 * F + 8 * (mloc or 0) + 8 * KC
    fused_per_launch = fused_ms / max(fused_n, 1)
    achieved = alg_pair * args.pairs / (fused_per_launch * 1e-3) / 1e9 if fused_n else None
    trq = tr.get("query_fused_kernel", {})
    line = {
        "metric": "query_reads_per_s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": workload_name(args, world),
                   "fraction_of_configs3": "%d steps x %d pairs = %.0f %% of configs[3]'s 50 M pairs" % (args.steps, args.pairs, 100.0 * args.steps * args.pairs / 50e6),
                   "scaling_note": "weak in DATABASE size (2.5 Gbp per GPU: N = 8 is configs[2]'s 20 Gbp); every read is answered by every shard as "
                                   "in the reference, so reads/s x database size is the quantity that grows with N",
                   "l2": "every step probes a %.1f GB bucket table at random and reads a fresh 300 MB batch (%d batches alternate): far beyond the 126 MB L2, no flush needed"
                         % (nk / 0.45 / 3 * 128 / 1e9, NB),
                   "parallelism": "one process per GPU; shard = target_id %% %d; per step: slice ingest, K1 on the slice, NCCL all-gather of sketches, "
                                  "shard query of all reads, NCCL all-to-all of top lists, slice-wise merge + decision (libmcgpu: mcg_query_batch_sharded)" % world,
                   "self_check": "device-resident and host-path runs give the same 22 statistics counters: %s" % same_stats},
        "classified": int(stats_dev[20] // max(args.steps, 1)) if world == 1 else None,
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": "reads/s", "ms_per_step": 1e3 * e2e_s / args.e2e_steps,
                "h2d_bytes_per_step": int(2 * nbytes + 16 * (n_loc + 1)) * world, "d2h_bytes_per_step": int(n_loc * (8 * KC + 12) + 176) * world,
                "steps": args.e2e_steps, "api": "mcg_query_batch_sharded (pinned host slices; every rank copies its slice over its own PCIe link)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                     "traffic": (trq.get("dram_bytes_per_pair") or 0) * args.pairs or None, "peak_source": peak_src,
                     "kernel": "mcg::q_fused_warp_kernel", "ms_per_launch": fused_per_launch, "launches_timed": int(fused_n),
                     "share_of_step": fused_ms / (dev_ms if world == 1 else e0.elapsed_time(e1)),
                     "algorithmic_bytes_per_pair": alg_pair, "mean_locations_per_pair": mloc,
                     "formula": "len1 + len2 + 32 B x F (F = 64 feature probes) + 8 B x M (locations gathered) + 8 B x K (SURVEY 8d), x %d pairs per launch" % args.pairs,
                     "note": "random 128-byte line fills (one per probe) + dependent location gathers: the kernel is latency / issue bound, see profiles/"},
        "sketch_in_query": {"ms_per_step": qk1_ms / max(args.steps, 1), "launches": int(qk1_n)},
        "phases_ms_per_step_rank0": phases,
    }
    line.update(out_extra)

    # ---- parity sample: oracle over the same database content
    if not args.no_check:
        try:
            line["parity"] = parity_sample(args, api, synth, tabs, db, comm, q, world, rank, dist, torch, dev, genomes, goffs, own, Gper, Gtot,
                                           h1[last_b], h2[last_b], offs, n_loc, e2e_res, log)
        except Exception as e:
            line["parity"] = {"error": repr(e)}

    # ---- extras
    if not args.no_extras:
        try:
            line.update(run_extras(args, ctx, api, synth, torch, dist, dev, stream, world, rank, barrier, max_over_ranks, check, peak, tr,
                                   db, comm, q, genomes, goffs, own, Gper, log))
        except Exception as e:       # extras must never hide the headline
            line["extras_error"] = repr(e)

    # ---- CPU baseline: the reference program on the host cores (rank 0, N = 1 only)
    if world == 1 and not args.no_cpu:
        try:
            del slices
            torch.cuda.empty_cache()
            r = reference_program_run(args, 1, 1, args.ref_genomes, args.ref_pairs, log)
            v = r["queries"] / r["query_s"][0]
            sample = reference_sample_text(r, args.ref_genomes, args.ref_pairs)
            line["cpu_baseline"] = {"value": v, "unit": "reads/s", "cores": r["np"] * r["threads"], "kind": "reference", "sample": sample}
            if "build" in line and r["build_added_s"]:
                line["build"]["cpu_baseline"] = {"value": r["build_bp"] / r["build_added_s"], "unit": "bp/s", "cores": r["np"], "kind": "reference",
                                                 "sample": "the reference program's build of that sample database (%d Mbp), its own 'Added N reference "
                                                           "sequences in X s' timer; build is single-threaded per MPI rank, np=%d"
                                                           % (r["build_bp"] // 1_000_000, r["np"])}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}

    if rank == 0:
        print_json(line)
    db.close()
    comm.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def parity_sample(args, api, synth, tabs, db, comm, q, world, rank, dist, torch, dev, genomes, goffs, own, Gper, Gtot,
                  hb1, hb2, offs, n_loc, e2e_res, log):
    """2000 pairs of rank 0's slice of the timed batch: per-shard lists of every GPU and the merged result of the timed
    e2e call against the oracle.  The oracle sketches EVERY genome of the shard (all host threads) and keeps the lists
    of the sample's features: lookups of those features see exactly the lists of the whole database."""
    from oracle import pyoracle as po
    O = po.Oracle()
    NS = 2000
    ns = min(NS, n_loc) if rank == 0 else 0
    # rank 0's sample reads go to every rank
    if world > 1:
        cnt = torch.tensor([ns], dtype=torch.int64, device=dev)
        dist.broadcast(cnt, 0)
        ns = int(cnt.item())
        s1 = torch.empty(ns * READLEN, dtype=torch.uint8, device=dev)
        s2 = torch.empty(ns * READLEN, dtype=torch.uint8, device=dev)
        if rank == 0:
            s1.copy_(hb1[:ns * READLEN]); s2.copy_(hb2[:ns * READLEN])
        dist.broadcast(s1, 0); dist.broadcast(s2, 0)
        b1, b2 = s1.cpu().numpy(), s2.cpu().numpy()
    else:
        b1, b2 = hb1.numpy()[:ns * READLEN].copy(), hb2.numpy()[:ns * READLEN].copy()
    so = offs[:ns + 1].copy()
    f1, _ = O.sketch_sequences(b1, so)
    f2, _ = O.sketch_sequences(b2, so)
    allowed = np.unique(np.concatenate([f1.ravel(), f2.ravel()]))
    allowed = allowed[allowed != 0xFFFFFFFF]
    threads = max(1, (os.cpu_count() or 1) // world)
    t0 = time.perf_counter()
    odb = O.db_create(shard=rank, nshards=world)
    gh = genomes[:Gper * GLEN].cpu().numpy()
    odb.add_targets_restricted(gh, goffs, own, allowed, threads=threads)
    odb.finalize()
    del gh
    t_oracle = time.perf_counter() - t0
    tax = po.TaxTables(tabs.lineage, tabs.rank, Gtot)
    oq = po.make_qopt(**QFLAGS)
    # (1) per-shard lists of this GPU (mcg_query_batch) == oracle shard
    gt, gh_, gn = db.query_batch(q, b1, so, b2, so)
    lt = np.full((ns, KC), -1, np.int32); lh = np.zeros((ns, KC), np.uint32); ln = np.zeros(ns, np.int32)
    for i in range(ns):
        t, h = odb.query_local(tax, oq, b1[so[i]:so[i + 1]], b2[so[i]:so[i + 1]])
        ln[i] = len(t); lt[i, :len(t)] = t; lh[i, :len(t)] = h
    shard_ok = bool((gn == ln).all() and (gt == lt).all() and (gh_ == lh).all())
    res = {"sample_pairs": ns, "oracle_threads": threads, "oracle_build_s": t_oracle, "features_in_sample": int(len(allowed))}
    if world > 1:
        allv = [None] * world if rank == 0 else None
        dist.gather_object((shard_ok, lt, lh, ln), allv, dst=0)
    else:
        allv = [(shard_ok, lt, lh, ln)]
    if rank != 0:
        return None
    res["per_shard_lists_equal_oracle"] = [bool(v[0]) for v in allv]
    # (2) merged lists + decision of the timed e2e call == oracle merge of the oracle's per-shard lists (rank order)
    mt, mh, mn, mb, mr = e2e_res
    ok = True
    nclass = 0
    for i in range(ns):
        it, ih = [], []
        for r in range(world):
            c = int(allv[r][3][i])
            it += allv[r][1][i, :c].tolist(); ih += allv[r][2][i, :c].tolist()
        et, eh, eb, er = O.merge_classify(tax, oq, it, ih)
        k = len(et)
        ok &= bool(mn[i] == k and (mt[i, :k] == et).all() and (mh[i, :k] == eh).all() and mb[i] == eb and mr[i] == er)
        nclass += eb >= 0
    res["merged_lists_and_taxa_equal_oracle"] = bool(ok)
    res["classified_in_sample"] = int(nclass)
    res["verdict"] = "bit-exact" if ok and all(res["per_shard_lists_equal_oracle"]) else "MISMATCH"
    return res


def run_extras(args, ctx, api, synth, torch, dist, dev, stream, world, rank, barrier, max_over_ranks, check, peak, tr,
               db, comm, q, genomes, goffs, own, Gper, log):
    """configs[1] (sketch-only 1 Gbp per GPU) and configs[4] (10 kbp reads) on the data already resident"""
    L = ctx.L
    out = {}
    # ---- sketching only: 1 Gbp of this rank's genomes (BASELINE.json configs[1])
    G1 = min(Gper, 250)
    bp = G1 * GLEN
    cfg = api.SketchCfg(K, S, WINLEN, STRIDE)
    o1g = goffs[:G1 + 1].copy()
    nwin = ctx.count_windows(cfg, o1g)
    d_off = torch.from_numpy(o1g).to(dev)
    d_feat = torch.empty(nwin * S, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    def sk():
        check(L.mcg_sketch_windows_dev(ctx.h, C.byref(cfg), genomes.data_ptr(), bp, d_off.data_ptr(), G1, nwin, d_feat.data_ptr(), None))
    for _ in range(3):
        sk()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ST = 20
    e0.record(stream)
    for _ in range(ST):
        sk()
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / ST
    hseq = genomes[:bp].cpu().pin_memory()
    hfeat = torch.empty(nwin * S, dtype=torch.int32).pin_memory()
    nw_out = C.c_int64()

    def ske():
        check(L.mcg_sketch_windows(ctx.h, C.byref(cfg), hseq.data_ptr(), o1g.ctypes.data, G1, hfeat.data_ptr(), None, C.byref(nw_out)))
    ske()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        ske()
    barrier()
    es = max_over_ranks(time.perf_counter() - t0) / 3
    same = bool((d_feat.cpu() == hfeat).all().item())
    ach = ALG_SKETCH * bp / (ms * 1e-3) / 1e9
    out["sketch"] = {"metric": "sketch_bp_per_s", "value": bp * world / (ms * 1e-3), "unit": "bp/s", "ms_per_step": ms,
                     "workload": "sketching only: %d Mbp of reference genomes per GPU (%d x 4 Mbp), k=16 sketchlen=16 winlen=128 winstride=113 "
                                 "(BASELINE.json configs[1]); weak scaling, no collective" % (bp // 1_000_000, G1),
                     "e2e": {"value": bp * world / es, "unit": "bp/s", "h2d_bytes_per_step": int(bp + 8 * (G1 + 1)) * world,
                             "d2h_bytes_per_step": int(nwin * S * 4) * world, "api": "mcg_sketch_windows (pinned host buffers)",
                             "same_as_device_path": same},
                     "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                  "traffic": (tr.get("sketch_kernel", {}).get("dram_bytes_per_bp") or 0) * bp or None,
                                  "kernel": "mcg::sketch_tpw_kernel", "algorithmic_bytes_per_bp": ALG_SKETCH,
                                  "note": "instruction-throughput bound (ALU pipe), see profiles/"}}
    if world == 1 and not args.no_cpu:
        try:
            from oracle import pyoracle as po
            cores = os.cpu_count() or 1
            impl, kind = (po.Reference(), "reference") if po.have_reference() else (po.Oracle(), "port")
            sg = int(max(2, min(G1, 10 * cores)))
            t, cs, nw = impl.sketch_throughput(hseq.numpy()[:sg * GLEN], o1g[:sg + 1].copy(), K, S, WINLEN, STRIDE, threads=cores)
            out["sketch"]["cpu_baseline"] = {"value": sg * GLEN / t, "unit": "bp/s", "cores": cores, "kind": kind,
                                             "sample": "the reference's hash_dna.h sketcher on %d of the genomes (%d Mbp), %d threads, %.1f s" % (sg, sg * 4, cores, t)}
        except Exception as e:
            out["sketch"]["cpu_baseline"] = {"error": repr(e)}
    del hseq, hfeat, d_feat
    log("sketch extra done")

    # ---- long reads: BASELINE.json configs[4] shape, 10 kbp single-end reads, many windows per read
    try:
        nlong_tot = 100_000
        bnd = [nlong_tot * r // world for r in range(world + 1)]
        nl_ = bnd[rank + 1] - bnd[rank]
        RL = 10_000
        lr = synth.long_reads_fast(genomes[:Gper * GLEN], GLEN, nl_, first_read=bnd[rank])
        lr = torch.cat([lr, torch.zeros(64, dtype=torch.uint8, device=dev)])
        lo = torch.from_numpy(np.arange(nl_ + 1, dtype=np.int64) * RL).to(dev)
        t_ = torch.empty((max(nl_, 1), KC), dtype=torch.int32, device=dev); h_ = torch.empty((max(nl_, 1), KC), dtype=torch.int32, device=dev)
        n_ = torch.empty(max(nl_, 1), dtype=torch.int32, device=dev); b_ = torch.empty(max(nl_, 1), dtype=torch.int32, device=dev)
        r_ = torch.empty(max(nl_, 1), dtype=torch.int32, device=dev); st_ = torch.zeros(22, dtype=torch.int64, device=dev)

        def lstep():
            check(L.mcg_query_batch_sharded_dev(db.h, comm.h, C.byref(q), lr.data_ptr(), lo.data_ptr(), nl_ * RL, None, None, 0, nl_, RL,
                                                t_.data_ptr(), h_.data_ptr(), n_.data_ptr(), b_.data_ptr(), r_.data_ptr(), st_.data_ptr()))
        lstep()
        barrier()
        st_.zero_()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        LS = 3
        f0.record(stream)
        for _ in range(LS):
            lstep()
        f1.record(stream)
        barrier()
        lms = max_over_ranks(f0.elapsed_time(f1)) / LS
        stl = st_.cpu().numpy()
        out["query_long"] = {"metric": "query_reads_per_s", "value": nlong_tot / (lms * 1e-3), "unit": "reads/s",
                             "bp_per_s": nlong_tot * RL / (lms * 1e-3), "ms_per_step": lms,
                             "workload": "long reads, BASELINE.json configs[4] shape: %d single-end reads x 10 kbp (10 %% substitutions, ~89 windows / "
                                         "~1400 features per read) per step = 10 %% of configs[4]'s 1 M reads, vs the same %d-shard database; staged path "
                                         "(find, scan, gather, size-binned per-read sorters)" % (nlong_tot, world),
                             "classified_per_step_this_rank": int(stl[20] // LS)}
    except Exception as e:
        out["query_long"] = {"error": repr(e)}
    log("long reads extra done")
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
