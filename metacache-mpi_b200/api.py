"""ctypes binding of the C-ABI (include/mcgpu.h) plus a thin host-side mirror of
the reference seams (sketcher / database / query) used by tests and bench.py.

There is no CPU fallback: if libmcgpu.so is missing or no B200 is visible every
compute call raises.
"""
import ctypes as C
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libmcgpu.so")
HEADER = os.path.join(ROOT, "include", "mcgpu.h")

NO_FEATURE = 0xFFFFFFFF
NUM_RANKS = 21
RANK_NONE = 21

_lib = None


class McgError(RuntimeError):
    pass


class SketchCfg(C.Structure):
    _fields_ = [("k", C.c_int32), ("sketchlen", C.c_int32), ("winlen", C.c_int64), ("winstride", C.c_int64)]


class QueryOpts(C.Structure):
    _fields_ = [("sketchlen", C.c_int32), ("winlen", C.c_int64), ("winstride", C.c_int64),
                ("maxcand", C.c_int32), ("lowest", C.c_int32), ("highest", C.c_int32),
                ("hitmin", C.c_int32), ("hitdiff", C.c_float), ("insertsize", C.c_int64),
                ("reference_compat", C.c_int32)]


def declared_symbols():
    """every function name include/mcgpu.h declares"""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mcg_[a-z0-9_]+)\s*\(", txt)))


def load_library():
    """dlopen libmcgpu.so and check that it exports every declared symbol"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise McgError("libmcgpu.so not built (%s): run `python __graft_entry__.py`; there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    if missing:
        raise McgError("libmcgpu.so lacks symbols declared in include/mcgpu.h: %s" % ", ".join(missing))
    vp, i64, i32, u64p = C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_uint64)
    L.mcg_last_error.restype = C.c_char_p
    L.mcg_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.mcg_ctx_destroy.argtypes = [vp]
    L.mcg_ctx_sync.argtypes = [vp]
    L.mcg_ctx_stream.restype = vp
    L.mcg_ctx_stream.argtypes = [vp]
    L.mcg_ctx_launches.restype = i64
    L.mcg_ctx_launches.argtypes = [vp]
    L.mcg_ctx_trim.argtypes = [vp]
    L.mcg_ctx_profile.argtypes = [vp, C.c_int]
    L.mcg_ctx_kernel_times.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(i64)]
    L.mcg_db_add_targets_dev.argtypes = [vp, vp, vp, vp, i64, vp]
    L.mcg_dev_alloc.argtypes = [vp, i64, C.POINTER(vp)]
    L.mcg_dev_free.argtypes = [vp, vp]
    L.mcg_host_alloc.argtypes = [i64, C.POINTER(vp)]
    L.mcg_host_free.argtypes = [vp]
    L.mcg_memcpy_h2d.argtypes = [vp, vp, vp, i64]
    L.mcg_memcpy_d2h.argtypes = [vp, vp, vp, i64]
    L.mcg_count_windows.restype = i64
    L.mcg_count_windows.argtypes = [C.POINTER(SketchCfg), vp, i64]
    L.mcg_sketch_windows.argtypes = [vp, C.POINTER(SketchCfg), vp, vp, i64, vp, vp, C.POINTER(i64)]
    L.mcg_sketch_windows_dev.argtypes = [vp, C.POINTER(SketchCfg), vp, i64, vp, i64, i64, vp, vp]
    L.mcg_db_create.argtypes = [vp, C.POINTER(SketchCfg), i32, i32, i32, C.POINTER(vp)]
    L.mcg_db_destroy.argtypes = [vp]
    L.mcg_db_add_targets.argtypes = [vp, vp, vp, vp, i64, vp]
    L.mcg_db_finalize.argtypes = [vp]
    L.mcg_db_counts.argtypes = [vp, u64p, u64p, u64p]
    L.mcg_db_export.argtypes = [vp, vp, vp, vp, vp]
    L.mcg_db_import.argtypes = [vp, vp, vp, vp, vp, C.c_uint64, C.c_uint64]
    L.mcg_db_key_sizes.argtypes = [vp, vp, vp]
    L.mcg_db_remove_keys.argtypes = [vp, vp, C.c_uint64, u64p]
    L.mcg_remove_overpopulated.argtypes = [C.POINTER(vp), i32, u64p]
    L.mcg_tax_upload.argtypes = [vp, vp, vp, i64, i64]
    L.mcg_query_batch.argtypes = [vp, C.POINTER(QueryOpts), vp, vp, vp, vp, i64, vp, vp, vp]
    L.mcg_query_batch_dev.argtypes = [vp, C.POINTER(QueryOpts), vp, vp, i64, vp, vp, i64, i64, i64, vp, vp, vp]
    L.mcg_query_matches.argtypes = [vp, C.POINTER(QueryOpts), vp, vp, vp, vp, i64, vp, vp, vp]
    L.mcg_merge_classify.argtypes = [vp, C.POINTER(QueryOpts), vp, vp, vp, i32, i64, vp, vp, vp, vp, vp, vp]
    L.mcg_merge_classify_dev.argtypes = [vp, C.POINTER(QueryOpts), vp, vp, vp, i32, i64, vp, vp, vp, vp, vp, vp]
    L.mcg_db_key_sizes_dev.argtypes = [vp, vp, vp]
    L.mcg_comm_get_id.argtypes = [vp]
    L.mcg_comm_create.argtypes = [vp, vp, i32, i32, C.POINTER(vp)]
    L.mcg_comm_destroy.argtypes = [vp]
    L.mcg_comm_rank.argtypes = [vp]
    L.mcg_comm_size.argtypes = [vp]
    L.mcg_db_remove_overpopulated_dist.argtypes = [vp, vp, u64p, u64p]
    L.mcg_query_batch_sharded.argtypes = [vp, vp, C.POINTER(QueryOpts), vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp]
    L.mcg_query_batch_sharded_dev.argtypes = [vp, vp, C.POINTER(QueryOpts), vp, vp, i64, vp, vp, i64, i64, i64,
                                              vp, vp, vp, vp, vp, vp]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise McgError("libmcgpu error %d: %s" % (rc, load_library().mcg_last_error().decode()))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u8(a):
    if isinstance(a, (bytes, bytearray)):
        a = np.frombuffer(a, dtype=np.uint8)
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a if a.size else np.zeros(1, np.uint8)


def make_query_opts(sketchlen=16, winlen=128, winstride=113, maxcand=2, lowest=0, highest=19,
                    hitmin=5, hitdiff=1.0, insertsize=0, reference_compat=0):
    return QueryOpts(sketchlen, winlen, winstride, maxcand, lowest, highest, hitmin,
                     float(np.float32(hitdiff)), insertsize, reference_compat)


def default_hitmin(sketchlen):
    """mode_query.cpp:247-260"""
    if sketchlen >= 6:
        return int(sketchlen / 3.0)
    return 2 if sketchlen >= 4 else 1


class Context:
    """one GPU: device + stream + scratch (mcg_ctx)"""

    def __init__(self, device=0, stream=None):
        L = load_library()
        self.L = L
        h = C.c_void_p()
        _check(L.mcg_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h)))
        self.h = h
        self.device = device
        self._dbs = []

    def close(self):
        if self.h:
            for d in list(self._dbs):      # shards hold a pointer to the ctx: destroy them first
                d.close()
            self.L.mcg_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _check(self.L.mcg_ctx_sync(self.h))

    def launches(self):
        return self.L.mcg_ctx_launches(self.h)

    def profile(self, on=True):
        _check(self.L.mcg_ctx_profile(self.h, 1 if on else 0))

    def kernel_times(self, which):
        """(summed device ms, launches) of kernel class `which` (0 sketch, 1 fused query, 2 radix scatter) since the last call"""
        ms, n = C.c_double(), C.c_int64()
        _check(self.L.mcg_ctx_kernel_times(self.h, which, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # ---- sketcher::operator() over all windows (hash_dna.h:115)
    def count_windows(self, cfg, offs):
        offs = np.ascontiguousarray(offs, np.int64)
        return self.L.mcg_count_windows(C.byref(cfg), _ptr(offs), len(offs) - 1)

    def sketch_windows(self, buf, offs, k=16, sketchlen=16, winlen=128, winstride=113, counts=True):
        cfg = SketchCfg(k, sketchlen, winlen, winstride)
        buf = _u8(buf)
        offs = np.ascontiguousarray(offs, np.int64)
        nseq = len(offs) - 1
        nwin = self.count_windows(cfg, offs)
        if nwin < 0:
            raise McgError("bad window parameters")
        feat = np.empty((max(nwin, 1), sketchlen), np.uint32)
        cnt = np.empty(max(nwin, 1), np.uint32) if counts else None
        n = C.c_int64()
        _check(self.L.mcg_sketch_windows(self.h, C.byref(cfg), _ptr(buf), _ptr(offs), nseq, _ptr(feat), _ptr(cnt), C.byref(n)))
        assert n.value == nwin
        return feat[:nwin], (cnt[:nwin] if counts else None)

    def upload_taxonomy(self, lineage, rank, ntargets):
        lineage = np.ascontiguousarray(lineage, np.int32)
        rank = np.ascontiguousarray(rank, np.uint8)
        _check(self.L.mcg_tax_upload(self.h, _ptr(lineage), _ptr(rank), len(rank), ntargets))

    def merge_classify(self, qopt, cand_tax, cand_hits, cand_n):
        """cand_*: [P][n][K] / [P][n] in rank order"""
        cand_tax = np.ascontiguousarray(cand_tax, np.int32)
        cand_hits = np.ascontiguousarray(cand_hits, np.uint32)
        cand_n = np.ascontiguousarray(cand_n, np.int32)
        P, n, K = cand_tax.shape
        ot = np.empty((max(n, 1), K), np.int32)
        oh = np.empty((max(n, 1), K), np.uint32)
        on = np.empty(max(n, 1), np.int32)
        best = np.empty(max(n, 1), np.int32)
        brank = np.empty(max(n, 1), np.int32)
        stats = np.zeros(22, np.uint64)
        _check(self.L.mcg_merge_classify(self.h, C.byref(qopt), _ptr(cand_tax), _ptr(cand_hits), _ptr(cand_n), P, n,
                                         _ptr(ot), _ptr(oh), _ptr(on), _ptr(best), _ptr(brank), _ptr(stats)))
        return ot[:n], oh[:n], on[:n], best[:n], brank[:n], stats


class Comm:
    """one rank of the NCCL communicator inside libmcgpu (mcg_comm).  `comm_id`: the 128 bytes of
    mcg_comm_get_id() from rank 0, carried to the other ranks by the launcher (see distributed.make_comm)"""

    def __init__(self, ctx, comm_id, rank, nranks):
        self.ctx, self.L = ctx, ctx.L
        h = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(comm_id)) if nranks > 1 else None
        _check(self.L.mcg_comm_create(ctx.h, buf, rank, nranks, C.byref(h)))
        self.h = h
        self.rank, self.nranks = rank, nranks

    @staticmethod
    def new_id():
        buf = (C.c_uint8 * 128)()
        _check(load_library().mcg_comm_get_id(buf))
        return bytes(buf)

    def close(self):
        if self.h:
            self.L.mcg_comm_destroy(self.h)
            self.h = None


class Database:
    """one shard (mcg_db): the sketch_database seams of the reference"""

    def __init__(self, ctx, k=16, sketchlen=16, winlen=128, winstride=113, maxlocs=254, shard=0, nshards=1):
        self.ctx = ctx
        self.L = ctx.L
        self.cfg = SketchCfg(k, sketchlen, winlen, winstride)
        h = C.c_void_p()
        _check(self.L.mcg_db_create(ctx.h, C.byref(self.cfg), maxlocs, shard, nshards, C.byref(h)))
        self.h = h
        self.shard, self.nshards = shard, nshards
        ctx._dbs.append(self)

    def close(self):
        if self.h:
            if self.ctx.h:
                self.L.mcg_db_destroy(self.h)
            self.h = None
            if self in self.ctx._dbs:
                self.ctx._dbs.remove(self)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_targets(self, buf, offs, target_ids):
        buf = _u8(buf)
        offs = np.ascontiguousarray(offs, np.int64)
        ids = np.ascontiguousarray(target_ids, np.uint32)
        n = len(ids)
        wins = np.zeros(max(n, 1), np.uint32)
        _check(self.L.mcg_db_add_targets(self.h, _ptr(buf), _ptr(offs), _ptr(ids), n, _ptr(wins)))
        return wins[:n]

    def finalize(self):
        _check(self.L.mcg_db_finalize(self.h))

    def counts(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(self.L.mcg_db_counts(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def export(self):
        _, nk, nl = self.counts()
        keys = np.empty(max(nk, 1), np.uint32)
        sizes = np.empty(max(nk, 1), np.uint32)
        tg = np.empty(max(nl, 1), np.uint32)
        wn = np.empty(max(nl, 1), np.uint32)
        _check(self.L.mcg_db_export(self.h, _ptr(keys), _ptr(sizes), _ptr(tg), _ptr(wn)))
        return keys[:nk], sizes[:nk], tg[:nl], wn[:nl]

    def import_(self, keys, sizes, tgts, wins):
        keys = np.ascontiguousarray(keys, np.uint32)
        sizes = np.ascontiguousarray(sizes, np.uint32)
        tgts = np.ascontiguousarray(tgts, np.uint32)
        wins = np.ascontiguousarray(wins, np.uint32)
        _check(self.L.mcg_db_import(self.h, _ptr(keys), _ptr(sizes), _ptr(tgts), _ptr(wins), len(keys), len(tgts)))

    def key_sizes(self):
        _, nk, _ = self.counts()
        keys = np.empty(max(nk, 1), np.uint32)
        sizes = np.empty(max(nk, 1), np.uint32)
        _check(self.L.mcg_db_key_sizes(self.h, _ptr(keys), _ptr(sizes)))
        return keys[:nk], sizes[:nk]

    def remove_keys(self, keys):
        keys = np.ascontiguousarray(keys, np.uint32)
        rem = C.c_uint64()
        _check(self.L.mcg_db_remove_keys(self.h, _ptr(keys), len(keys), C.byref(rem)))
        return rem.value

    def query_batch(self, qopt, buf1, offs1, buf2=None, offs2=None):
        buf1 = _u8(buf1)
        offs1 = np.ascontiguousarray(offs1, np.int64)
        n = len(offs1) - 1
        if buf2 is not None:
            buf2 = _u8(buf2)
            offs2 = np.ascontiguousarray(offs2, np.int64)
        K = qopt.maxcand
        ot = np.empty((max(n, 1), K), np.int32)
        oh = np.empty((max(n, 1), K), np.uint32)
        on = np.empty(max(n, 1), np.int32)
        _check(self.L.mcg_query_batch(self.h, C.byref(qopt), _ptr(buf1), _ptr(offs1), _ptr(buf2), _ptr(offs2), n,
                                      _ptr(ot), _ptr(oh), _ptr(on)))
        return ot[:n], oh[:n], on[:n]

    def query_batch_sharded(self, comm, qopt, buf1, offs1, buf2=None, offs2=None):
        """collective: this rank's slice of the batch in, merged lists + decision + statistics of the slice out"""
        buf1 = _u8(buf1)
        offs1 = np.ascontiguousarray(offs1, np.int64)
        n = len(offs1) - 1
        if buf2 is not None:
            buf2 = _u8(buf2)
            offs2 = np.ascontiguousarray(offs2, np.int64)
        K = qopt.maxcand
        ot = np.empty((max(n, 1), K), np.int32)
        oh = np.empty((max(n, 1), K), np.uint32)
        on = np.empty(max(n, 1), np.int32)
        best = np.empty(max(n, 1), np.int32)
        brank = np.empty(max(n, 1), np.int32)
        stats = np.zeros(22, np.uint64)
        _check(self.L.mcg_query_batch_sharded(self.h, comm.h if comm is not None else None, C.byref(qopt), _ptr(buf1), _ptr(offs1),
                                              _ptr(buf2), _ptr(offs2), n, _ptr(ot), _ptr(oh), _ptr(on), _ptr(best), _ptr(brank),
                                              _ptr(stats)))
        return ot[:n], oh[:n], on[:n], best[:n], brank[:n], stats

    def remove_overpopulated_dist(self, comm):
        a, b = C.c_uint64(), C.c_uint64()
        _check(self.L.mcg_db_remove_overpopulated_dist(self.h, comm.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def query_matches(self, qopt, buf1, offs1, buf2=None, offs2=None):
        buf1 = _u8(buf1)
        offs1 = np.ascontiguousarray(offs1, np.int64)
        n = len(offs1) - 1
        if buf2 is not None:
            buf2 = _u8(buf2)
            offs2 = np.ascontiguousarray(offs2, np.int64)
        cnt = np.zeros(max(n, 1), np.int64)
        _check(self.L.mcg_query_matches(self.h, C.byref(qopt), _ptr(buf1), _ptr(offs1), _ptr(buf2), _ptr(offs2), n,
                                        _ptr(cnt), None, None))
        tot = int(cnt[:n].sum())
        tg = np.empty(max(tot, 1), np.uint32)
        wn = np.empty(max(tot, 1), np.uint32)
        _check(self.L.mcg_query_matches(self.h, C.byref(qopt), _ptr(buf1), _ptr(offs1), _ptr(buf2), _ptr(offs2), n,
                                        _ptr(cnt), _ptr(tg), _ptr(wn)))
        return cnt[:n], tg[:tot], wn[:tot]


def remove_overpopulated(dbs):
    L = load_library()
    arr = (C.c_void_p * len(dbs))(*[d.h for d in dbs])
    rem = C.c_uint64()
    _check(L.mcg_remove_overpopulated(arr, len(dbs), C.byref(rem)))
    return rem.value
