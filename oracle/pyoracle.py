"""oracle/pyoracle.py -- TEST INFRASTRUCTURE, not product code.

ctypes bindings for
  * liboracle.so          (mc_oracle.c, the plain-C restatement)      -> class Oracle
  * _ref/libref_l1.so     (ref_harness.cpp around the real reference) -> class Reference
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs import this module; the product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NUM_RANKS = 21
RANK_NONE = 21

_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    """compile liboracle.so (and oracle/_ref when /root/reference exists)"""
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "mc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/src/main.cpp"):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


def have_reference():
    return os.path.exists(os.path.join(HERE, "_ref", "libref_l1.so"))


def _bytes(s):
    if isinstance(s, (bytes, bytearray)):
        return bytes(s)
    if isinstance(s, str):
        return s.encode()
    return np.ascontiguousarray(s, dtype=np.uint8).tobytes()


def concat(seqs):
    """list of byte strings -> (uint8 buffer, int64 offsets[n+1])"""
    bs = [_bytes(s) for s in seqs]
    offs = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        offs[1:] = np.cumsum([len(b) for b in bs])
    buf = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    if buf.size == 0:
        buf = np.zeros(1, np.uint8)
    return buf, offs


class QOpt(C.Structure):
    _fields_ = [("sketchlen", C.c_int), ("winlen", C.c_int64), ("stride", C.c_int64),
                ("maxcand", C.c_int), ("lowest", C.c_int), ("highest", C.c_int),
                ("hitmin", C.c_int), ("hitdiff", C.c_float), ("insertsize", C.c_int64),
                ("reference_compat", C.c_int)]


class Tax(C.Structure):
    _fields_ = [("ntaxa", C.c_int64), ("ntargets", C.c_int64),
                ("lineage", C.c_void_p), ("rank", C.c_void_p)]


def make_qopt(sketchlen=16, winlen=128, stride=113, maxcand=2, lowest=0, highest=19,
              hitmin=5, hitdiff=1.0, insertsize=0, reference_compat=0):
    return QOpt(sketchlen, winlen, stride, maxcand, lowest, highest, hitmin,
                float(np.float32(hitdiff)), insertsize, reference_compat)


class Oracle:
    """plain-C restatement (oracle/mc_oracle.c)"""

    def __init__(self):
        build()
        L = C.CDLL(os.path.join(HERE, "liboracle.so"))
        self.L = L
        L.mco_tmh.restype = C.c_uint32
        L.mco_tmh.argtypes = [C.c_uint32]
        L.mco_revcomp.restype = C.c_uint32
        L.mco_revcomp.argtypes = [C.c_uint32, C.c_int]
        L.mco_canonical.restype = C.c_uint32
        L.mco_canonical.argtypes = [C.c_uint32, C.c_int]
        L.mco_kmers.restype = C.c_int64
        L.mco_kmers.argtypes = [C.c_char_p, C.c_int64, C.c_int, _u32p, _u16p]
        L.mco_num_windows.restype = C.c_int64
        L.mco_num_windows.argtypes = [C.c_int64] * 3
        L.mco_window.argtypes = [C.c_int64] * 4 + [C.POINTER(C.c_int64)] * 2
        L.mco_sketch_window.restype = C.c_int
        L.mco_sketch_window.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, _u32p]
        L.mco_sketch_sequences.restype = C.c_int64
        L.mco_sketch_sequences.argtypes = [_u8p, _i64p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64, _u32p, _u32p]
        L.mco_sketch_throughput.restype = C.c_double
        L.mco_sketch_throughput.argtypes = [_u8p, _i64p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int,
                                            C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.mco_db_create.restype = C.c_void_p
        L.mco_db_create.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int]
        L.mco_db_destroy.argtypes = [C.c_void_p]
        L.mco_db_maxlocs.restype = C.c_int
        L.mco_db_maxlocs.argtypes = [C.c_void_p]
        L.mco_db_add_target.restype = C.c_int64
        L.mco_db_add_target.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_uint32]
        L.mco_db_add_targets_restricted.restype = C.c_int64
        L.mco_db_add_targets_restricted.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                                    C.c_void_p, C.c_int64, C.c_int]
        L.mco_db_finalize.argtypes = [C.c_void_p]
        L.mco_db_counts.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint64)] * 3
        L.mco_db_export.argtypes = [C.c_void_p, _u32p, _u32p, _u32p, _u32p]
        L.mco_remove_overpopulated.restype = C.c_uint64
        L.mco_remove_overpopulated.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.mco_matches.restype = C.c_int64
        L.mco_matches.argtypes = [C.c_void_p, C.POINTER(QOpt), C.c_char_p, C.c_int64, C.c_char_p, C.c_int64,
                                  _u32p, _u32p, C.c_int64]
        L.mco_query_local.restype = C.c_int
        L.mco_query_local.argtypes = [C.c_void_p, C.POINTER(Tax), C.POINTER(QOpt), C.c_char_p, C.c_int64,
                                      C.c_char_p, C.c_int64, _i32p, _u32p, C.c_int]
        L.mco_merge_classify.restype = C.c_int
        L.mco_merge_classify.argtypes = [C.POINTER(Tax), C.POINTER(QOpt), _i32p, _u32p, C.c_int,
                                         _i32p, _u32p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int)]
        L.mco_statistics.argtypes = [_i32p, C.c_int64, _u64p]
        L.mco_query_batch.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(Tax), C.POINTER(QOpt),
                                      _u8p, _i64p, C.c_void_p, C.c_void_p, C.c_int64,
                                      _i32p, _u32p, _i32p, _i32p, _i32p, _u64p]

    # ---- L1
    def tmh(self, x):
        return self.L.mco_tmh(x)

    def revcomp(self, x, k):
        return self.L.mco_revcomp(x, k)

    def canonical(self, x, k):
        return self.L.mco_canonical(x, k)

    def kmers(self, s, k):
        s = _bytes(s)
        n = max(len(s), 1)
        km = np.zeros(n, np.uint32)
        am = np.zeros(n, np.uint16)
        c = self.L.mco_kmers(s, len(s), k, km, am)
        return km[:c], am[:c]

    def windows(self, length, winlen, stride):
        n = self.L.mco_num_windows(length, winlen, stride)
        out = []
        b, e = C.c_int64(), C.c_int64()
        for w in range(n):
            self.L.mco_window(length, winlen, stride, w, C.byref(b), C.byref(e))
            out.append((b.value, e.value))
        return out

    def sketch_window(self, s, k=16, sketchlen=16):
        s = _bytes(s)
        out = np.zeros(max(sketchlen, 1), np.uint32)
        c = self.L.mco_sketch_window(s, len(s), k, sketchlen, out)
        return out[:c].copy()

    def sketch_sequences(self, buf, offs, k=16, sketchlen=16, winlen=128, stride=113):
        nseq = len(offs) - 1
        nwin = sum(self.L.mco_num_windows(int(offs[i + 1] - offs[i]), winlen, stride) for i in range(nseq))
        feat = np.zeros((max(nwin, 1), sketchlen), np.uint32)
        cnt = np.zeros(max(nwin, 1), np.uint32)
        n = self.L.mco_sketch_sequences(buf, offs, nseq, k, sketchlen, winlen, stride, feat, cnt)
        assert n == nwin
        return feat[:nwin], cnt[:nwin]

    def sketch_throughput(self, buf, offs, k=16, sketchlen=16, winlen=128, stride=113, threads=1):
        cs, nw = C.c_uint64(), C.c_uint64()
        t = self.L.mco_sketch_throughput(buf, offs, len(offs) - 1, k, sketchlen, winlen, stride, threads,
                                         C.byref(cs), C.byref(nw))
        return t, cs.value, nw.value

    # ---- L2
    def db_create(self, k=16, sketchlen=16, winlen=128, stride=113, maxlocs=254, shard=0, nshards=1):
        return OracleDB(self, k, sketchlen, winlen, stride, maxlocs, shard, nshards)

    def remove_overpopulated(self, dbs):
        arr = (C.c_void_p * len(dbs))(*[d.h for d in dbs])
        return self.L.mco_remove_overpopulated(arr, len(dbs))

    # ---- L3/L4
    def merge_classify(self, tax, qopt, in_tax, in_hits):
        in_tax = np.ascontiguousarray(in_tax, np.int32)
        in_hits = np.ascontiguousarray(in_hits, np.uint32)
        cap = max(len(in_tax), 1)
        ot = np.zeros(cap, np.int32)
        oh = np.zeros(cap, np.uint32)
        best, brank = C.c_int32(), C.c_int()
        it = in_tax if len(in_tax) else np.zeros(1, np.int32)
        ih = in_hits if len(in_hits) else np.zeros(1, np.uint32)
        n = self.L.mco_merge_classify(C.byref(tax.c), C.byref(qopt), it, ih, len(in_tax), ot, oh, cap,
                                      C.byref(best), C.byref(brank))
        return ot[:n].copy(), oh[:n].copy(), best.value, brank.value

    def statistics(self, ranks):
        ranks = np.ascontiguousarray(ranks, np.int32)
        out = np.zeros(22, np.uint64)
        self.L.mco_statistics(ranks if len(ranks) else np.zeros(1, np.int32), len(ranks), out)
        return out

    def query_batch(self, dbs, tax, qopt, buf1, offs1, buf2, offs2):
        n = len(offs1) - 1
        K = qopt.maxcand
        arr = (C.c_void_p * len(dbs))(*[d.h for d in dbs])
        ot = np.full((max(n, 1), K), -1, np.int32)
        oh = np.zeros((max(n, 1), K), np.uint32)
        nc = np.zeros(max(n, 1), np.int32)
        best = np.zeros(max(n, 1), np.int32)
        brank = np.zeros(max(n, 1), np.int32)
        st = np.zeros(22, np.uint64)
        p2 = buf2.ctypes.data_as(C.c_void_p) if buf2 is not None else None
        o2 = offs2.ctypes.data_as(C.c_void_p) if buf2 is not None else None
        self.L.mco_query_batch(arr, len(dbs), C.byref(tax.c), C.byref(qopt), buf1, offs1, p2, o2, n,
                               ot, oh, nc, best, brank, st)
        return ot[:n], oh[:n], nc[:n], best[:n], brank[:n], st


class OracleDB:
    def __init__(self, o, k, sketchlen, winlen, stride, maxlocs, shard, nshards):
        self.o = o
        self.h = o.L.mco_db_create(k, sketchlen, winlen, stride, maxlocs, shard, nshards)

    def __del__(self):
        try:
            self.o.L.mco_db_destroy(self.h)
        except Exception:
            pass

    def maxlocs(self):
        return self.o.L.mco_db_maxlocs(self.h)

    def add_target(self, seq, target_id):
        s = _bytes(seq)
        return self.o.L.mco_db_add_target(self.h, s, len(s), target_id)

    def add_targets_restricted(self, buf, offs, ids, allowed, threads=1):
        """many targets at once, multi-threaded, keeping only the features in `allowed` (see mc_oracle.h)"""
        buf = np.ascontiguousarray(buf, np.uint8)
        offs = np.ascontiguousarray(offs, np.int64)
        ids = np.ascontiguousarray(ids, np.uint32)
        allowed = np.unique(np.ascontiguousarray(allowed, np.uint32))
        return self.o.L.mco_db_add_targets_restricted(self.h, buf.ctypes.data, offs.ctypes.data, ids.ctypes.data, len(ids),
                                                      allowed.ctypes.data, len(allowed), threads)

    def finalize(self):
        self.o.L.mco_db_finalize(self.h)

    def counts(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.o.L.mco_db_counts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def export(self):
        _, nk, nl = self.counts()
        keys = np.zeros(max(nk, 1), np.uint32)
        sizes = np.zeros(max(nk, 1), np.uint32)
        tg = np.zeros(max(nl, 1), np.uint32)
        wn = np.zeros(max(nl, 1), np.uint32)
        self.o.L.mco_db_export(self.h, keys, sizes, tg, wn)
        return keys[:nk], sizes[:nk], tg[:nl], wn[:nl]

    def matches(self, qopt, s1, s2=b""):
        s1, s2 = _bytes(s1), _bytes(s2)
        z = np.zeros(1, np.uint32)
        n = self.o.L.mco_matches(self.h, C.byref(qopt), s1, len(s1), s2, len(s2), z, z, 0)
        tg = np.zeros(max(n, 1), np.uint32)
        wn = np.zeros(max(n, 1), np.uint32)
        self.o.L.mco_matches(self.h, C.byref(qopt), s1, len(s1), s2, len(s2), tg, wn, n)
        return tg[:n], wn[:n]

    def query_local(self, tax, qopt, s1, s2=b"", cap=64):
        s1, s2 = _bytes(s1), _bytes(s2)
        ot = np.zeros(cap, np.int32)
        oh = np.zeros(cap, np.uint32)
        n = self.o.L.mco_query_local(self.h, C.byref(tax.c), C.byref(qopt), s1, len(s1), s2, len(s2), ot, oh, cap)
        assert n <= cap
        return ot[:n].copy(), oh[:n].copy()


class TaxTables:
    """flattened taxonomy (same layout as the product's mcg_tax_upload)"""

    def __init__(self, lineage, rank, ntargets):
        self.lineage = np.ascontiguousarray(lineage, np.int32)
        self.rank = np.ascontiguousarray(rank, np.uint8)
        self.ntargets = ntargets
        self.c = Tax(len(self.rank), ntargets, self.lineage.ctypes.data, self.rank.ctypes.data)


class Reference:
    """the real reference, via oracle/_ref/libref_l1.so"""

    def __init__(self):
        L = C.CDLL(os.path.join(HERE, "_ref", "libref_l1.so"))
        self.L = L
        L.ref_tmh.restype = C.c_uint32
        L.ref_tmh.argtypes = [C.c_uint32]
        L.ref_revcomp.restype = C.c_uint32
        L.ref_revcomp.argtypes = [C.c_uint32, C.c_int]
        L.ref_canonical.restype = C.c_uint32
        L.ref_canonical.argtypes = [C.c_uint32, C.c_int]
        L.ref_kmers.restype = C.c_int64
        L.ref_kmers.argtypes = [C.c_char_p, C.c_int64, C.c_int, _u32p, _u16p]
        L.ref_windows.restype = C.c_int64
        L.ref_windows.argtypes = [C.c_int64] * 3 + [_i64p, _i64p]
        L.ref_sketch_window.restype = C.c_int
        L.ref_sketch_window.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, _u32p]
        L.ref_sketch_sequence.restype = C.c_int64
        L.ref_sketch_sequence.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64, _u32p, _u32p]
        L.ref_sketch_throughput.restype = C.c_double
        L.ref_sketch_throughput.argtypes = [_u8p, _i64p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int,
                                            C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.refdb_create.restype = C.c_void_p
        L.refdb_create.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int]
        L.refdb_destroy.argtypes = [C.c_void_p]
        L.refdb_max_locations.restype = C.c_int
        L.refdb_max_locations.argtypes = [C.c_void_p]
        L.refdb_set_query_params.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64]
        L.refdb_set_taxonomy.argtypes = [C.c_void_p, _i64p, _i64p, _u8p, C.c_int64]
        L.refdb_add_target.restype = C.c_int
        L.refdb_add_target.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int, C.c_int]
        for f in ("refdb_target_count", "refdb_feature_count", "refdb_dead_feature_count", "refdb_location_count"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p]
        L.refdb_dump.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refdb_remove_keys.restype = C.c_uint64
        L.refdb_remove_keys.argtypes = [C.c_void_p, _u32p, C.c_uint32]
        L.refdb_matches.restype = C.c_int64
        L.refdb_matches.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, _u32p, _u32p, C.c_int64]
        L.refdb_query_local.restype = C.c_int
        L.refdb_query_local.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int, C.c_int,
                                        C.c_int64, _i64p, _u64p, C.c_int]
        L.refdb_merge_classify.restype = C.c_int
        L.refdb_merge_classify.argtypes = [C.c_void_p, _i64p, _u64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_float, _i64p, _u64p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int)]
        L.refdb_lineage.argtypes = [C.c_void_p, C.c_int64, _i64p]
        L.ref_statistics.argtypes = [_i32p, C.c_int64, _u64p]

    def tmh(self, x):
        return self.L.ref_tmh(x)

    def revcomp(self, x, k):
        return self.L.ref_revcomp(x, k)

    def canonical(self, x, k):
        return self.L.ref_canonical(x, k)

    def kmers(self, s, k):
        s = _bytes(s)
        n = max(len(s), 1)
        km = np.zeros(n, np.uint32)
        am = np.zeros(n, np.uint16)
        c = self.L.ref_kmers(s, len(s), k, km, am)
        return km[:c], am[:c]

    def windows(self, length, winlen, stride):
        cap = max(length // max(stride, 1) + 4, 4)
        b = np.zeros(cap, np.int64)
        e = np.zeros(cap, np.int64)
        n = self.L.ref_windows(length, winlen, stride, b, e)
        return [(int(b[i]), int(e[i])) for i in range(n)]

    def sketch_window(self, s, k=16, sketchlen=16):
        s = _bytes(s)
        out = np.zeros(max(sketchlen, 1), np.uint32)
        c = self.L.ref_sketch_window(s, len(s), k, sketchlen, out)
        return out[:c].copy()

    def sketch_sequence(self, s, k=16, sketchlen=16, winlen=128, stride=113):
        s = _bytes(s)
        cap = max(len(s) // max(stride, 1) + 4, 4)
        feat = np.zeros((cap, sketchlen), np.uint32)
        cnt = np.zeros(cap, np.uint32)
        n = self.L.ref_sketch_sequence(s, len(s), k, sketchlen, winlen, stride, feat, cnt)
        return feat[:n].copy(), cnt[:n].copy()

    def sketch_throughput(self, buf, offs, k=16, sketchlen=16, winlen=128, stride=113, threads=1):
        cs, nw = C.c_uint64(), C.c_uint64()
        t = self.L.ref_sketch_throughput(buf, offs, len(offs) - 1, k, sketchlen, winlen, stride, threads,
                                         C.byref(cs), C.byref(nw))
        return t, cs.value, nw.value

    def db_create(self, k=16, sketchlen=16, winlen=128, stride=113, maxlocs=254):
        return ReferenceDB(self, k, sketchlen, winlen, stride, maxlocs)

    def statistics(self, ranks):
        ranks = np.ascontiguousarray(ranks, np.int32)
        out = np.zeros(22, np.uint64)
        self.L.ref_statistics(ranks if len(ranks) else np.zeros(1, np.int32), len(ranks), out)
        return out


class ReferenceDB:
    """one rank's mc::database (sketch_database.h)"""

    def __init__(self, r, k, sketchlen, winlen, stride, maxlocs):
        self.r = r
        self.h = r.L.refdb_create(k, sketchlen, winlen, stride, maxlocs)

    def __del__(self):
        try:
            self.r.L.refdb_destroy(self.h)
        except Exception:
            pass

    def maxlocs(self):
        return self.r.L.refdb_max_locations(self.h)

    def set_query_params(self, sketchlen, winlen, stride):
        self.r.L.refdb_set_query_params(self.h, sketchlen, winlen, stride)

    def set_taxonomy(self, ids, parents, ranks):
        self.r.L.refdb_set_taxonomy(self.h, np.ascontiguousarray(ids, np.int64),
                                    np.ascontiguousarray(parents, np.int64),
                                    np.ascontiguousarray(ranks, np.uint8), len(ids))

    def add_target(self, seq, name, parent_taxid, num_procs, my_id):
        s = _bytes(seq)
        return self.r.L.refdb_add_target(self.h, s, len(s), name.encode(), parent_taxid, num_procs, my_id)

    def counts(self):
        L = self.r.L
        return (L.refdb_feature_count(self.h), L.refdb_dead_feature_count(self.h), L.refdb_location_count(self.h))

    def dump(self):
        """non-empty buckets sorted by key: keys, sizes, tgts, wins"""
        nk, nl = C.c_uint64(), C.c_uint64()
        self.r.L.refdb_dump(self.h, C.byref(nk), C.byref(nl), None, None, None, None)
        keys = np.zeros(max(nk.value, 1), np.uint32)
        sizes = np.zeros(max(nk.value, 1), np.uint32)
        tg = np.zeros(max(nl.value, 1), np.uint32)
        wn = np.zeros(max(nl.value, 1), np.uint32)
        self.r.L.refdb_dump(self.h, C.byref(nk), C.byref(nl), keys.ctypes.data, sizes.ctypes.data,
                            tg.ctypes.data, wn.ctypes.data)
        keys, sizes, tg, wn = keys[:nk.value], sizes[:nk.value], tg[:nl.value], wn[:nl.value]
        order = np.argsort(keys, kind="stable")
        starts = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64) if len(sizes) else np.zeros(0, np.int64)
        idx = np.concatenate([np.arange(starts[i], starts[i] + sizes[i]) for i in order]) if len(order) else np.zeros(0, np.int64)
        return keys[order], sizes[order], tg[idx.astype(np.int64)], wn[idx.astype(np.int64)]

    def remove_keys(self, keys):
        keys = np.ascontiguousarray(keys, np.uint32)
        if len(keys) == 0:
            return 0
        return self.r.L.refdb_remove_keys(self.h, keys, len(keys))

    def matches(self, s1, s2=b""):
        s1, s2 = _bytes(s1), _bytes(s2)
        z = np.zeros(1, np.uint32)
        n = self.r.L.refdb_matches(self.h, s1, len(s1), s2, len(s2), z, z, 0)
        tg = np.zeros(max(n, 1), np.uint32)
        wn = np.zeros(max(n, 1), np.uint32)
        self.r.L.refdb_matches(self.h, s1, len(s1), s2, len(s2), tg, wn, n)
        return tg[:n], wn[:n]

    def query_local(self, s1, s2, maxcand, lowest, insertsize=0, cap=64):
        s1, s2 = _bytes(s1), _bytes(s2)
        ot = np.zeros(cap, np.int64)
        oh = np.zeros(cap, np.uint64)
        n = self.r.L.refdb_query_local(self.h, s1, len(s1), s2, len(s2), maxcand, lowest, insertsize, ot, oh, cap)
        assert n <= cap
        return ot[:n].copy(), oh[:n].copy()

    def merge_classify(self, in_tax, in_hits, maxcand, lowest, highest, hitmin, hitdiff):
        in_tax = np.ascontiguousarray(in_tax, np.int64)
        in_hits = np.ascontiguousarray(in_hits, np.uint64)
        cap = max(len(in_tax), 1)
        ot = np.zeros(cap, np.int64)
        oh = np.zeros(cap, np.uint64)
        best, brank = C.c_int64(), C.c_int()
        it = in_tax if len(in_tax) else np.zeros(1, np.int64)
        ih = in_hits if len(in_hits) else np.zeros(1, np.uint64)
        n = self.r.L.refdb_merge_classify(self.h, it, ih, len(in_tax), maxcand, lowest, highest, hitmin,
                                          float(np.float32(hitdiff)), ot, oh, cap, C.byref(best), C.byref(brank))
        return ot[:n].copy(), oh[:n].copy(), best.value, brank.value

    def lineage(self, taxid):
        out = np.zeros(NUM_RANKS, np.int64)
        self.r.L.refdb_lineage(self.h, taxid, out)
        return out
