"""The C-ABI library loads, exports every symbol include/mcgpu.h declares, and fails loudly
(no CPU fallback) when no GPU is present.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np

from metacache_mpi_b200 import api


def test_exports_every_declared_symbol():
    names = api.declared_symbols()
    assert len(names) >= 30 and "mcg_sketch_windows" in names and "mcg_merge_classify" in names
    L = api.load_library()
    for n in names:
        assert hasattr(L, n), n
    assert L.mcg_version() == 100


def test_header_cites_reference_seams():
    txt = open(api.HEADER).read()
    for cite in ("hash_dna.h:115", "sketch_database.h:519", "sketch_database.h:806", "querying.h:905",
                 "classification.cpp:236", "candidates.h:119", "hash_multimap.h:923"):
        assert cite in txt, cite
    assert "torch" not in re.sub(r"/\*.*?\*/", "", txt, flags=re.S)


def test_count_windows_is_host_only(oracle):
    L = api.load_library()
    cfg = api.SketchCfg(16, 16, 128, 113)
    offs = np.array([0, 0, 5, 133, 433, 674, 10674], np.int64)
    want = sum(len(oracle.windows(int(b - a), 128, 113)) for a, b in zip(offs[:-1], offs[1:]))
    assert L.mcg_count_windows(C.byref(cfg), offs.ctypes.data, len(offs) - 1) == want


def test_no_cpu_fallback_without_gpu():
    L = api.load_library()
    if L.mcg_device_count() > 0:
        return          # on the GPU box the gpu-marked tests exercise the compute path
    h = C.c_void_p()
    rc = L.mcg_ctx_create(0, None, C.byref(h))
    assert rc == -2 and b"no CPU fallback" in L.mcg_last_error()
    try:
        api.Context(0)
        raise AssertionError("Context() must raise without a GPU")
    except api.McgError:
        pass


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "metacache-mpi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) and f != "smoke.py":
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or "oracle/" not in src, os.path.join(dirpath, f)
                assert "pyoracle" not in src and "mc_oracle" not in src, os.path.join(dirpath, f)
