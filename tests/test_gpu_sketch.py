"""K1 parity: CUDA sketches (through the C-ABI) == oracle, bit-exact"""
import numpy as np
import pytest

from metacache_mpi_b200 import api, synth
from oracle import pyoracle as po
from util import lcg_window

pytestmark = pytest.mark.gpu


def _check(ctx, oracle, seqs, **kw):
    buf, offs = po.concat(seqs)
    fo, co = oracle.sketch_sequences(buf, offs, k=kw.get("k", 16), sketchlen=kw.get("sketchlen", 16),
                                     winlen=kw.get("winlen", 128), stride=kw.get("winstride", 113))
    fg, cg = ctx.sketch_windows(buf, offs, **kw)
    assert fg.shape == fo.shape
    bad = np.nonzero((fg != fo).any(axis=1))[0]
    assert len(bad) == 0, "first differing window %d: gpu %s oracle %s" % (bad[0], fg[bad[0]], fo[bad[0]])
    assert (cg == co).all()


def test_known_answers(ctx, oracle):
    W = lcg_window()
    f, c = ctx.sketch_windows(*po.concat([W]))
    want = [0x03245e21, 0x03bdcb2e, 0x0635c602, 0x09aed6dd, 0x0b4abdff, 0x0ca44a78, 0x0e219d11, 0x0fcdf1a7,
            0x14267ca1, 0x151d5a6e, 0x16433551, 0x1b266a1a, 0x1c0d83c2, 0x1c5a4795, 0x1ed8d177, 0x20311862]
    assert f.shape == (1, 16) and list(f[0]) == want and c[0] == 16
    Wn = list(W)
    Wn[40], Wn[41], Wn[100] = "N", "n", "R"
    f, _ = ctx.sketch_windows(*po.concat(["".join(Wn)]))
    assert list(f[0][:4]) == [0x03245e21, 0x0635c602, 0x09aed6dd, 0x0b4abdff]
    f, c = ctx.sketch_windows(*po.concat([W[:20]]))
    assert c[0] == 5 and list(f[0][:5]) == [0x0635c602, 0x1c5a4795, 0x4b362346, 0x4e903cb8, 0xb7466199]
    assert (f[0][5:] == 0xFFFFFFFF).all()


def test_genomes_default_params(ctx, oracle):
    _check(ctx, oracle, [synth.genome(g, 200000) for g in range(3)])


def test_ragged_and_edge_sequences(ctx, oracle):
    rng = np.random.default_rng(5)
    seqs = [b"", b"A", b"ACGT" * 3, b"ACGTACGTACGTACGT", b"N" * 300, b"A" * 500, b"ACGT" * 100,
            b"acgtnACGTN" * 40]
    for n in (15, 16, 17, 127, 128, 129, 150, 240, 241, 242, 300, 1000, 4097):
        seqs.append(synth.random_bases(rng, n))
    _check(ctx, oracle, seqs)
    # reads: many short sequences
    _check(ctx, oracle, [synth.random_bases(rng, int(n)) for n in rng.integers(0, 400, size=700)])


@pytest.mark.parametrize("k,s,wl,st", [(16, 16, 128, 113), (12, 8, 64, 50), (16, 32, 256, 241), (5, 16, 128, 113),
                                       (16, 16, 20, 5), (1, 4, 10, 10), (16, 24, 127, 100), (8, 16, 128, 200),
                                       (16, 16, 1000, 985), (11, 32, 128, 128)])
def test_parameter_sweep(ctx, oracle, k, s, wl, st):
    seqs = [synth.genome(1, 20000), synth.genome(2, 777), b"ACGT" * 50, b"A" * 300]
    _check(ctx, oracle, seqs, k=k, sketchlen=s, winlen=wl, winstride=st)


def test_low_complexity_and_duplicates(ctx, oracle):
    # few distinct k-mers: the unfiltered fallback and the duplicate removal must agree
    seqs = [b"AC" * 300, b"ACG" * 200, (b"ACGTTGCA" * 16 + b"N") * 10, b"A" * 128 + b"C" * 128]
    _check(ctx, oracle, seqs)
    _check(ctx, oracle, seqs, sketchlen=32)


def test_full_size_properties(ctx):
    """size-independent properties at a larger size: ascending, unique, padded, strand-symmetric"""
    g = synth.genome(0, 5_000_000, n_runs=False, lower=False)
    buf, offs = po.concat([g])
    f, c = ctx.sketch_windows(buf, offs)
    valid = np.arange(16)[None, :] < c[:, None]
    assert (f[~valid] == 0xFFFFFFFF).all()
    d = np.diff(f.astype(np.int64), axis=1)
    assert (d[valid[:, 1:]] > 0).all()
    # reverse complement of a stride-aligned stretch gives the same sketches in reverse window order
    n = 113 * 1000 + 15
    seg = g[:n]
    f1, _ = ctx.sketch_windows(*po.concat([seg]))
    f2, _ = ctx.sketch_windows(*po.concat([synth.revcomp_ascii(seg)]))
    # 1000 full windows + one 15-base tail window (empty sketch) on both strands
    assert len(f1) == 1001 and (f1[:1000] == f2[:1000][::-1]).all()


def test_pipelined_host_path_many_chunks(ctx, oracle):
    """> 32 MB of sequence: the host entry point cuts the input into chunks of windows that are copied,
    sketched and copied back on three streams; chunk seams fall inside sequences"""
    rng = np.random.default_rng(11)
    seqs = [synth.genome(0, 30_000_000), synth.random_bases(rng, 100), synth.genome(2, 41_000_003), b"", synth.genome(4, 7_000_001)]
    buf, offs = po.concat(seqs)
    fg, cg = ctx.sketch_windows(buf, offs)
    fo, co = oracle.sketch_sequences(buf, offs)
    assert fg.shape == fo.shape
    bad = np.nonzero((fg != fo).any(axis=1))[0]
    assert len(bad) == 0, "first differing window %d" % bad[0]
    assert (cg == co).all()


@pytest.mark.parametrize("extra", ["-6", "40", "200"])
def test_rare_selection_paths(extra):
    """thread-per-window kernel with its filter moved (MCG_SK_EXTRA): too few survivors (exact
    warp redo), 49..64 survivors (fourth sort block), column overflow, keep-everything -- all
    must still equal the oracle.  Own process: the knob is read once per process."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from __graft_entry__ import load_package; load_package()\n"
        "from metacache_mpi_b200 import api, synth\n"
        "from oracle import pyoracle as po\n"
        "ctx = api.Context(0); orc = po.Oracle()\n"
        "rng = np.random.default_rng(11)\n"
        "seqs = [synth.genome(1, 60000), synth.genome(2, 777), b'ACGT' * 200, b'A' * 300, b'N' * 200 + synth.genome(3, 500).tobytes()]\n"
        "seqs += [synth.random_bases(rng, int(n)) for n in rng.integers(0, 400, size=200)]\n"
        "buf, offs = po.concat(seqs)\n"
        "for (k, s, wl, st) in [(16, 16, 128, 113), (12, 8, 64, 50), (16, 16, 256, 241), (16, 5, 128, 113)]:\n"
        "    fo, co = orc.sketch_sequences(buf, offs, k=k, sketchlen=s, winlen=wl, stride=st)\n"
        "    fg, cg = ctx.sketch_windows(buf, offs, k=k, sketchlen=s, winlen=wl, winstride=st)\n"
        "    assert fg.shape == fo.shape and (fg == fo).all() and (cg == co).all(), (k, s, wl, st)\n"
        "print('OK')\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MCG_SK_EXTRA=extra)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "OK" in out.stdout, out.stderr[-2000:]
