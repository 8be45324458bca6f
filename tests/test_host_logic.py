"""host-side logic: synthetic data determinism, taxonomy flattening, option defaults"""
import numpy as np

from metacache_mpi_b200 import api, synth, taxonomy


def test_synth_is_deterministic_and_shaped():
    a, b = synth.genome(3, 10000), synth.genome(3, 10000)
    assert (a == b).all() and len(a) == 10000
    even, odd = synth.genome(2, 20000), synth.genome(3, 20000)
    same = (np.char.upper(even.view("S1")) == np.char.upper(odd.view("S1"))).mean()
    assert 0.93 < same < 0.99          # odd genome = 5 % substitutions of its even sibling
    assert (a == ord("N")).sum() >= 50 and np.isin(a, np.frombuffer(b"ACGTNacgt", np.uint8)).all()
    g = [synth.genome(i, 5000) for i in range(3)]
    b1, o1, b2, o2, t = synth.make_read_pairs(g, 50, seed=1)
    c1, p1, c2, p2, t2 = synth.make_read_pairs(g, 50, seed=1)
    assert (b1 == c1).all() and (b2 == c2).all() and (t == t2).all()
    assert len(b1) == 50 * 150 and o1[-1] == 50 * 150 and (np.diff(o1) == 150).all()


def test_taxonomy_tables_shape_and_walk():
    ids, par, rk, species = synth.make_taxonomy(6)
    t = taxonomy.TaxonomyTables(ids, par, rk, species)
    assert t.lineage.shape == (6 + len(ids), 21) and t.lineage.dtype == np.int32
    for g in range(6):
        lin = t.lineage[g]
        assert lin[0] == g                                   # sequence level = the target itself
        assert t.taxid[lin[4]] == 100000 + g                 # species
        assert t.taxid[lin[6]] == 10000 + g // 2             # genus (through an unranked node)
        assert t.taxid[lin[10]] == 1000 + g // 4             # family
        assert t.taxid[lin[19]] == 2 and t.taxid[lin[20]] == 1
        assert (lin[[1, 2, 3, 5, 7, 8, 9]] == -1).all()
    # unknown parent / parent < 1: lineage ends at the target
    t2 = taxonomy.TaxonomyTables(ids, par, rk, np.array([0, -5, 999999, 100000, 100001, 100002]))
    assert (t2.lineage[0][1:] == -1).all() and (t2.lineage[1][1:] == -1).all() and (t2.lineage[2][1:] == -1).all()
    assert t2.parent_id[1] == 0


def test_option_defaults_follow_reference():
    assert api.default_hitmin(16) == 5 and api.default_hitmin(5) == 2 and api.default_hitmin(3) == 1   # mode_query.cpp:247
    q = api.make_query_opts(hitdiff=float(np.float32(np.float64(np.float32(80.0)) * 0.01)))       # query_options.cpp:167
    assert abs(q.hitdiff - 0.8) < 1e-6


def test_sorting_network_include_is_valid():
    """csrc/sortnet16.inc (the register network of the thread-per-window sketch kernel) must be a
    sorting network: 0-1 principle over all 2^16 inputs, checked on the committed file itself"""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "metacache-mpi_b200", "csrc", "sortnet16.inc")).read()
    ces = [(int(a), int(b)) for a, b in re.findall(r"SN_CE\((\d+), (\d+)\)", txt)]
    assert 55 <= len(ces) <= 63 and all(0 <= a < b < 16 for a, b in ces)
    import numpy as np
    v = ((np.arange(1 << 16)[:, None] >> np.arange(16)[None, :]) & 1).astype(np.uint8)
    for a, b in ces:
        lo = np.minimum(v[:, a], v[:, b])
        hi = np.maximum(v[:, a], v[:, b])
        v[:, a], v[:, b] = lo, hi
    assert (v[:, :-1] <= v[:, 1:]).all()


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference PROGRAM under the MPI shim on the host cores, no GPU) prints one
    JSON line with the keys of the measurement contract; here on a tiny sample (2 genomes, 2000 pairs)"""
    import json
    import os
    import subprocess
    import sys
    import pytest
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "oracle", "_ref", "metacache_mpi_ref")):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-genomes", "2", "--ref-pairs", "2000"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    lines = [l for l in r.stdout.decode().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "query_reads_per_s" and d["unit"] == "reads/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "sample" in d["config"] and "configs[3]" in d["config"]["workload"]
    assert d["build"]["value"] > 0


def test_counter_based_generator_follows_its_specification():
    """synth.genome_fast / read_pairs_fast (bench data, torch ops on any device) against the scalar definition"""
    import torch
    from metacache_mpi_b200 import synth
    L = 20000
    g2, g3 = synth.genome_fast(2, L), synth.genome_fast(3, L)
    for pos in (0, 1, 49, 50, 777, L - 1):
        want = "ACGT"[synth.hash_py(0, 2, pos) >> 62]
        if (synth.hash_py(3, 2, pos // 50) >> 32) < int(0.001 * (1 << 32)):
            want = "N"
        if (synth.hash_py(4, 2, pos) >> 32) < int(0.01 * (1 << 32)) and want != "N":
            want = want.lower()
        elif (synth.hash_py(4, 2, pos) >> 32) < int(0.01 * (1 << 32)):
            want = chr(ord("N") | 0x20)
        assert chr(int(g2[pos])) == want
    diff = float((g2 != g3).float().mean())
    assert 0.03 < diff < 0.08                                  # 5 % substitutions (3/4 of them change the base) + own N / case
    assert (synth.genome_fast(2, L) == g2).all()               # pure function of (genome, position)
    cat = torch.cat([g2, g3])
    b1, b2, gs = synth.read_pairs_fast(cat, L, 500)
    assert b1.numel() == 500 * 150 and set(gs.tolist()) == {0, 1}
    c1, c2, _ = synth.read_pairs_fast(cat, L, 100, first_pair=200)
    assert (c1 == b1[200 * 150:300 * 150]).all() and (c2 == b2[200 * 150:300 * 150]).all()     # slices of one stream
    assert 0.0005 < float((b1 == 78).float().mean()) < 0.01


def test_oracle_restricted_build_answers_like_the_whole_database(oracle):
    """bench.py's parity sample: a database restricted to the sample's features (multi-threaded build) gives the same
    per-read lists, taxa and statistics as the whole database"""
    import numpy as np
    from oracle import pyoracle as po
    from util import World
    w = World(G=8, L=30000, npairs=60)
    full = w.oracle_shards(oracle, 2, maxlocs=5)
    f1, _ = oracle.sketch_sequences(w.b1, w.o1)
    f2, _ = oracle.sketch_sequences(w.b2, w.o2)
    allowed = np.unique(np.concatenate([f1.ravel(), f2.ravel()]))
    allowed = allowed[allowed != 0xFFFFFFFF]
    rs = []
    for r in range(2):
        d = oracle.db_create(shard=r, nshards=2, maxlocs=5)
        d.add_targets_restricted(w.buf, w.offs, np.arange(w.G, dtype=np.uint32), allowed, threads=3)
        d.finalize()
        rs.append(d)
        assert d.counts()[0] < full[r].counts()[0]
    q = po.make_qopt(maxcand=4, lowest=4, hitmin=4, hitdiff=0.8)
    a = oracle.query_batch(full, w.tax, q, w.b1, w.o1, w.b2, w.o2)
    b = oracle.query_batch(rs, w.tax, q, w.b1, w.o1, w.b2, w.o2)
    assert all((x == y).all() for x, y in zip(a, b))


def test_top_list_selection_rule_equals_replayed_insertion(oracle):
    """The fused query kernel does not replay candidates.h:236-285 insert() candidate by candidate: it takes, K times, the
    candidate with the most hits (the earliest of those) and retires every candidate of its taxon (query.cu
    warp_toplist_select32).  Property checked here on the CPU, against the oracle's insertion (itself pinned to the
    reference's insert by tests/test_oracle_vs_reference.py): both give the same list for random candidate streams with
    many equal hit counts, taxa that appear several times, targets that stay at sequence level, and every list length."""
    from oracle import pyoracle as po
    from metacache_mpi_b200 import taxonomy as tx
    G = 24
    ids, par, rk, species = synth.make_taxonomy(G)
    species = species.copy()
    species[[3, 8, 9, 20]] = 0                                   # no taxon above the sequence: never merged, never lifted
    tabs = tx.TaxonomyTables(ids, par, rk, species)
    tax = po.TaxTables(tabs.lineage, tabs.rank, G)
    rng = np.random.default_rng(11)
    checked = 0
    for lowest in (0, 4, 6, 10):                                 # sequence, species, genus, family
        for K in (1, 2, 4, 7, 32):
            for trial in range(60):
                n = int(rng.integers(0, 33))
                t = rng.integers(0, G, n).astype(np.int32)       # (the kernel sees ascending distinct targets; any order must hold)
                if trial % 2 == 0:
                    t = np.sort(rng.permutation(G)[:min(n, G)]).astype(np.int32)
                    n = len(t)
                h = rng.integers(1, 4 if trial % 3 else 40, n).astype(np.uint32)
                q = po.make_qopt(maxcand=K, lowest=lowest, hitmin=1, hitdiff=1.0)
                want_t, want_h, _, _ = oracle.merge_classify(tax, q, t, h)
                # the selection rule, on lifted taxa
                lifted, seq = [], []
                for x in t:
                    a = tabs.lineage[x][lowest] if lowest > 0 and tabs.lineage[x][lowest] >= 0 else x
                    lifted.append(int(a)); seq.append(tabs.rank[a] == 0)
                alive = [True] * n
                got_t, got_h = [], []
                for _ in range(K):
                    best = -1
                    for j in range(n):
                        if alive[j] and (best < 0 or h[j] > h[best]):
                            best = j
                    if best < 0:
                        break
                    got_t.append(lifted[best]); got_h.append(int(h[best]))
                    for j in range(n):
                        if j == best or (not seq[best] and not seq[j] and lifted[j] == lifted[best]):
                            alive[j] = False
                assert list(want_t) == got_t and list(want_h) == got_h, (lowest, K, t, h)
                checked += 1
    assert checked == 4 * 5 * 60
