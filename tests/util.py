"""shared builders for the parity tests (seeded, small)"""
import numpy as np

from metacache_mpi_b200 import synth, taxonomy
from oracle import pyoracle as po


class World:
    """G synthetic genomes + taxonomy + reads, with oracle shards for P ranks"""

    def __init__(self, G=8, L=30000, tie_block=True, seed=3, npairs=300, readlen=150):
        self.G, self.L = G, L
        self.genomes = synth.make_genomes(G, L)
        if tie_block and G > 2 and L >= 113 * 90:
            # a stride-aligned block shared by most genomes: equal hit counts across shards
            b = 113 * 44
            for g in range(2, G):
                self.genomes[g][b:b + 113 * 40] = self.genomes[0][b:b + 113 * 40]
        self.ids, self.par, self.rk, self.species = synth.make_taxonomy(G)
        self.tabs = taxonomy.TaxonomyTables(self.ids, self.par, self.rk, self.species)
        self.tax = po.TaxTables(self.tabs.lineage, self.tabs.rank, G)
        self.b1, self.o1, self.b2, self.o2, self.truth = synth.make_read_pairs(self.genomes, npairs, readlen=readlen, seed=seed)
        self.buf, self.offs = po.concat(self.genomes)

    def oracle_shards(self, O, P, **kw):
        dbs = [O.db_create(shard=r, nshards=P, **kw) for r in range(P)]
        for r in range(P):
            for g in range(self.G):
                dbs[r].add_target(self.genomes[g], g)
            dbs[r].finalize()
        return dbs

    def gpu_shards(self, api, ctx, P, **kw):
        dbs = [api.Database(ctx, shard=r, nshards=P, **kw) for r in range(P)]
        ids = np.arange(self.G, dtype=np.uint32)
        for r in range(P):
            dbs[r].add_targets(self.buf, self.offs, ids)
            dbs[r].finalize()
        return dbs


def lcg_window(seed=42, n=128):
    st = seed
    out = []
    for _ in range(n):
        st = (st * 1664525 + 1013904223) & 0xFFFFFFFF
        out.append("ACGT"[st >> 30])
    return "".join(out)
