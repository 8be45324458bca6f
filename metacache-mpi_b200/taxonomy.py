"""Host-side flattening of the taxonomy into the tables the device consumes
(mirror of taxonomy.h:576-597 `ranks()` and sketch_database.h:150,546: the
sequence-level taxon of target t has id -(t)-1 and parent = the target's taxid).

Taxon INDEX space: [0, ntargets) = sequence-level taxa of the targets (index ==
target id); [ntargets, ntaxa) = the taxa above sequence level in the order given.
"""
import numpy as np

NUM_RANKS = 21
RANK_SEQUENCE = 0
RANK_NONE = 21


class TaxonomyTables:
    def __init__(self, ids, parents, ranks, target_parent_taxids):
        ids = np.asarray(ids, np.int64)
        parents = np.asarray(parents, np.int64)
        ranks = np.asarray(ranks, np.uint8)
        tpar = np.asarray(target_parent_taxids, np.int64)
        nt = len(tpar)
        self.ntargets = nt
        self.ntaxa = nt + len(ids)
        # taxon id per index
        self.taxid = np.concatenate([-(np.arange(nt, dtype=np.int64)) - 1, ids])
        self.parent_id = np.concatenate([np.where(tpar < 1, 0, tpar), parents])
        self.rank = np.concatenate([np.zeros(nt, np.uint8), ranks]).astype(np.uint8)
        self.index_of = {int(t): i for i, t in enumerate(self.taxid)}
        # index-space parent (-1 = none / missing)
        pidx = np.full(self.ntaxa, -1, np.int64)
        for i in range(self.ntaxa):
            p = int(self.parent_id[i])
            if p != 0 and p in self.index_of:
                pidx[i] = self.index_of[p]
        self.parent_index = pidx
        # lineage of the taxa above sequence level first (memoised walk), then targets
        lin = np.full((self.ntaxa, NUM_RANKS), -1, np.int32)
        for i in range(nt, self.ntaxa):
            lin[i] = self._walk(i)
        if nt:
            # target lineage = own slot + lineage of parent chain (same walk)
            par = pidx[:nt]
            has = par >= 0
            lin[:nt][has] = lin[par[has]]
            # a later (higher) ancestor overwrites a lower one of the same rank, so the
            # target itself (rank Sequence) is only kept if no ancestor has rank Sequence
            own = np.arange(nt, dtype=np.int32)
            keep = lin[:nt, RANK_SEQUENCE] < 0
            lin[:nt, RANK_SEQUENCE] = np.where(keep, own, lin[:nt, RANK_SEQUENCE])
        self.lineage = np.ascontiguousarray(lin)

    def _walk(self, i):
        """taxonomy.h:576-597: walk towards the root; every ranked taxon on the way
        (including the start) writes its slot, later ones overwrite"""
        lin = np.full(NUM_RANKS, -1, np.int32)
        seen = 0
        cur = i
        while cur >= 0 and seen <= self.ntaxa:
            r = int(self.rank[cur])
            if r != RANK_NONE:
                lin[r] = cur
            nxt = int(self.parent_index[cur])
            if nxt == cur:
                break
            # parent id 0 or unknown parent terminates the walk
            cur = nxt
            seen += 1
        return lin

    def ids_of(self, idx):
        idx = np.asarray(idx)
        out = np.zeros(idx.shape, np.int64)
        m = idx >= 0
        out[m] = self.taxid[idx[m]]
        return out
