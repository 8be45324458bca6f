"""Multi-GPU plumbing of the query path: one process per GPU, torch.distributed (NCCL on
GPUs, gloo in the CPU tests).  Mirrors the reference's MPI structure:

  partitioning : target t lives on rank t % world           (sketch_database.h:540)
  reads        : every rank sees every read (broadcast once per batch)
  merge        : per-read top lists are gathered to every rank in RANK ORDER and inserted
                 in that order (querying.h:892-1071 tree reduce == rank order 0..P-1 for
                 power-of-two P)
"""
import numpy as np
import torch
import torch.distributed as dist


def owned_targets(ntargets, rank, world):
    """target ids of this rank's shard, ascending"""
    return np.arange(rank, ntargets, world, dtype=np.uint32)


def broadcast_reads(tensors, src=0):
    """broadcast a list of tensors (read bytes, offsets) from `src` to all ranks"""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for t in tensors:
            dist.broadcast(t, src=src)
    return tensors


def gather_candidates(tax, hits, ncand):
    """local [n][K], [n][K], [n] -> [P][n][K], [P][n][K], [P][n] in rank order on every rank"""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return tax.unsqueeze(0), hits.unsqueeze(0), ncand.unsqueeze(0)
    P = dist.get_world_size()
    n = tax.shape[0]
    # concatenated along dim 0 (the layout both gloo and NCCL accept), viewed as [P][n][...]
    gt = torch.empty((P * n,) + tuple(tax.shape[1:]), dtype=tax.dtype, device=tax.device)
    gh = torch.empty((P * n,) + tuple(hits.shape[1:]), dtype=hits.dtype, device=hits.device)
    gn = torch.empty((P * n,), dtype=ncand.dtype, device=ncand.device)
    dist.all_gather_into_tensor(gt, tax.contiguous())
    dist.all_gather_into_tensor(gh, hits.contiguous())
    dist.all_gather_into_tensor(gn, ncand.contiguous())
    return gt.view((P, n) + tuple(tax.shape[1:])), gh.view((P, n) + tuple(hits.shape[1:])), gn.view(P, n)


def reduce_key_sizes(keys, sizes, maxlocs):
    """-remove-overpopulated-features across ranks (mode_build.cpp:847-1074): every rank passes
    the (key, size) pairs of its non-empty buckets; returns the keys whose summed size exceeds
    maxlocs-1, identical on every rank"""
    keys = torch.as_tensor(np.ascontiguousarray(keys).astype(np.int64))
    sizes = torch.as_tensor(np.ascontiguousarray(sizes).astype(np.int64))
    if dist.is_initialized() and dist.get_world_size() > 1:
        P = dist.get_world_size()
        n = torch.tensor([keys.numel()], dtype=torch.int64)
        ns = [torch.zeros(1, dtype=torch.int64) for _ in range(P)]
        dist.all_gather(ns, n)
        m = int(max(int(x.item()) for x in ns))
        pk = torch.full((m,), -1, dtype=torch.int64)
        ps = torch.zeros((m,), dtype=torch.int64)
        pk[:keys.numel()] = keys
        ps[:sizes.numel()] = sizes
        gk = [torch.empty_like(pk) for _ in range(P)]
        gs = [torch.empty_like(ps) for _ in range(P)]
        dist.all_gather(gk, pk)
        dist.all_gather(gs, ps)
        keys = torch.cat([gk[r][:int(ns[r].item())] for r in range(P)])
        sizes = torch.cat([gs[r][:int(ns[r].item())] for r in range(P)])
    if keys.numel() == 0:
        return np.zeros(0, np.uint32)
    uk, inv = torch.unique(keys, return_inverse=True)
    tot = torch.zeros(uk.numel(), dtype=torch.int64).index_add_(0, inv, sizes)
    return uk[tot > (maxlocs - 1)].numpy().astype(np.uint32)
