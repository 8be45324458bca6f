"""Launcher side of the multi-GPU path: one process per GPU (torchrun), torch.distributed for the
rendezvous only.  The data plane -- exchange of sketches and per-read top lists over NVLink, the
distributed -remove-overpopulated-features reduction -- lives in libmcgpu.so behind mcg_comm_*,
mcg_query_batch_sharded and mcg_db_remove_overpopulated_dist (csrc/comm.cu, csrc/query.cu).

  partitioning : target t lives on rank t % world                      (sketch_database.h:540)
  reads        : a batch is cut into `world` slices in rank order; rank r ingests and sketches slice r
  merge        : rank r merges the per-shard top lists of slice r in RANK ORDER (querying.h:892-1071:
                 tree reduce == rank order 0..P-1 for power-of-two P)
"""
import numpy as np
import torch.distributed as dist


def owned_targets(ntargets, rank, world):
    """target ids of this rank's shard, ascending"""
    return np.arange(rank, ntargets, world, dtype=np.uint32)


def slice_bounds(n, world):
    """read index ranges of the `world` slices of a batch of n reads (slice r = [b[r], b[r + 1]))"""
    return [n * r // world for r in range(world + 1)]


def make_comm(ctx):
    """mcg_comm of this process: rank 0 draws the id, the torch.distributed store carries it (any backend)"""
    from . import api
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return api.Comm(ctx, b"", 0, 1)
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [api.Comm.new_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return api.Comm(ctx, box[0], rank, world)
