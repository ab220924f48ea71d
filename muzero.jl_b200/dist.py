"""Host-side multi-GPU plumbing (one process per GPU, torch.distributed for rendezvous only).

Self-play shards by game id and needs no collective: rank r plays a contiguous block of the global game-id range and
every random stream is keyed by the global game id, so results do not depend on the number of ranks.  The learner is
data-parallel: the NCCL communicator lives inside the C library (`mz_comm_init`); torch.distributed only carries the
128-byte ncclUniqueId from rank 0 to the others.
"""
import numpy as np


def shard_games(rank, world, first_game, n_games):
    """Contiguous block of game ids [first, first+count) owned by `rank`."""
    lo = first_game + (n_games * rank) // world
    hi = first_game + (n_games * (rank + 1)) // world
    return lo, hi - lo


def broadcast_unique_id(make_uid, rank, device=None):
    """rank 0 creates the id (Context.comm_unique_id), everyone receives it.  Works on gloo (CPU) and nccl (GPU)."""
    import torch
    import torch.distributed as dist
    uid = torch.from_numpy(np.ascontiguousarray(make_uid() if rank == 0 else np.zeros(128, np.uint8)))
    if device is not None:
        uid = uid.to(device)
    dist.broadcast(uid, 0)
    return uid.cpu().numpy()


def attach_communicator(ctx, rank, world, device=None):
    """Create the library's NCCL communicator for a data-parallel learner."""
    from .capi import Context
    uid = broadcast_unique_id(Context.comm_unique_id, rank, device)
    ctx.comm_init(rank, world, uid)
