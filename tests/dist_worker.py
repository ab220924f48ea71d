"""Worker for tests/test_distributed_cpu.py (gloo, world_size 2): sharded self-play + gradient averaging on the CPU
using the oracle as the per-rank engine, checking the HOST-SIDE multi-rank logic of muzero.jl_b200/dist.py."""
import os
import pickle
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from muzero_jl_b200 import dist as mzdist  # noqa: E402
from oracle import oracle as O  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
cfg = O.default_config(num_iters=8)
blob = O.init_weights(cfg, 1337)
N, FIRST = 24, 500
lo, cnt = mzdist.shard_games(rank, world, FIRST, N)
h = O.self_play(cfg, blob, lo, cnt, 1.0, 1)
# the unique-id exchange used for the NCCL communicator (any 128-byte payload)
uid = mzdist.broadcast_unique_id(lambda: (np.arange(128) * 7 % 251).astype(np.uint8), rank)
# data-parallel gradient: every rank holds grad = 2*theta (reference_l2); sum over ranks, average in the update
g = torch.from_numpy(blob + blob)
dist.all_reduce(g)
g = g.numpy() * np.float32(1.0 / world)
out = [None] * world
dist.gather_object(dict(rank=rank, lo=lo, cnt=cnt, T=h["T"], actions=h["actions"], cv=h["child_visits"], rv=h["root_values"], sims=h["sims"],
                        uid=uid, grad_ok=bool(np.array_equal(g, blob + blob))), out if rank == 0 else None, dst=0)
if rank == 0:
    with open(os.environ["MZ_DIST_OUT"], "wb") as f:
        pickle.dump(out, f)
dist.destroy_process_group()
