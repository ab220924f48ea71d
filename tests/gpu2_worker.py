"""Worker of tests/test_gpu_multi.py: one process per GPU (rank = device), the product library on every rank.
  * sharded self-play: rank r plays its block of the global game-id range; no collective
  * data-parallel learner: mz_learn_step_batch in BPTT mode on the rank's shard of a batch, ncclAllReduce inside the library
Results go to a pickle per rank; the parent compares them with the single-GPU run and the oracle."""
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from muzero_jl_b200 import capi, dist as mzdist  # noqa: E402

rank, world = int(sys.argv[1]), int(sys.argv[2])
work = sys.argv[3]
job = pickle.load(open(os.path.join(work, "job.pkl"), "rb"))
cfg = capi.default_config(**job["cfg"])
ctx = capi.Context(cfg, device=rank)
ctx.set_weights(job["blob"])
out = {"rank": rank}
# ---- sharded self-play (no collective): per-game results must not depend on the sharding ----
lo, cnt = mzdist.shard_games(rank, world, job["first_game"], job["n_games"])
sims, moves = ctx.self_play(lo, cnt, 1.0)
h = ctx.history_export()
out.update(lo=lo, cnt=cnt, sims=sims, hist={k: h[k] for k in h})
# ---- data-parallel learner ----
ctx.comm_init(rank, world, job["uid"])
out["comm_mode"] = ctx.comm_mode()
for name in ("halves", "equal_gscale"):
    ctx.set_weights(job["blob"]); ctx.optimizer_reset()
    shard = {k: v[rank::world] if name == "halves" else np.array_split(v, world)[rank] for k, v in job["batch_" + name].items()}
    losses = [ctx.learn_step(t, capi.GRAD_BPTT, shard) for t in (1, 2)]
    out["w_" + name] = ctx.get_weights(); out["losses_" + name] = np.stack(losses)
# the library's own sampling path under the communicator: every rank draws from ITS replay shard, weights stay identical
ctx.set_weights(job["blob"]); ctx.optimizer_reset()
ctx.learn_steps(1, 3, capi.GRAD_BPTT)
out["w_own_batches"] = ctx.get_weights()
ctx.comm_destroy()
# the same data-parallel steps through ncclAllReduce + the ADAM kernel (MUZERO_B200_DP=nccl): for two ranks a + b has one rounding, so
# the fused peer-memory update must give the same bits
os.environ["MUZERO_B200_DP"] = "nccl"
ctx.comm_init(rank, world, job["uid2"])
out["comm_mode_nccl"] = ctx.comm_mode()
ctx.set_weights(job["blob"]); ctx.optimizer_reset()
shard = {k: v[rank::world] for k, v in job["batch_halves"].items()}
for t in (1, 2):
    ctx.learn_step(t, capi.GRAD_BPTT, shard)
out["w_halves_nccl"] = ctx.get_weights()
ctx.comm_destroy()
ctx.close()
pickle.dump(out, open(os.path.join(work, "out%d.pkl" % rank), "wb"))
