"""Small case for compute-sanitizer (memcheck / racecheck / synccheck): mz_k_search_lat in API and SLOTS mode, and the use_batch_norm kernels."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi
ctx = capi.Context(capi.default_config(num_slots=64, num_iters=8, replay_buffer_size=128)); ctx.init_weights(5)
st = np.zeros((5, 63), np.float32); st[:, 18:27] = 1
vc, rv = ctx.run_mcts(st, np.full(5, 0x1ff, np.uint32), np.ones(5, np.int32), True, np.arange(5, dtype=np.uint64), np.ones(5, np.int32))
print("lat run_mcts", vc.sum(1), "lat self_play", ctx.self_play(0, 3, 1.0), "arena", ctx.arena(100, 4, capi.OPP_EXPERT, 2, 0.0))
ctx.close()
ctx = capi.Context(capi.default_config(num_slots=32, num_iters=6, replay_buffer_size=64, use_batch_norm=1)); ctx.init_weights(5)
print("bn self_play", ctx.self_play(0, 40, 1.0), ctx.learn_steps(1, 2))
ctx.close()
print("sanitize lat ok")
