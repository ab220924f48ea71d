"""Small end-to-end case of the split-precision tensor-core paths for compute-sanitizer: self-play, run_mcts, network callables, arena,
learner forward (reference_l2) and the tensor-core BPTT (two batch sizes: one CTA / several CTAs with a ragged tail)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi

for B in (24, 70):
    ctx = capi.Context(capi.default_config(num_slots=64, num_iters=12, replay_buffer_size=128, nn_mode=capi.NN_SPLIT_MMA, batch_size=B, intermediate_rewards=1))
    ctx.init_weights(5)
    print("self_play", ctx.self_play(0, 80, 1.0))
    st = np.zeros((40, 63), np.float32); st[:, 18:27] = 1
    vc, rv = ctx.run_mcts(st, np.full(40, 0x1ff, np.uint32), np.ones(40, np.int32), True, np.arange(40, dtype=np.uint64), np.ones(40, np.int32))
    print("mcts", vc.sum(), ctx.representation(st).shape, ctx.learn_step(1))
    print("arena", ctx.arena(1000, 70, capi.OPP_EXPERT, 2, 0.0))
    print("bptt path", ctx.learner_path(capi.GRAD_BPTT), ctx.learn_steps(2, 2, capi.GRAD_BPTT))
    print("after update", ctx.self_play(1000, 40, 0.5))
    ctx.close()
print("sanitize case ok")
