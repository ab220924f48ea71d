"""Small end-to-end case for compute-sanitizer: exact + tensor-core + ResNet self-play, arena, run_mcts API, replay gather (uniform and PER), learner steps."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi

for mode in (capi.NN_FP32_EXACT, capi.NN_BF16_TC):
    ctx = capi.Context(capi.default_config(num_slots=64, num_iters=12, replay_buffer_size=128, nn_mode=mode))
    ctx.init_weights(5)
    print("self_play", ctx.self_play(0, 80, 1.0))
    st = np.zeros((40, 63), np.float32); st[:, 18:27] = 1
    vc, rv = ctx.run_mcts(st, np.full(40, 0x1ff, np.uint32), np.ones(40, np.int32), True, np.arange(40, dtype=np.uint64), np.ones(40, np.int32))
    print("mcts", vc.sum(), ctx.learn_step(1), ctx.history_export()["T"].sum())
    print("arena", ctx.arena(1000, 70, capi.OPP_EXPERT, 2, 0.0), ctx.arena(2000, 70, capi.OPP_RANDOM, 1, 1.0))
    if mode == capi.NN_FP32_EXACT:
        print("bptt", ctx.learn_steps(2, 2, capi.GRAD_BPTT))
    ctx.close()
ctx = capi.Context(capi.default_config(num_slots=64, num_iters=12, replay_buffer_size=128, per=1, per_alpha=1))
ctx.init_weights(5)
print("per self_play", ctx.self_play(0, 80, 1.0), ctx.learn_steps(1, 2))
ctx.close()
ctx = capi.Context(capi.resnet_config(num_slots=30, num_iters=6, replay_buffer_size=64))
ctx.init_weights(5)
print("resnet self_play", ctx.self_play(0, 40, 1.0), ctx.arena(500, 20, capi.OPP_RANDOM, 1, 0.0))
ctx.close()
print("sanitize case ok")
