"""Small end-to-end case for compute-sanitizer: exact + tensor-core self-play, run_mcts API, replay gather, learner step."""
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
    ctx.close()
print("sanitize case ok")
