import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from muzero_jl_b200 import capi
ctx = capi.Context(capi.default_config(num_slots=256, num_iters=50)); ctx.init_weights(1)
for n in (1,):
    st = np.zeros((n, 63), np.float32); st[:, 18:27] = 1
    args = (st, np.full(n, 0x1ff, np.uint32), np.ones(n, np.int32), True, np.arange(n, dtype=np.uint64), np.ones(n, np.int32))
    ctx.run_mcts(*args); ctx.run_mcts(*args)
