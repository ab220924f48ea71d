"""Profiling driver for mz_k_search_lat: run_mcts calls with N roots (env N, default 8) through the C ABI."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi
n = int(os.environ.get("N", 8))
ctx = capi.Context(capi.default_config(num_slots=256, num_iters=50)); ctx.init_weights(1337)
st = np.zeros((n, 63), np.float32); st[:, 18:27] = 1
args = (st, np.full(n, 0x1ff, np.uint32), np.ones(n, np.int32), True, np.arange(n, dtype=np.uint64), np.ones(n, np.int32))
for _ in range(3):
    vc, rv = ctx.run_mcts(*args)
print(vc[0], rv[0]); ctx.close()
