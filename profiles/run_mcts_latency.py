"""Latency of one run_mcts call through the C ABI vs the number of roots (the reference calls it with one root at a time)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi
for mode, name in ((capi.NN_FP32_EXACT, "exact"), (capi.NN_BF16_TC, "tensor-core")):
    ctx = capi.Context(capi.default_config(num_slots=4096, num_iters=50, nn_mode=mode)); ctx.init_weights(1)
    for n in (1, 32, 1024, 4096):
        st = np.zeros((n, 63), np.float32); st[:, 18:27] = 1
        args = (st, np.full(n, 0x1ff, np.uint32), np.ones(n, np.int32), True, np.arange(n, dtype=np.uint64), np.ones(n, np.int32))
        ctx.run_mcts(*args)
        t0 = time.perf_counter()
        for _ in range(10):
            ctx.run_mcts(*args)
        dt = (time.perf_counter() - t0) / 10
        print("%-12s n = %5d roots: %.3f ms per call, %.2f M simulations/s" % (name, n, dt * 1e3, n * 50 / dt / 1e6))
    ctx.close()
