"""Latency of one run_mcts call through the C ABI vs the number of roots (the reference calls it with one root at a time).
MUZERO_B200_LAT=0 switches mz_k_search_lat off (the batched kernels then serve every call)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi
for lat in ("", "0"):
    if lat:
        os.environ["MUZERO_B200_LAT"] = lat
    for mode, name in ((capi.NN_FP32_EXACT, "exact"), (capi.NN_SPLIT_MMA, "split"), (capi.NN_BF16_TC, "bf16")):
        ctx = capi.Context(capi.default_config(num_slots=4096, num_iters=50, nn_mode=mode)); ctx.init_weights(1)
        for n in (1, 8, 32, 74, 148, 1024, 4096):
            st = np.zeros((n, 63), np.float32); st[:, 18:27] = 1
            args = (st, np.full(n, 0x1ff, np.uint32), np.ones(n, np.int32), True, np.arange(n, dtype=np.uint64), np.ones(n, np.int32))
            ctx.run_mcts(*args)
            t0 = time.perf_counter()
            for _ in range(20):
                ctx.run_mcts(*args)
            dt = (time.perf_counter() - t0) / 20
            ctx.kernel_time_reset(True); ctx.run_mcts(*args); ms, nl = ctx.kernel_time(0); ctx.kernel_time_reset(False)
            print("LAT=%-2s %-6s n = %5d roots: %.3f ms per call (search kernel %.3f ms), %.2f M simulations/s" % (lat or "on", name, n, dt * 1e3, ms, n * 50 / dt / 1e6))
        ctx.close()
