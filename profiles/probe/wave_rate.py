import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from muzero_jl_b200 import capi
for mode, name in ((capi.NN_FP32_EXACT, "exact"), (capi.NN_SPLIT_MMA, "split")):
    ctx = capi.Context(capi.default_config(num_slots=4096, num_iters=50, nn_mode=mode, replay_buffer_size=20000)); ctx.init_weights(1)
    ctx.self_play(0, 4096, 1.0)
    best = 1e9
    for r in range(4):
        t0 = time.perf_counter(); s, m = ctx.self_play(10000 * (r + 1), 4096, 1.0); dt = time.perf_counter() - t0; best = min(best, dt / s)
    print("%s: %.1f M simulations/s (one wave of 4096 games, wall clock)" % (name, 1e-6 / best))
    ctx.close()
