"""Device time of the search kernel per move and per simulation round for every network mode (CUDA events around each launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi
G, S = int(os.environ.get("G", 4096)), int(os.environ.get("S", 50))
ONLY = os.environ.get("MODES", "split_mma,bf16_tcgen05,fp32_exact").split(",")
for mode, name in ((capi.NN_SPLIT_MMA, "split_mma"), (capi.NN_BF16_TC, "bf16_tcgen05"), (capi.NN_FP32_EXACT, "fp32_exact")):
    if name not in ONLY:
        continue
    ctx = capi.Context(capi.default_config(num_slots=G, num_iters=S, replay_buffer_size=max(10000, G), nn_mode=mode)); ctx.init_weights(1337)
    ctx.self_play(0, G, 1.0)
    ctx.kernel_time_reset(True)
    sims = 0
    for i in range(3):
        sims += ctx.self_play((1 + i) * G, G, 1.0)[0]
    ms, n = ctx.kernel_time(0)
    print("%-13s %6.1f M simulations/s of search-kernel time; %.3f ms per launch (%d launches), %.1f us = %.0f cycles per simulation round"
          % (name, sims / ms / 1e3, ms / n, n, 1e3 * ms / n / S, 1e3 * ms / n / S * 1965))
    ctx.close()
