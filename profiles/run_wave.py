"""Profiling driver: N self-play waves of G games x S simulations (what bench.py times), nothing else.
env: G, S, N, NN (fp32 | tc)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi

G = int(os.environ.get("G", 4096)); S = int(os.environ.get("S", 50)); N = int(os.environ.get("N", 2))
mode = capi.NN_BF16_TC if os.environ.get("NN", "fp32") == "tc" else capi.NN_FP32_EXACT
ctx = capi.Context(capi.default_config(num_slots=G, num_iters=S, replay_buffer_size=max(10000, G), nn_mode=mode))
ctx.init_weights(1337)
for i in range(N):
    sims, moves = ctx.self_play(i * G, G, 1.0)
print("waves", N, "sims", sims, "moves", moves, "launches", ctx.launch_count())
ctx.close()
