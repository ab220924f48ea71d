"""Profiling driver: N self-play waves of G games x S simulations (what bench.py times), nothing else.
env: G, S, N, NN (fp32 | tc | sp | rn | cn)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi

G = int(os.environ.get("G", 4096)); S = int(os.environ.get("S", 50)); N = int(os.environ.get("N", 2))
import time
nn = os.environ.get("NN", "fp32")
mode = capi.NN_BF16_TC if nn == "tc" else capi.NN_SPLIT_MMA if nn == "sp" else capi.NN_FP32_EXACT
if nn == "cn":     # 6x7 Connect game + ResNet (BASELINE.json configs[3]: 200 simulations per move)
    ctx = capi.Context(capi.connect_config(num_slots=G, num_iters=S, replay_buffer_size=max(10000, G)))
elif nn == "rn":     # ResNet networks (BASELINE.json config 3: 16384 concurrent games, bf16 inference)
    ctx = capi.Context(capi.resnet_config(num_slots=G, num_iters=S, replay_buffer_size=max(10000, G)))
else:
    ctx = capi.Context(capi.default_config(num_slots=G, num_iters=S, replay_buffer_size=max(10000, G), nn_mode=mode))
ctx.init_weights(1337)
for i in range(N):
    t0 = time.time()
    sims, moves = ctx.self_play(i * G, G, 1.0)
    dt = time.time() - t0
    print("wave", i, "sims", sims, "moves", moves, "wall %.1f ms" % (dt * 1e3), "%.1f M sims/s" % (sims / dt / 1e6))
print("waves", N, "sims", sims, "moves", moves, "launches", ctx.launch_count())
ctx.close()
