"""Device time per kernel family over a few self-play waves (mz_kernel_time): 0 search, 1 save/refill, 6 env/aux.  env: G, S, N, NN"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi
G = int(os.environ.get("G", 4096)); S = int(os.environ.get("S", 50)); N = int(os.environ.get("N", 3)); nn = os.environ.get("NN", "fp32")
cfg = capi.default_config(num_slots=G, num_iters=S, replay_buffer_size=max(10000, G), nn_mode=capi.NN_BF16_TC if nn == "tc" else capi.NN_FP32_EXACT)
ctx = capi.Context(cfg); ctx.init_weights(1337)
ctx.self_play(0, G, 1.0)
ctx.kernel_time_reset(True)
t0 = time.perf_counter()
for i in range(N):
    ctx.self_play((i + 1) * G, G, 1.0)
wall = (time.perf_counter() - t0) * 1e3
tot = 0.0
for fam, name in ((0, "search"), (1, "save/refill"), (6, "env/aux")):
    ms, n = ctx.kernel_time(fam)
    tot += ms
    print("%-12s %8.3f ms over %4d launches = %7.1f us each" % (name, ms, n, 1e3 * ms / max(n, 1)))
print("wall %.2f ms for %d waves; kernels %.2f ms; host gaps %.2f ms per wave" % (wall, N, tot, (wall - tot) / N))
