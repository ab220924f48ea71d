"""Throughput of one self_play call of 4 x G games on G slots under the two refill policies (MUZERO_B200_REFILL=immediate | wave), per network mode."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi
for name, mk in (("split", lambda: capi.default_config(num_slots=4096, num_iters=50, replay_buffer_size=40000, nn_mode=capi.NN_SPLIT_MMA)),
                 ("exact", lambda: capi.default_config(num_slots=4096, num_iters=50, replay_buffer_size=40000)),
                 ("resnet", lambda: capi.resnet_config(num_slots=16384, num_iters=50, replay_buffer_size=70000))):
    for pol in ("immediate", "wave"):
        os.environ["MUZERO_B200_REFILL"] = pol
        ctx = capi.Context(mk()); ctx.init_weights(1337)
        G = ctx.cfg.num_slots
        ctx.self_play(0, G, 1.0)
        t0 = time.perf_counter(); sims, moves = ctx.self_play(G, 4 * G, 1.0); dt = time.perf_counter() - t0
        print("%-7s %-10s %6.1f M simulations/s (%d games, %d launches)" % (name, pol, sims / dt / 1e6, 4 * G, ctx.launch_count()))
        ctx.close()
