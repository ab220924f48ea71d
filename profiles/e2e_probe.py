import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from muzero_jl_b200 import capi
G, S = 4096, 50
cfg = capi.default_config(num_slots=G, num_iters=S, replay_buffer_size=10000)
stream = torch.cuda.Stream()
ctx = capi.Context(cfg, device=0, stream=stream.cuda_stream)
ctx.init_weights(1337); blob = ctx.get_weights()
host_hist = ctx.history_buffers(G, pinned=True)
host_blob = torch.from_numpy(blob).pin_memory().numpy()
for i in range(8):
    t0 = time.perf_counter(); ctx.set_weights(host_blob); t1 = time.perf_counter()
    ctx.self_play(i * G, G, 1.0); t2 = time.perf_counter()
    info = ctx.replay_info(); t3 = time.perf_counter()
    h = ctx.history_export(key0=info["first_key"] + info["n_games"] - G, n=G, out=host_hist); t4 = time.perf_counter()
    print("set_weights %.2f ms, self_play %.2f ms, replay_info %.2f ms, history_export %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
