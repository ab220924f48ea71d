"""Learner step time (get_batch gather + unroll forward + backward + reduce + ADAM) for the tensor-core and the fp32 SIMT BPTT paths."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi
for B in (32, 4096):
    ctx = capi.Context(capi.default_config(nn_mode=capi.NN_SPLIT_MMA, num_slots=1024, replay_buffer_size=10000, batch_size=B)); ctx.init_weights(3)
    ctx.self_play(0, 3000, 1.0)
    for name, env in (("tcgen05", None), ("fp32 SIMT", "1")):
        if env: os.environ["MUZERO_B200_BPTT_SIMT"] = env
        elif "MUZERO_B200_BPTT_SIMT" in os.environ: del os.environ["MUZERO_B200_BPTT_SIMT"]
        ctx.learn_steps(1, 5, capi.GRAD_BPTT)
        ctx.kernel_time_reset(True)
        t0 = time.perf_counter(); l = ctx.learn_steps(6, 20, capi.GRAD_BPTT); dt = (time.perf_counter() - t0) / 20
        ms3, n3 = ctx.kernel_time(3); ms4, n4 = ctx.kernel_time(4)
        ctx.kernel_time_reset(False)
        print("B %4d %-10s %.3f ms/step wall (%.2f M samples/s); kernels: unroll+loss %.3f ms, reduce+adam %.3f ms; losses %s" % (B, name, dt * 1e3, B / dt / 1e6, ms3 / 20, ms4 / 20, l))
    ctx.close()
