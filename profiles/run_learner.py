"""Profiling driver: N learner steps (get_batch gather + unroll forward [+ backward] + loss + gradient reduce + ADAM) at batch B.
env: B, N, MODE (bptt | l2), STACK (stacked_observations, default 1; 2 = 99 inputs: the fp32 kernel mz_k_learn_bptt)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi

B = int(os.environ.get("B", 4096)); N = int(os.environ.get("N", 3))
mode = capi.GRAD_BPTT if os.environ.get("MODE", "bptt") == "bptt" else capi.GRAD_REFERENCE_L2
STACK = int(os.environ.get("STACK", 1))
ctx = capi.Context(capi.default_config(num_slots=1024, num_iters=10, replay_buffer_size=4096, batch_size=B, stacked_observations=STACK))
ctx.init_weights(1337)
ctx.self_play(0, 2048, 1.0)
losses = ctx.learn_steps(1, N, mode)
print("steps", N, "B", B, "losses", losses, "launches", ctx.launch_count())
ctx.close()
