"""Profiling driver: learner steps (BPTT, B = 4096) on the tensor-core path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi
B = int(os.environ.get("B", 4096))
ctx = capi.Context(capi.default_config(nn_mode=capi.NN_SPLIT_MMA, num_slots=1024, replay_buffer_size=10000, batch_size=B)); ctx.init_weights(3)
ctx.self_play(0, 3000, 1.0)
print(ctx.learn_steps(1, 6, capi.GRAD_BPTT), ctx.learner_path(capi.GRAD_BPTT))
ctx.close()
