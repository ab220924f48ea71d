"""Which network shapes the fused BPTT kernel accepts (its saved activations live in one CTA's shared memory)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi
for kw in (dict(), dict(stacked_observations=2), dict(depth_dynamics=4, depth_state_head=4), dict(depth_prediction=5, depth_policy=3, depth_value=3),
           dict(num_unroll_steps=8), dict(num_unroll_steps=10), dict(depth_representation=6), dict(stacked_observations=2, depth_dynamics=3, depth_state_head=3, depth_reward=2)):
    ctx = capi.Context(capi.default_config(num_slots=64, num_iters=5, **kw)); ctx.init_weights(1); ctx.self_play(0, 64, 1.0)
    try:
        ctx.learn_gradients(ctx.get_batch(1), capi.GRAD_BPTT); r = "ok"
    except capi.MuZeroB200Error as e:
        r = str(e)[-90:]
    print(kw, "->", r); ctx.close()
