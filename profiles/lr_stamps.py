"""Cycle stamps of CTA 0 of mz_k_learn_bptt_tc (profiling build: make -C muzero.jl_b200/csrc -B OUT=../../profiles/libmuzero_timers.so EXTRA=-DMZ_PHASE_TIMERS;
run with MUZERO_B200_LIB=profiles/libmuzero_timers.so MUZERO_B200_LR_STAMPS=1)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi
B = int(os.environ.get("B", 4096)); N = 10
ctx = capi.Context(capi.default_config(nn_mode=capi.NN_SPLIT_MMA, num_slots=1024, replay_buffer_size=10000, batch_size=B)); ctx.init_weights(3)
ctx.self_play(0, 3000, 1.0)          # (zeroes the counters at its start)
ctx.learn_steps(1, N, capi.GRAD_BPTT)
raw = ctx.phase_cycles().astype(float)
names = ["set-up + representation", "fwd staging", "fwd rounds (own)", "fwd wait other group", "fwd rows", "drain + bwd weights", "bwd loss grad + staging", "bwd rounds (own)", "bwd wait other group", "representation bwd + teardown"]
for who, o in (("thread 0 (prediction group)", 0), ("thread 128 (dynamics group)", 16)):
    v = raw[o:o + 10] / N
    print(who + ": total %.0f cycles per launch; " % v.sum() + ", ".join("%s %.0f" % (names[i], v[i]) for i in range(10)))
ctx.close()
