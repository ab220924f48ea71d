"""End-to-end use of the path: self-play -> BPTT learner -> competitive play against the random / expert opponent on a
context of its own (the quality metric the reference's readme asks for and never measures).  env: ROUNDS, GAMES, STEPS, BATCH, LR"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi

ROUNDS = int(os.environ.get("ROUNDS", 8)); GAMES = int(os.environ.get("GAMES", 4096)); STEPS = int(os.environ.get("STEPS", 50))
BATCH = int(os.environ.get("BATCH", 1024)); LR = float(os.environ.get("LR", 0.003)); S = int(os.environ.get("S", 50))
NN = {"fp32": capi.NN_FP32_EXACT, "split": capi.NN_SPLIT_MMA, "tc": capi.NN_BF16_TC}[os.environ.get("NN", "fp32")]   # split: search and learner on the tensor cores
kw = dict(num_slots=4096, num_iters=S, replay_buffer_size=20000, batch_size=BATCH, lr_init=LR, training_steps=ROUNDS * STEPS, nn_mode=NN)
cfg = capi.default_config(**{k: v for k, v in kw.items() if hasattr(capi.default_config(), k)})
train = capi.Context(cfg); train.init_weights(1337)
arena = capi.Context(capi.default_config(num_slots=2048, num_iters=S, replay_buffer_size=4096, nn_mode=NN))
print("nn_mode", os.environ.get("NN", "fp32"), "learner path", train.learner_path(capi.GRAD_BPTT))


def evaluate(tag):
    arena.set_weights(train.get_weights())
    out = []
    for opp, name in ((capi.OPP_RANDOM, "random"), (capi.OPP_EXPERT, "expert")):
        for side in (1, 2):
            arena.replay_clear()
            r = arena.arena(10 ** 6, 2000, opp, side, 0.0)
            out.append("%s/p%d %4d-%4d-%4d" % (name, side, r["wins"], r["draws"], r["losses"]))
    print("%-10s wins-draws-losses of 2000: %s" % (tag, "   ".join(out)), flush=True)


evaluate("init")
game = 0; step = 0
for rnd in range(ROUNDS):
    t0 = time.time()
    sims, _ = train.self_play(game, GAMES, 1.0); game += GAMES
    losses = train.learn_steps(step + 1, STEPS, capi.GRAD_BPTT); step += STEPS
    print("round %d: %d sims, %d steps, losses %s, %.2f s" % (rnd, sims, STEPS, losses, time.time() - t0), flush=True)
    evaluate("round %d" % rnd)
