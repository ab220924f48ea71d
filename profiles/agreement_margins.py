import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import common
from oracle import oracle as O
from muzero_jl_b200 import capi
from test_oracle_resnet import _randomised_blob
for seed in (11, 12, 13):
    cfg = capi.resnet_config(num_slots=128, replay_buffer_size=512, num_iters=30, exploration_eps=0.0)
    ctx = capi.Context(cfg); ocfg = common.oracle_config(cfg)
    blob = _randomised_blob(ocfg, seed); ctx.set_weights(blob)
    n = 120
    st, legal, tp = common.random_stacked(ocfg, n, seed=seed + 10)
    gid = np.arange(n, dtype=np.uint64) + 1000; mv = np.ones(n, np.int32)
    vc, rv, pri = ctx.run_mcts(st, legal, tp, False, gid, mv, priors=True)
    O.set_bf16(True)
    same = top = 0; perr = 0.0
    for i in range(n):
        ovc, orv, opri = O.run_mcts(ocfg, blob, st[i], int(legal[i]), int(tp[i]), False, int(gid[i]), 1)
        same += int(np.array_equal(ovc, vc[i])); top += int(np.argmax(ovc) == np.argmax(vc[i])); perr = max(perr, float(np.max(np.abs(opri - pri[i]))))
    O.set_bf16(False)
    print("resnet seed", seed, "identical", same, "/", n, "same top", top, "prior err", perr)
    ctx.close()
