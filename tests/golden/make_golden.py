"""Generates tests/golden/*.npz from the CPU oracle (oracle/mz_oracle.c).

The reference (pure Julia) cannot be executed in this image and ships no golden vectors, so these fixtures
freeze the ORACLE's outputs (parity unpinned beyond the known-answer tests in tests/test_oracle_kat.py).
They guard against silent drift of either the oracle or the CUDA path.  Re-run: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle as O  # noqa: E402
import common  # noqa: E402


def main():
    # 1. networks + MCTS on fixed roots (S=50, noise off and on)
    cfg = O.default_config(num_iters=50, exploration_eps=0.0)
    blob = O.init_weights(cfg, 1337)
    st, legal, tp = common.random_stacked(cfg, 48, seed=2024)
    game = np.arange(48, dtype=np.uint64) + 1000; move = (np.arange(48) % 9 + 1).astype(np.int32)
    h = np.stack([O.representation(cfg, blob, s) for s in st])
    v, p = zip(*[O.prediction(cfg, blob, x) for x in h])
    res = [O.run_mcts(cfg, blob, st[i], int(legal[i]), int(tp[i]), True, int(game[i]), int(move[i])) for i in range(48)]
    cfg_n = O.default_config(num_iters=50, exploration_eps=0.25)
    res_n = [O.run_mcts(cfg_n, blob, st[i], int(legal[i]), int(tp[i]), True, int(game[i]), int(move[i])) for i in range(48)]
    np.savez_compressed(os.path.join(HERE, "mcts_s50.npz"), stacked=st, legal=legal, to_play=tp, game=game, move=move, hidden=h,
                        value=np.array(v, np.float32), policy=np.stack(p),
                        vc=np.stack([r[0] for r in res]), rv=np.array([r[1] for r in res], np.float32), pri=np.stack([r[2] for r in res]),
                        vc_noise=np.stack([r[0] for r in res_n]), rv_noise=np.array([r[1] for r in res_n], np.float32),
                        pri_noise=np.stack([r[2] for r in res_n]), weight_checksum=np.float64(blob.astype(np.float64).sum()))
    # 2. self-play histories (reference defaults: S=10, noise on, T=1) + a batch + one learner step
    cfg = O.default_config()
    blob = O.init_weights(cfg, 1337)
    hist = O.self_play(cfg, blob, 0, 64, 1.0, 2)
    batch = O.get_batch(cfg, hist, step=1, first_key=1)
    pv, pr, pp, losses = O.learn_forward(cfg, blob, batch)
    w = blob.copy(); m = np.zeros_like(w); vv = np.zeros_like(w)
    l1 = O.learn_step(cfg, w, m, vv, 1, batch)
    b2 = O.get_batch(cfg, hist, step=2, first_key=1)
    l2 = O.learn_step(cfg, w, m, vv, 2, b2)
    np.savez_compressed(os.path.join(HERE, "selfplay_learn.npz"), **{"hist_" + k: hist[k] for k in common.HIST_KEYS},
                        **{"batch_" + k: batch[k] for k in common.BATCH_KEYS}, pred_values=pv, pred_rewards=pr, pred_policies=pp,
                        losses=losses, losses_step1=l1, losses_step2=l2, weights_after_2=w, sims=np.int64(hist["sims"]))
    # 3. paths added later: BPTT gradient, prioritised replay, ResNet networks (Float32 and bf16 emulation) and a ResNet search
    cfg = O.default_config(batch_size=24, per=1, intermediate_rewards=1)
    blob = O.init_weights(cfg, 77)
    hist = O.self_play(cfg, blob, 0, 40, 1.0, 2)
    q_pos, q_game = O.per_priorities(cfg, hist)
    pb = O.get_batch_per(cfg, hist, q_pos, q_game, 3)
    _, grad = O.learn_gradients_w(cfg, blob, pb, fwd64=False)
    pv, _, _, pl = O.learn_forward_w(cfg, blob, pb)
    q2, g2 = q_pos.copy(), q_game.copy()
    O.per_update(cfg, hist, q2, g2, pb["index"], pv, pb["values"])
    rcfg = O.resnet_config(num_iters=20, exploration_eps=0.0)
    rblob = O.init_weights(rcfg, 5)
    st, legal, tp = common.random_stacked(rcfg, 12, seed=31)
    rh = np.stack([O.representation(rcfg, rblob, x) for x in st])
    O.set_bf16(True)
    try:
        rh16 = np.stack([O.representation(rcfg, rblob, x) for x in st])
        rv16, rp16 = zip(*[O.prediction(rcfg, rblob, x) for x in rh16])
        rvc = np.stack([O.run_mcts(rcfg, rblob, st[i], int(legal[i]), int(tp[i]), False, 500 + i, 1)[0] for i in range(12)])
    finally:
        O.set_bf16(False)
    np.savez_compressed(os.path.join(HERE, "bptt_per_resnet.npz"), **{"hist_" + k: hist[k] for k in common.HIST_KEYS}, q_pos=q_pos, q_game=q_game,
                        **{"pb_" + k: pb[k] for k in common.BATCH_KEYS + ("weights",)}, grad=grad.astype(np.float32), losses=pl, q_pos_after=q2,
                        q_game_after=g2, rn_stacked=st, rn_legal=legal, rn_to_play=tp, rn_hidden_f32=rh, rn_hidden_bf16=rh16,
                        rn_value_bf16=np.array(rv16, np.float32), rn_policy_bf16=np.stack(rp16), rn_visit_counts_bf16=rvc)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
