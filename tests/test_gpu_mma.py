"""GPU tests of the split-precision tensor-core path (nn_mode = MZ_NN_SPLIT_MMA, mz_kernels_sp.cuh): bf16 hi + lo operands on
tcgen05.mma, fp32 accumulation in TMEM.  Not bit-exact by construction (the tensor core sums in its own order); what is asserted:
network outputs within 2e-5 of the Float32 oracle, visit counts identical to the FLOAT32 oracle on >= 99 % of 1024 roots, the same
most-visited action on >= 99.9 %, and everything that does not depend on network arithmetic (histories' structure, temperature rule,
slot-count independence) exactly."""
import numpy as np
import pytest

import common
from oracle import oracle as O

pytestmark = pytest.mark.gpu
MMA_ATOL = 2e-5


@pytest.fixture(scope="module")
def capi():
    from muzero_jl_b200 import capi
    return capi


def make(capi, **kw):
    kw.setdefault("num_slots", 256); kw.setdefault("replay_buffer_size", 1024)
    cfg = capi.default_config(nn_mode=capi.NN_SPLIT_MMA, **kw)
    return capi.Context(cfg), common.oracle_config(cfg)


@pytest.mark.parametrize("kw", [dict(), dict(width_hidden=48, depth_prediction=1, depth_value=0, depth_policy=2, depth_dynamics=0, depth_state_head=1, depth_reward=2),
                                dict(width_hidden=32, stacked_observations=0, depth_representation=0)])
def test_mma_networks_match_float32_oracle(capi, kw):
    ctx, ocfg = make(capi, **kw)
    ctx.init_weights(1337); blob = ctx.get_weights()
    rng = np.random.default_rng(5)
    blob[:] = blob + np.where(blob == 0, rng.uniform(-0.2, 0.2, blob.shape), 0).astype(np.float32)     # non-zero biases
    ctx.set_weights(blob)
    n = 70
    st, legal, tp = common.random_stacked(ocfg, n, seed=9)
    h = ctx.representation(st)
    oh = np.stack([O.representation(ocfg, blob, x) for x in st])
    assert np.max(np.abs(h - oh)) <= MMA_ATOL * max(1.0, float(np.max(np.abs(oh))))
    v, p = ctx.prediction(oh)
    ov, op = zip(*[O.prediction(ocfg, blob, x) for x in oh])
    assert np.max(np.abs(v - np.array(ov))) <= MMA_ATOL and np.max(np.abs(p - np.stack(op))) <= MMA_ATOL
    for scale in (1.0, 64.0):                                             # hidden states after several in-place doublings (Q6)
        sa = np.concatenate([oh * 2 * scale, np.full((n, 9), np.float32(5.0 / 9.0), np.float32)], axis=1)
        nh, r = ctx.dynamics(sa)
        onh, orr = zip(*[O.dynamics(ocfg, blob, x) for x in sa])
        onh = np.stack(onh)
        # 16 mantissa bits per operand: the error scales with the magnitude of the inputs
        assert np.max(np.abs(nh - onh)) <= MMA_ATOL * max(1.0, float(np.max(np.abs(onh)))) and np.max(np.abs(r - np.array(orr))) <= MMA_ATOL * max(1.0, scale / 4)
    # the oracle's emulation of the same operand split (same products, sequential accumulation instead of the tensor core's order)
    O.set_bf16(2)
    try:
        eh = np.stack([O.representation(ocfg, blob, x) for x in st])
    finally:
        O.set_bf16(0)
    assert np.max(np.abs(h - eh)) <= MMA_ATOL * max(1.0, float(np.max(np.abs(oh))))
    ctx.close()


@pytest.mark.parametrize("eps", [0.0, 0.25])
def test_mma_run_mcts_agrees_with_float32_oracle(capi, eps):
    n, S = 1024, 50
    ctx, ocfg = make(capi, num_iters=S, exploration_eps=eps, num_slots=1024)
    ctx.init_weights(7); blob = ctx.get_weights()
    st, legal, tp = common.random_stacked(ocfg, n, seed=77)
    game = np.arange(n, dtype=np.uint64) + 100; move = (np.arange(n) % 9 + 1).astype(np.int32)
    vc, rv, pri = ctx.run_mcts(st, legal, tp, True, game, move, priors=True)
    assert np.all(vc.sum(1) == S) and np.all((vc > 0) <= ((legal[:, None] >> np.arange(9)) & 1).astype(bool))
    exact = [O.run_mcts(ocfg, blob, st[i], int(legal[i]), int(tp[i]), True, int(game[i]), int(move[i])) for i in range(n)]
    same = np.mean([vc[i].tolist() == exact[i][0].tolist() for i in range(n)])
    best = np.mean([int(np.argmax(vc[i])) == int(np.argmax(exact[i][0])) for i in range(n)])
    perr = max(float(np.max(np.abs(pri[i] - exact[i][2]))) for i in range(n))
    print("split-precision MMA path, %d roots x %d simulations, eps %.2f: visit counts identical to the Float32 oracle %.4f, same most-visited action %.4f, "
          "max prior error %.2e" % (n, S, eps, same, best, perr))
    assert same >= 0.99 and best >= 0.999 and perr <= 1e-5
    ident = [i for i in range(n) if vc[i].tolist() == exact[i][0].tolist()]
    # identical root visit counts do not imply identical subtrees: a different path below the root moves the root value by O(1 / S)
    rerr = np.abs(rv[ident] - np.array([exact[i][1] for i in ident]))
    assert np.median(rerr) <= 1e-5 and np.max(rerr) <= 0.05
    ctx.close()


def test_mma_self_play_agrees_with_float32_oracle_and_is_independent_of_slot_count(capi):
    res = []
    for slots in (64, 4096):                                              # 1 and 28 trees per CTA
        ctx, ocfg = make(capi, num_slots=slots, replay_buffer_size=max(1024, slots), num_iters=30)
        ctx.init_weights(11); blob = ctx.get_weights()
        sims, moves = ctx.self_play(0, 600, 1.0)
        assert sims == moves * 30
        h = ctx.history_export(); order = np.argsort(h["game_id"])
        res.append({k: h[k][order] for k in common.HIST_KEYS}); ctx.close()
    for k in common.HIST_KEYS:
        assert np.array_equal(res[0][k], res[1][k]), k
    o = O.self_play(ocfg, blob, 0, 600, 1.0, 8)
    same_game = np.mean([all(np.array_equal(res[0][k][i], o[k][i]) for k in ("T", "actions", "child_visits")) for i in range(600)])
    same_first = np.mean([np.array_equal(res[0]["child_visits"][i, 0], o["child_visits"][i, 0]) for i in range(600)])
    print("split-precision MMA self-play: %.4f of 600 games identical to the Float32 oracle in every ply, %.4f in the first ply" % (same_game, same_first))
    assert same_first >= 0.99 and same_game >= 0.93
    for j in range(600):
        T = res[0]["T"][j]
        assert 6 <= T <= 9 and np.allclose(res[0]["child_visits"][j, :T].sum(1), 1.0, atol=1e-6) and np.all(res[0]["rewards"][j, :T - 1] == 0)


def test_mma_randomised_configurations_agree(capi):
    for seed in range(4):
        rng = np.random.default_rng(300 + seed)
        kw = dict(num_iters=int(rng.integers(5, 41)), stacked_observations=int(rng.integers(0, 2)), depth_representation=int(rng.integers(0, 4)),
                  depth_prediction=int(rng.integers(0, 4)), depth_dynamics=int(rng.integers(0, 4)), depth_policy=int(rng.integers(0, 3)),
                  depth_value=int(rng.integers(0, 3)), depth_reward=int(rng.integers(0, 3)), depth_state_head=int(rng.integers(0, 4)),
                  width_hidden=int(rng.choice([32, 48, 64])), num_slots=int(rng.integers(8, 200)), seed=int(rng.integers(1, 1 << 30)),
                  exploration_eps=float(np.float32(rng.choice([0.0, 0.25]))))
        ctx, ocfg = make(capi, **kw)
        ctx.init_weights(seed + 3); blob = ctx.get_weights()
        n = 200
        st, legal, tp = common.random_stacked(ocfg, n, seed=seed)
        gid = np.arange(n, dtype=np.uint64) + 50; mv = np.ones(n, np.int32)
        vc, rv = ctx.run_mcts(st, legal, tp, True, gid, mv)
        same = np.mean([vc[i].tolist() == O.run_mcts(ocfg, blob, st[i], int(legal[i]), int(tp[i]), True, int(gid[i]), 1)[0].tolist() for i in range(n)])
        assert same >= 0.97, (same, kw)
        ctx.close()


def test_mma_rejects_wide_layers_and_resnet(capi):
    with pytest.raises(capi.MuZeroB200Error) as e:
        capi.Context(capi.default_config(num_slots=32, nn_mode=capi.NN_SPLIT_MMA, stacked_observations=2))     # 99 inputs
    assert e.value.code == capi.E_UNSUPPORTED
    with pytest.raises(capi.MuZeroB200Error):
        capi.Context(capi.resnet_config(num_slots=32, nn_mode=capi.NN_SPLIT_MMA))


def test_mma_learner_forward_and_stale_image(capi):
    """The K-step unroll on this path (mz_k_learn_forward_sp): predictions and losses against the Float32 oracle; the reference_l2 update
    does not depend on the forward pass, so the weights after learning steps are BIT-identical to the exact path's; and self-play after
    an update uses the updated weights (the tensor-core image is rebuilt on the device)."""
    ctx, ocfg = make(capi, num_slots=256, replay_buffer_size=1024, batch_size=96)
    ex = capi.Context(capi.default_config(num_slots=256, replay_buffer_size=1024, batch_size=96))
    ctx.init_weights(21); blob = ctx.get_weights(); ex.set_weights(blob)
    ex.self_play(0, 300, 1.0)
    ctx.history_import(ex.history_export())
    b = ctx.get_batch(3)
    pv, pr, pp, losses = ctx.learn_forward(b)
    opv, opr, opp, ol = O.learn_forward(ocfg, blob, b)
    assert np.max(np.abs(pv - opv)) <= MMA_ATOL and np.max(np.abs(pr - opr)) <= MMA_ATOL and np.max(np.abs(pp - opp)) <= MMA_ATOL
    assert np.allclose(losses, ol, rtol=2e-5)
    assert np.array_equal(pv[:, 0], pv[:, 1]) and np.array_equal(pp[:, 0], pp[:, 1]) and np.all(pr[:, 0] == 0)        # Q19, Learning.jl:352
    la = ctx.learn_steps(1, 5); lb = ex.learn_steps(1, 5)
    assert np.array_equal(ctx.get_weights(), ex.get_weights())            # gradient = 2 * theta (Q20): independent of the forward arithmetic
    assert np.allclose(la, lb, rtol=2e-5)
    # the search now runs on the updated weights
    w2 = ctx.get_weights()
    st, legal, tp = common.random_stacked(ocfg, 64, seed=5)
    gid = np.arange(64, dtype=np.uint64) + 7; mv = np.ones(64, np.int32)
    vc, _ = ctx.run_mcts(st, legal, tp, True, gid, mv)
    same = sum(vc[i].tolist() == O.run_mcts(ocfg, w2, st[i], int(legal[i]), int(tp[i]), True, int(gid[i]), 1)[0].tolist() for i in range(64))
    assert same >= 62, same
    h = ctx.representation(st)
    oh = np.stack([O.representation(ocfg, w2, x) for x in st])
    assert np.max(np.abs(h - oh)) <= MMA_ATOL * max(1.0, float(np.max(np.abs(oh))))
    ctx.close(); ex.close()


# ---- MZ_GRAD_BPTT on the tensor cores (mz_learner_tc.cuh): forward in split precision, backward with bf16 operands ----------------
# Stated tolerance: per network, max |g - g_oracle| <= 1e-2 * max |g_oracle| against the oracle's Float64 backward (bf16 operands carry 8
# mantissa bits; measured 1e-3 .. 3e-3).  The exact fp32 kernel (nn_mode = fp32, tests/test_gpu_parity.py: 2e-5) stays the tight check.
BPTT_TC_RTOL = 1e-2


def _tc_batch(ocfg, blob, B, seed):
    rng = np.random.default_rng(seed)
    hist = O.self_play(ocfg, blob, 0, 16, 1.0, 2)
    c2 = O.Config.from_buffer_copy(ocfg); c2.batch_size = B
    batch = O.get_batch(c2, hist, step=seed)
    batch["rewards"] = batch["rewards"] + (rng.standard_normal(batch["rewards"].shape) * 0.3).astype(np.float32)
    return batch


def _tc_check(ocfg, g, og, tol=BPTT_TC_RTOL):
    nr, npred = O.num_params(ocfg, 0), O.num_params(ocfg, 1)
    worst = 0.0
    for lo, hi in ((0, nr), (nr, nr + npred), (nr + npred, g.shape[0])):
        scale = np.max(np.abs(og[lo:hi]))
        err = np.max(np.abs(g[lo:hi].astype(np.float64) - og[lo:hi]))
        assert err <= tol * scale, (lo, hi, err, scale)
        worst = max(worst, err / scale)
    return worst


@pytest.mark.parametrize("kw,B", [({}, 32), ({"intermediate_rewards": 1}, 77), ({}, 5), ({"num_unroll_steps": 1, "intermediate_rewards": 1}, 40),
                                  ({"num_unroll_steps": 2}, 33), ({"depth_value": 0, "depth_policy": 2, "depth_reward": 0, "depth_state_head": 1, "intermediate_rewards": 1}, 64),
                                  ({"width_hidden": 48, "depth_prediction": 1, "depth_dynamics": 0, "depth_representation": 1, "intermediate_rewards": 1}, 100)])
def test_bptt_on_tensor_cores_matches_oracle(capi, kw, B):
    ctx, ocfg = make(capi, batch_size=B, **kw)
    ctx.init_weights(5)
    rng = np.random.default_rng(8)
    blob = ctx.get_weights() + (rng.standard_normal(ctx.num_params()) * 0.02).astype(np.float32)
    ctx.set_weights(blob)
    batch = _tc_batch(ocfg, blob, B, 3)
    assert ctx.learner_path(capi.GRAD_BPTT) == 2 and ctx.learner_path(capi.GRAD_REFERENCE_L2) == 1
    g, losses = ctx.learn_gradients(batch, capi.GRAD_BPTT)
    _, og = O.learn_gradients(ocfg, blob, batch, fwd64=False)
    worst = _tc_check(ocfg, g, og)
    print("tensor-core BPTT, %s, B = %d: worst per-network gradient error %.2e of the network's largest entry" % (kw, B, worst))
    # the forward of the fused kernel is the forward-only kernel's: same predictions and losses, within 2e-5 of the Float32 oracle
    pv, pr, pp, l_fwd = ctx.learn_forward(batch)
    assert np.array_equal(losses, l_fwd)
    opv, opr, opp, ol = O.learn_forward(ocfg, blob, batch)
    assert np.max(np.abs(pv - opv)) <= MMA_ATOL and np.max(np.abs(pp - opp)) <= MMA_ATOL and np.allclose(losses, ol, rtol=2e-5)
    g2, _ = ctx.learn_gradients(batch, capi.GRAD_BPTT)
    assert np.array_equal(g, g2)                                          # deterministic
    ctx.close()


def test_bptt_on_tensor_cores_large_batch_and_training(capi):
    ctx, ocfg = make(capi, batch_size=4096, replay_buffer_size=4096)
    ctx.init_weights(9); blob = ctx.get_weights()
    small = _tc_batch(ocfg, blob, 32, 4)
    big = {k: np.concatenate([v] * 128) for k, v in small.items()}
    gs, _ = ctx.learn_gradients(small, capi.GRAD_BPTT)
    gb, _ = ctx.learn_gradients(big, capi.GRAD_BPTT)                       # 128 tiles, 16 chunks: the mean gradient of 128 copies is the gradient of one
    assert np.max(np.abs(gs - gb)) <= 1e-4 * np.max(np.abs(gs))
    _, og = O.learn_gradients(ocfg, blob, small, fwd64=False)
    _tc_check(ocfg, gb, og)
    ctx.close()
    # the update is ADAM on that gradient, and training lowers the objective (Learning.jl:287)
    ctx, ocfg = make(capi, num_slots=256, replay_buffer_size=1024, batch_size=256)
    ctx.init_weights(12)
    ex = capi.Context(capi.default_config(num_slots=256, replay_buffer_size=1024, batch_size=256)); ex.set_weights(ctx.get_weights())
    ex.self_play(0, 512, 1.0); ctx.history_import(ex.history_export()); ex.close()

    def objective():
        _, _, _, l = ctx.learn_forward(ctx.get_batch(999))
        return float(np.sum(l.astype(np.float64)))
    before = objective()
    w0 = ctx.get_weights(); m = np.zeros_like(w0); v = np.zeros_like(w0)
    b1 = ctx.get_batch(1); g1, _ = ctx.learn_gradients(b1, capi.GRAD_BPTT)
    ctx.learn_step(1, capi.GRAD_BPTT, b1)
    O.adam_apply(w0, m, v, g1, 1)
    assert np.array_equal(ctx.get_weights(), w0)
    ctx.learn_steps(2, 60, capi.GRAD_BPTT)
    after = objective()
    assert np.isfinite(after) and after < 0.5 * before, (before, after)
    ctx.close()


def test_bptt_on_tensor_cores_with_per_weights(capi):
    """conf.PER on the tensor-core learner: the importance weights enter the loss and its gradient (Learning.jl:272-281); gradient against the
    oracle's weighted Float64 backward within the stated tolerance, losses within 2e-5, and the weights matter."""
    ctx, ocfg = make(capi, num_slots=64, replay_buffer_size=160, per=1, batch_size=40, intermediate_rewards=1)
    ctx.init_weights(6)
    rng = np.random.default_rng(4)
    blob = ctx.get_weights() + (rng.standard_normal(ctx.num_params()) * 0.02).astype(np.float32)
    ctx.set_weights(blob)
    ex = capi.Context(capi.default_config(num_slots=64, replay_buffer_size=160, per=1, batch_size=40, intermediate_rewards=1)); ex.set_weights(blob)
    ex.self_play(0, 100, 1.0)
    b = ex.get_batch_per(2); ex.close()
    b["rewards"] = b["rewards"] + (rng.standard_normal(b["rewards"].shape) * 0.3).astype(np.float32)
    assert ctx.learner_path(capi.GRAD_BPTT) == 2
    g, losses = ctx.learn_gradients(b, capi.GRAD_BPTT)
    _, _, _, ol = O.learn_forward_w(ocfg, blob, b)
    assert np.allclose(losses, ol, rtol=2e-5)
    _, og = O.learn_gradients_w(ocfg, blob, b, fwd64=False)
    _tc_check(ocfg, g, og)
    b1 = dict(b); b1["weights"] = np.ones_like(b["weights"])
    g1, l1 = ctx.learn_gradients(b1, capi.GRAD_BPTT)
    assert not np.array_equal(l1, losses) and not np.array_equal(g1, g)
    ctx.close()


@pytest.mark.parametrize("seed", range(5))
def test_split_one_launch_per_wave_equals_one_launch_per_move(seed):
    """mz_k_search_sp plays a whole wave in one launch (single-wave self-play; calls for more games than slots are cut into waves); the
    weight-set fill accounting then carries a per-ply bias.  Under randomly drawn network shapes (different numbers of rounds, weight
    sets and members per set) the games must be the ones the one-launch-per-move loop plays (MUZERO_B200_PERSIST=0)."""
    import os
    from muzero_jl_b200 import capi
    rng = np.random.default_rng(900 + seed)
    kw = dict(num_iters=int(rng.integers(3, 40)), stacked_observations=int(rng.integers(0, 2)), exploration_eps=float(np.float32(rng.choice([0.0, 0.25]))),
              depth_representation=int(rng.integers(0, 4)), depth_prediction=int(rng.integers(0, 4)), depth_dynamics=int(rng.integers(0, 4)),
              depth_policy=int(rng.integers(0, 3)), depth_value=int(rng.integers(0, 3)), depth_reward=int(rng.integers(0, 3)),
              depth_state_head=int(rng.integers(0, 4)), width_hidden=int(rng.choice([32, 64])), max_moves=int(rng.choice([9, 9, 5])),
              num_slots=int(rng.choice([32, 64, 100])), replay_buffer_size=1024, nn_mode=capi.NN_SPLIT_MMA, seed=int(rng.integers(1, 1 << 30)))
    games = int(rng.choice([kw["num_slots"], kw["num_slots"] // 2 + 1, 2 * kw["num_slots"] + 7]))
    temperature = float(rng.choice([0.0, 1.0]))
    res = []
    for persist in ("1", "0"):
        old = os.environ.get("MUZERO_B200_PERSIST"); os.environ["MUZERO_B200_PERSIST"] = persist
        try:
            ctx = capi.Context(capi.default_config(**kw))
        finally:
            if old is None:
                del os.environ["MUZERO_B200_PERSIST"]
            else:
                os.environ["MUZERO_B200_PERSIST"] = old
        ctx.init_weights(seed + 5)
        n0 = ctx.launch_count()
        sims, moves = ctx.self_play(100, games, temperature)
        launches = ctx.launch_count() - n0
        h = ctx.history_export(); order = np.argsort(h["game_id"])
        res.append((sims, moves, launches, {k: h[k][order] for k in common.HIST_KEYS}, ctx.replay_info()))
        ctx.close()
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1], kw
    assert res[0][2] < res[1][2], (res[0][2], res[1][2])                  # fewer launches: one search per wave instead of one per move
    for k in common.HIST_KEYS:
        assert np.array_equal(res[0][3][k], res[1][3][k]), (k, kw)
    assert res[0][4]["n_games"] == res[1][4]["n_games"] and res[0][4]["total_samples"] == res[1][4]["total_samples"]
