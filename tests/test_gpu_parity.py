"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs and against the committed golden fixtures.  Integer / index results (env
transitions, visit counts, selected actions, sampled indices) must be bit-exact; in MZ_NN_FP32_EXACT mode the
network outputs, priors, root values, histories and targets are bit-exact too (the arithmetic contract fixes the
summation order); loss scalars use a parallel reduction and are compared with rtol 2e-6."""
import ctypes as C
import os

import numpy as np
import pytest

import common
from oracle import oracle as O

pytestmark = pytest.mark.gpu
LOSS_RTOL = 2e-6


@pytest.fixture(scope="module")
def capi():
    from muzero_jl_b200 import capi
    return capi


def make_ctx(capi, **kw):
    kw.setdefault("num_slots", 256); kw.setdefault("replay_buffer_size", 1024)
    cfg = capi.default_config(**kw)
    return capi.Context(cfg), common.oracle_config(cfg)


def test_device_is_blackwell(capi):
    ctx, _ = make_ctx(capi)
    info = ctx.device_info()
    assert info["cc"][0] >= 10 and info["sm_count"] >= 100
    ctx.close()


def test_networks_bit_exact(capi):
    ctx, ocfg = make_ctx(capi)
    ctx.init_weights(1337)
    blob = ctx.get_weights()
    assert np.array_equal(blob, O.init_weights(ocfg, 1337))
    rng = np.random.default_rng(1)
    for B in (1, 31, 32, 100):
        st = rng.normal(size=(B, 63)).astype(np.float32)
        h = ctx.representation(st)
        assert np.array_equal(h, np.stack([O.representation(ocfg, blob, x) for x in st]))
        v, p = ctx.prediction(h)
        ov, op = zip(*[O.prediction(ocfg, blob, x) for x in h])
        assert np.array_equal(v, np.array(ov, np.float32)) and np.array_equal(p, np.stack(op))
        sa = rng.normal(size=(B, 36)).astype(np.float32)
        nh, r = ctx.dynamics(sa)
        oh, orr = zip(*[O.dynamics(ocfg, blob, x) for x in sa])
        assert np.array_equal(nh, np.stack(oh)) and np.array_equal(r, np.array(orr, np.float32))
    # set/get round trip per net
    w2 = rng.normal(size=ctx.num_params(capi.NET_PREDICTION)).astype(np.float32)
    ctx.set_weights(w2, capi.NET_PREDICTION)
    assert np.array_equal(ctx.get_weights(capi.NET_PREDICTION), w2)
    assert np.array_equal(ctx.get_weights(capi.NET_DYNAMICS), blob[18331 + 23242:])
    ctx.close()


def test_env_exhaustive_against_state_graph(capi):
    """Every reachable (board, legal action) pair of games/tictactoe/game.jl's rules: 6046 boards (KAT-env-3)."""
    ctx, ocfg = make_ctx(capi)
    L = O.lib()
    seen = {}
    stack = [O.Env()]; L.mzo_env_reset(C.byref(ocfg), C.byref(stack[0]))
    rows = []
    while stack:
        e = stack.pop()
        key = (e.p1, e.p2)
        if key in seen:
            continue
        seen[key] = 1
        if L.mzo_env_is_terminated(C.byref(ocfg), C.byref(e)):
            continue
        m = L.mzo_env_legal_mask(C.byref(ocfg), C.byref(e))
        for a in range(1, 10):
            if m >> (a - 1) & 1:
                n = O.Env(e.p1, e.p2, e.player, e.moves); L.mzo_env_step(C.byref(ocfg), C.byref(n), a)
                rows.append((e.p1, e.p2, e.player, a, n.p1, n.p2, n.player, L.mzo_env_reward(C.byref(ocfg), C.byref(n), e.player),
                             L.mzo_env_is_terminated(C.byref(ocfg), C.byref(n)), L.mzo_env_legal_mask(C.byref(ocfg), C.byref(n))))
                stack.append(n)
    assert len(seen) == 6046
    rows = np.array(rows, dtype=np.int64)
    p1 = rows[:, 0].astype(np.uint64); p2 = rows[:, 1].astype(np.uint64); pl = rows[:, 2].astype(np.int32)
    assert np.array_equal(ctx.env_legal(p1, p2, pl) != 0, np.ones(len(rows), bool))
    reward, done, legal = ctx.env_step(p1, p2, pl, rows[:, 3].astype(np.int32))
    assert np.array_equal(p1, rows[:, 4].astype(np.uint64)) and np.array_equal(p2, rows[:, 5].astype(np.uint64))
    assert np.array_equal(pl, rows[:, 6].astype(np.int32))
    assert np.array_equal(reward, rows[:, 7].astype(np.float32))
    assert np.array_equal(done, rows[:, 8].astype(np.int32))
    assert np.array_equal(legal, rows[:, 9].astype(np.uint32))
    obs = ctx.env_observation(p1[:64], p2[:64])
    for i in range(64):
        o = np.zeros(27, np.float32); e = O.Env(int(p1[i]), int(p2[i]), 1, 0); L.mzo_env_observation(C.byref(ocfg), C.byref(e), O._p(o))
        assert np.array_equal(obs[i], o)
    q1, q2, qp = ctx.env_reset(5)
    assert not q1.any() and not q2.any() and np.all(qp == 1)
    ctx.close()


@pytest.mark.parametrize("S,eps,tie,n", [(10, 0.0, 0, 70), (50, 0.0, 0, 300), (50, 0.25, 0, 64), (25, 0.0, 1, 33)])
def test_run_mcts_bit_exact(capi, S, eps, tie, n):
    ctx, ocfg = make_ctx(capi, num_iters=S, exploration_eps=eps, tie_mode=tie)
    ctx.init_weights(7); blob = ctx.get_weights()
    st, legal, tp = common.random_stacked(ocfg, n, seed=S + n)
    game = np.arange(n, dtype=np.uint64) + 100; move = (np.arange(n) % 9 + 1).astype(np.int32)
    vc, rv, pri = ctx.run_mcts(st, legal, tp, True, game, move, priors=True)
    for i in range(n):
        ovc, orv, opri = O.run_mcts(ocfg, blob, st[i], int(legal[i]), int(tp[i]), True, int(game[i]), int(move[i]))
        assert vc[i].tolist() == ovc.tolist(), i
        assert rv[i] == orv and np.array_equal(pri[i], opri), i
    assert np.all(vc.sum(1) == S)
    ctx.close()


def test_run_mcts_rejects_empty_legal(capi):
    ctx, _ = make_ctx(capi)
    ctx.init_weights(1)
    with pytest.raises(capi.MuZeroB200Error):
        ctx.run_mcts(np.zeros((1, 63), np.float32), [0], [1], True, [0], [1])
    ctx.close()


def test_golden_mcts(capi):
    g = np.load(os.path.join(common.ROOT, "tests", "golden", "mcts_s50.npz"))
    for eps, sfx in ((0.0, ""), (0.25, "_noise")):
        ctx, _ = make_ctx(capi, num_iters=50, exploration_eps=eps)
        ctx.init_weights(1337)
        assert abs(float(ctx.get_weights().astype(np.float64).sum()) - float(g["weight_checksum"])) == 0
        vc, rv, pri = ctx.run_mcts(g["stacked"], g["legal"], g["to_play"], True, g["game"], g["move"], priors=True)
        assert np.array_equal(vc, g["vc" + sfx]) and np.array_equal(rv, g["rv" + sfx]) and np.array_equal(pri, g["pri" + sfx])
        if not sfx:
            h = ctx.representation(g["stacked"]); v, p = ctx.prediction(h)
            assert np.array_equal(h, g["hidden"]) and np.array_equal(v, g["value"]) and np.array_equal(p, g["policy"])
        ctx.close()


def test_select_action_bit_exact(capi):
    ctx, ocfg = make_ctx(capi)
    rng = np.random.default_rng(5); n = 500
    legal = rng.integers(1, 512, n).astype(np.uint32)
    vc = (rng.integers(0, 12, (n, 9)) * ((legal[:, None] >> np.arange(9)) & 1)).astype(np.int32)
    vc[np.arange(n), [int(np.log2(int(l) & -int(l))) for l in legal]] += 1   # at least one visit
    game = rng.integers(0, 1 << 20, n).astype(np.uint64); move = rng.integers(1, 10, n).astype(np.int32)
    for T in (0.0, 1.0, 0.5, 0.25, 0.7, float("inf")):
        act = ctx.select_action(vc, legal, T, game, move)
        ref = [O.select_action(ocfg, vc[i], int(legal[i]), T, int(game[i]), int(move[i])) for i in range(n)]
        assert act.tolist() == ref, T
    ctx.close()


@pytest.mark.parametrize("temperature,eps,slots,games", [(1.0, 0.25, 64, 200), (0.0, 0.0, 96, 96), (1.0, 0.0, 32, 45)])
def test_self_play_bit_exact(capi, temperature, eps, slots, games):
    ctx, ocfg = make_ctx(capi, exploration_eps=eps, num_slots=slots, replay_buffer_size=256)
    ctx.init_weights(3); blob = ctx.get_weights()
    sims, moves = ctx.self_play(1000, games, temperature)
    o = O.self_play(ocfg, blob, 1000, games, temperature, 4)
    assert sims == o["sims"] and moves == int(o["T"].sum())
    info = ctx.replay_info()
    assert info["n_games"] == games and info["first_key"] == 1 and info["total_samples"] == int(o["T"].sum())
    h = ctx.history_export()
    assert sorted(h["game_id"].tolist()) == list(range(1000, 1000 + games))
    for j in range(games):
        i = int(h["game_id"][j]) - 1000
        for k in common.HIST_KEYS:
            assert np.array_equal(h[k][j], o[k][i]), (k, i)
    st = ctx.search_stats()
    assert 1.0 <= st["mean_depth"] <= 10 and 1.0 <= st["mean_legal"] <= 9
    ctx.close()


def test_self_play_is_independent_of_slot_count(capi):
    res = []
    for slots in (32, 160):
        ctx, _ = make_ctx(capi, num_slots=slots, replay_buffer_size=512)
        ctx.init_weights(11)
        ctx.self_play(0, 300, 1.0)
        h = ctx.history_export(); order = np.argsort(h["game_id"])
        res.append({k: h[k][order] for k in common.HIST_KEYS}); ctx.close()
    for k in common.HIST_KEYS:
        assert np.array_equal(res[0][k], res[1][k]), k


def test_replay_ring_eviction_and_counters(capi):
    ctx, ocfg = make_ctx(capi, num_slots=32, replay_buffer_size=64)
    ctx.init_weights(2)
    ctx.self_play(0, 100, 1.0)
    info = ctx.replay_info()
    assert info["n_games"] == 64 and info["first_key"] == 37       # FIFO eviction beyond replay_buffer_size (ReplayBuffer.jl:156-160)
    h = ctx.history_export()
    assert info["total_samples"] == int(h["T"].sum())
    ctx.replay_clear()
    assert ctx.replay_info()["n_games"] == 0
    ctx.close()


def test_get_batch_and_learner_against_oracle_and_golden(capi):
    g = np.load(os.path.join(common.ROOT, "tests", "golden", "selfplay_learn.npz"))
    ctx, ocfg = make_ctx(capi, num_slots=64, replay_buffer_size=64)
    ctx.init_weights(1337); blob = ctx.get_weights()
    sims, _ = ctx.self_play(0, 64, 1.0)
    assert sims == int(g["sims"])
    h = ctx.history_export(); order = np.argsort(h["game_id"])
    for k in common.HIST_KEYS:
        assert np.array_equal(h[k][order], g["hist_" + k]), k
    # the ring's key order is the save order; rebuild it in game-id order so sampled indices match the oracle's buffer
    ctx.replay_clear()
    hist = {k: g["hist_" + k] for k in common.HIST_KEYS}
    ctx.history_import(hist, game_id=np.arange(64))
    for step in (1, 2, 9):
        b = ctx.get_batch(step)
        ob = O.get_batch(ocfg, hist, step, first_key=1)
        for k in common.BATCH_KEYS:
            assert np.array_equal(b[k], ob[k]), (k, step)
    b = ctx.get_batch(1)
    for k in common.BATCH_KEYS:
        assert np.array_equal(b[k], g["batch_" + k]), k
    pv, pr, pp, losses = ctx.learn_forward(b)
    assert np.array_equal(pv, g["pred_values"]) and np.array_equal(pr, g["pred_rewards"]) and np.array_equal(pp, g["pred_policies"])
    assert np.allclose(losses, g["losses"], rtol=LOSS_RTOL, atol=0)
    l1 = ctx.learn_step(1); l2 = ctx.learn_step(2)
    assert np.allclose(l1, g["losses_step1"], rtol=LOSS_RTOL, atol=0) and np.allclose(l2, g["losses_step2"], rtol=LOSS_RTOL, atol=0)
    assert np.array_equal(ctx.get_weights(), g["weights_after_2"])     # reference_l2 update is elementwise: bit-exact
    ctx.close()


def test_learner_large_batch_properties(capi):
    """Size-independent properties at a throughput-sized batch: forward rows 0 and 1 coincide (Q19), rewards row 0 = 0,
    policies are distributions, and the per-sample result does not depend on the batch it is part of."""
    ctx, ocfg = make_ctx(capi, num_slots=256, replay_buffer_size=2048, batch_size=4096)
    ctx.init_weights(4)
    ctx.self_play(0, 1024, 1.0)
    b = ctx.get_batch(3)
    pv, pr, pp, losses = ctx.learn_forward(b)
    assert np.array_equal(pv[:, 0], pv[:, 1]) and np.array_equal(pp[:, 0], pp[:, 1]) and not pr[:, 0].any()
    assert np.allclose(pp.sum(-1), 1.0, atol=1e-5) and np.all(np.isfinite(losses))
    sub = {k: b[k][100:133] for k in ("obs", "actions", "values", "rewards", "policies", "gscale")}
    pv2, pr2, pp2, _ = ctx.learn_forward(sub)
    assert np.array_equal(pv2, pv[100:133]) and np.array_equal(pr2, pr[100:133]) and np.array_equal(pp2, pp[100:133])
    blob = ctx.get_weights()
    opv, opr, opp, ol = O.learn_forward(common.oracle_config(ctx.cfg), blob, {k: v[:64] for k, v in b.items()})
    assert np.array_equal(opv, pv[:64]) and np.array_equal(opp, pp[:64])
    ctx.close()


# ---- tensor-core network path (MZ_NN_BF16_TC): tolerance + agreement, not bit-exactness -----------------------------
TC_ATOL = 5e-3          # bf16 operands: one bf16 ulp of an O(1) activation is 4e-3


def test_tc_networks_within_bf16_tolerance(capi):
    ctx, ocfg = make_ctx(capi, nn_mode=capi.NN_BF16_TC)
    ctx.init_weights(1337); blob = ctx.get_weights()
    rng = np.random.default_rng(2)
    st, legal, tp = common.random_stacked(ocfg, 70, seed=9)
    O.set_bf16(True)
    try:
        h = ctx.representation(st)
        oh = np.stack([O.representation(ocfg, blob, x) for x in st])
        assert np.allclose(h, oh, atol=TC_ATOL) and np.median(np.abs(h - oh)) < 1e-5
        v, p = ctx.prediction(oh)
        ov, op = zip(*[O.prediction(ocfg, blob, x) for x in oh])
        assert np.allclose(v, np.array(ov), atol=TC_ATOL) and np.allclose(p, np.stack(op), atol=TC_ATOL)
        assert np.median(np.abs(p - np.stack(op))) < 1e-5
        sa = np.concatenate([oh * 2, np.full((70, 9), np.float32(5.0 / 9.0), np.float32)], axis=1)
        nh, r = ctx.dynamics(sa)
        onh, orr = zip(*[O.dynamics(ocfg, blob, x) for x in sa])
        assert np.allclose(nh, np.stack(onh), atol=TC_ATOL) and np.allclose(r, np.array(orr), atol=TC_ATOL)
        assert np.median(np.abs(nh - np.stack(onh))) < 1e-5
    finally:
        O.set_bf16(False)
    # against the reference's Float32 arithmetic: bf16 tolerance
    oh32 = np.stack([O.representation(ocfg, blob, x) for x in st])
    assert np.allclose(h, oh32, atol=5e-2)
    ctx.close()


def test_tc_run_mcts_agreement(capi):
    n, S = 256, 50
    ctx, ocfg = make_ctx(capi, num_iters=S, exploration_eps=0.0, nn_mode=capi.NN_BF16_TC)
    ctx.init_weights(7); blob = ctx.get_weights()
    st, legal, tp = common.random_stacked(ocfg, n, seed=77)
    game = np.arange(n, dtype=np.uint64) + 100; move = (np.arange(n) % 9 + 1).astype(np.int32)
    vc, rv, pri = ctx.run_mcts(st, legal, tp, True, game, move, priors=True)
    assert np.all(vc.sum(1) == S) and np.all((vc > 0) <= ((legal[:, None] >> np.arange(9)) & 1).astype(bool))
    exact = [O.run_mcts(ocfg, blob, st[i], int(legal[i]), int(tp[i]), True, int(game[i]), int(move[i])) for i in range(n)]
    O.set_bf16(True)
    try:
        emu = [O.run_mcts(ocfg, blob, st[i], int(legal[i]), int(tp[i]), True, int(game[i]), int(move[i])) for i in range(n)]
    finally:
        O.set_bf16(False)
    same_emu = np.mean([vc[i].tolist() == emu[i][0].tolist() for i in range(n)])
    same_exact = np.mean([vc[i].tolist() == exact[i][0].tolist() for i in range(n)])
    best_exact = np.mean([int(np.argmax(vc[i])) == int(np.argmax(exact[i][0])) for i in range(n)])
    print("TC visit counts identical to bf16-emulating oracle: %.3f, to the Float32 oracle: %.3f, same most-visited action: %.3f"
          % (same_emu, same_exact, best_exact))
    assert same_emu >= 0.95        # same arithmetic up to accumulation order
    assert best_exact >= 0.85      # bf16 vs Float32 networks
    assert np.allclose(pri, np.stack([e[2] for e in emu]), atol=TC_ATOL)
    assert np.allclose(rv, np.array([e[1] for e in emu]), atol=5e-2)
    ctx.close()


def test_tc_self_play_is_well_formed(capi):
    ctx, ocfg = make_ctx(capi, nn_mode=capi.NN_BF16_TC, num_slots=128, replay_buffer_size=512)
    ctx.init_weights(3)
    sims, moves = ctx.self_play(0, 300, 1.0)
    h = ctx.history_export()
    assert sims == moves * ctx.cfg.num_iters and sorted(h["game_id"].tolist()) == list(range(300))
    assert h["T"].min() >= 6 and h["T"].max() <= 9
    for j in range(300):
        T = h["T"][j]
        assert np.allclose(h["child_visits"][j, :T].sum(1), 1.0, atol=1e-6) and np.all(h["rewards"][j, :T - 1] == 0)
    ctx.close()


def test_learn_steps_equals_repeated_learn_step(capi):
    ws = []
    for fused in (False, True):
        ctx, ocfg = make_ctx(capi, num_slots=64, replay_buffer_size=128)
        ctx.init_weights(21); ctx.self_play(0, 64, 1.0)
        if fused:
            losses = ctx.learn_steps(1, 7)
        else:
            for t in range(1, 8):
                losses = ctx.learn_step(t)
        ws.append((ctx.get_weights(), losses)); ctx.close()
    assert np.array_equal(ws[0][0], ws[1][0]) and np.allclose(ws[0][1], ws[1][1], rtol=LOSS_RTOL)


# ---- grad_mode = MZ_GRAD_BPTT: gradient of the reference's loss through the unroll (mz_k_learn_bptt) -----------------
# The CUDA backward accumulates in Float32 (32-row tiles, then tiles in order); the oracle's backward is Float64 around
# the SAME Float32 forward activations (identical relu masks).  Tolerance: 2e-5 of the largest gradient entry of the
# network the parameter belongs to.
BPTT_RTOL = 2e-5


def _bptt_batch(ocfg, blob, B, seed):
    rng = np.random.default_rng(seed)
    hist = O.self_play(ocfg, blob, 0, 16, 1.0, 2)
    c2 = O.Config.from_buffer_copy(ocfg); c2.batch_size = B
    batch = O.get_batch(c2, hist, step=seed)
    batch["rewards"] = batch["rewards"] + (rng.standard_normal(batch["rewards"].shape) * 0.3).astype(np.float32)
    return batch


def _check_grad(ocfg, g, og):
    nr, npred = O.num_params(ocfg, 0), O.num_params(ocfg, 1)
    for lo, hi in ((0, nr), (nr, nr + npred), (nr + npred, g.shape[0])):
        scale = np.max(np.abs(og[lo:hi]))
        err = np.max(np.abs(g[lo:hi].astype(np.float64) - og[lo:hi]))
        assert err <= BPTT_RTOL * scale, (lo, hi, err, scale)


@pytest.mark.parametrize("kw,B", [({}, 32), ({"intermediate_rewards": 1}, 32), ({"intermediate_rewards": 1}, 77), ({}, 5),
                                  ({"num_unroll_steps": 1, "intermediate_rewards": 1}, 40), ({"num_unroll_steps": 2}, 33),   # (K = 0 is degenerate in the reference: gradient_scale = min(K, ..) = 0, ReplayBuffer.jl:212)
                                  ({"depth_value": 0, "depth_policy": 2, "depth_reward": 0, "depth_state_head": 1, "intermediate_rewards": 1}, 64),
                                  ({"stacked_observations": 2}, 32), ({"stacked_observations": 2, "intermediate_rewards": 1, "num_unroll_steps": 3}, 45)])   # 99 inputs: wider than the 64-row buffers
def test_bptt_gradients_match_oracle(capi, kw, B):
    ctx, ocfg = make_ctx(capi, batch_size=B, **kw)
    ctx.init_weights(5)
    rng = np.random.default_rng(8)
    blob = ctx.get_weights() + (rng.standard_normal(ctx.num_params()) * 0.02).astype(np.float32)   # non-zero biases
    ctx.set_weights(blob)
    batch = _bptt_batch(ocfg, blob, B, 3)
    g, losses = ctx.learn_gradients(batch, capi.GRAD_BPTT)
    _, og = O.learn_gradients(ocfg, blob, batch, fwd64=False)
    _check_grad(ocfg, g, og)
    # the fused kernel's forward is the forward kernel's arithmetic: predictions and losses are bit-identical
    pv, pr, pp, l_fwd = ctx.learn_forward(batch)
    assert np.array_equal(losses, l_fwd)
    opv, opr, opp, ol = O.learn_forward(ocfg, blob, batch)
    assert np.array_equal(pv, opv) and np.array_equal(pp, opp) and np.allclose(losses, ol, rtol=LOSS_RTOL)
    # deterministic: same bits on a second run
    g2, _ = ctx.learn_gradients(batch, capi.GRAD_BPTT)
    assert np.array_equal(g, g2)
    # reference_l2 mode through the same entry point: exactly 2*theta (Q20)
    g3, _ = ctx.learn_gradients(batch, capi.GRAD_REFERENCE_L2)
    assert np.array_equal(g3, blob + blob)
    ctx.close()


def test_bptt_refuses_layers_wider_than_its_backward_tile(capi):
    """the backward tile covers 64 outputs / 64 inputs with an input gradient: a 72-wide hidden layer is refused (not silently truncated),
    while everything else of such a context (forward, the reference's own update) works"""
    ctx, ocfg = make_ctx(capi, batch_size=16, width_hidden=72)
    ctx.init_weights(5); blob = ctx.get_weights()
    batch = _bptt_batch(ocfg, blob, 16, 3)
    with pytest.raises(capi.MuZeroB200Error) as e:
        ctx.learn_gradients(batch, capi.GRAD_BPTT)
    assert e.value.code == capi.E_UNSUPPORTED
    pv, pr, pp, _ = ctx.learn_forward(batch)
    opv, opr, opp, _ = O.learn_forward(ocfg, blob, batch)
    assert np.array_equal(pv, opv) and np.array_equal(pp, opp)
    g, _ = ctx.learn_gradients(batch, capi.GRAD_REFERENCE_L2)
    assert np.array_equal(g, blob + blob)
    ctx.close()


def test_bptt_update_is_adam_on_the_gradient(capi):
    ctx, ocfg = make_ctx(capi, batch_size=48, intermediate_rewards=1)
    ctx.init_weights(6); blob = ctx.get_weights()
    m = np.zeros_like(blob); v = np.zeros_like(blob); ob = blob.copy()
    for t in (1, 2, 3):
        batch = _bptt_batch(ocfg, ob, 48, 10 + t)
        g, _ = ctx.learn_gradients(batch, capi.GRAD_BPTT)
        ctx.learn_step(t, capi.GRAD_BPTT, batch)
        O.adam_apply(ob, m, v, g, t)                      # Flux.ADAM on the CUDA gradient: the update itself is bit-exact
        assert np.array_equal(ctx.get_weights(), ob)
    ctx.close()


def test_bptt_large_batch_properties(capi):
    """Size-independent properties at a throughput-sized batch (B = 4096, 128 tiles): a batch made of 128 copies of a
    32-sample batch has the same mean gradient as the 32-sample batch (the loss is a batch mean and Q21's mean_i(1/g_i)
    is unchanged by replication), and the training loop decreases the data loss."""
    ctx, ocfg = make_ctx(capi, batch_size=4096, replay_buffer_size=4096)
    ctx.init_weights(9); blob = ctx.get_weights()
    small = _bptt_batch(ocfg, blob, 32, 4)
    big = {k: np.concatenate([v] * 128) for k, v in small.items()}
    gs, _ = ctx.learn_gradients(small, capi.GRAD_BPTT)
    gb, _ = ctx.learn_gradients(big, capi.GRAD_BPTT)
    assert np.max(np.abs(gs - gb)) <= 1e-4 * np.max(np.abs(gs))
    _, og = O.learn_gradients(ocfg, blob, small, fwd64=False)
    _check_grad(ocfg, gb, og)
    ctx.close()


def test_bptt_training_reduces_the_loss(capi):
    """The objective the reference's optimiser sees (data loss + sum(theta^2) per net, Learning.jl:287) goes down."""
    ctx, ocfg = make_ctx(capi, num_slots=256, replay_buffer_size=1024, batch_size=256)
    ctx.init_weights(12)
    ctx.self_play(0, 512, 1.0)

    def objective():
        _, _, _, l = ctx.learn_forward(ctx.get_batch(999))
        return float(np.sum(l.astype(np.float64)))
    before = objective()
    ctx.learn_steps(1, 60, capi.GRAD_BPTT)
    after = objective()
    assert np.isfinite(after) and after < 0.5 * before, (before, after)
    ctx.close()


def test_reanalyse_bit_exact_and_used_by_targets(capi):
    """reanalysed_predicted_root_values (Constructors.jl:13): produced on the device for a range of games, bit-exact vs
    prediction(representation(stacked observation)) of the oracle, and consumed by compute_target_value (ReplayBuffer.jl:8)."""
    ctx, ocfg = make_ctx(capi, num_slots=64, replay_buffer_size=128, batch_size=64)
    ctx.init_weights(31); blob = ctx.get_weights()
    ctx.self_play(0, 96, 1.0)
    h = ctx.history_export()
    n = len(h["T"])
    v0, f0 = ctx.reanalysed_export()
    assert not f0.any()
    b_before = ctx.get_batch(5)
    ctx.reanalyse(key0=17, n=40)                                # a sub-range: keys 17..56
    vals, flags = ctx.reanalysed_export()
    assert flags.sum() == 40 and flags[16:56].all()
    L = O.lib()
    s = O.sizes(ocfg)
    hist2 = {k: v.copy() for k, v in h.items()}
    for j in range(16, 56):
        T = h["T"][j]
        for t in range(T):
            st = np.zeros(s["stack"], np.float32)
            L.mzo_stack_observations(C.byref(ocfg), O._p(np.ascontiguousarray(h["obs"][j])), O._p(np.ascontiguousarray(h["actions"][j]), C.c_int32), t + 1, O._p(st))
            v, _ = O.prediction(ocfg, blob, O.representation(ocfg, blob, st))
            assert vals[j, t] == v, (j, t)
            hist2["root_values"][j, t] = v
        assert not vals[j, T:].any()
    b_after = ctx.get_batch(5)
    ob = O.get_batch(ocfg, hist2, step=5)
    for k in common.BATCH_KEYS:
        assert np.array_equal(b_after[k], ob[k]), k
    assert not np.array_equal(b_after["values"], b_before["values"])     # the bootstrap values changed for the reanalysed games
    # a newly saved game clears the flag of the ring slot it takes
    ctx.self_play(96, 64, 1.0)
    _, f2 = ctx.reanalysed_export()
    assert f2.sum() < 40
    ctx.close()


@pytest.mark.parametrize("mode", ["l2", "bptt"])
def test_checkpoint_resume_is_bit_identical(capi, mode, tmp_path):
    """weights + ADAM moments + step count + replay histories: a restored context continues exactly like the original."""
    gm = capi.GRAD_BPTT if mode == "bptt" else capi.GRAD_REFERENCE_L2
    ctx, ocfg = make_ctx(capi, num_slots=64, replay_buffer_size=128, batch_size=48)
    ctx.init_weights(41); ctx.self_play(0, 96, 1.0)
    ctx.learn_steps(1, 4, gm)
    ck = ctx.checkpoint()
    assert ck["steps_done"] == 4 and ck["adam_m"].any() and ck["adam_v"].any()
    np.savez(tmp_path / "ck.npz", **ck)
    la = ctx.learn_steps(5, 3, gm); wa = ctx.get_weights(); ctx.close()
    ctx2, _ = make_ctx(capi, num_slots=64, replay_buffer_size=128, batch_size=48)
    with np.load(tmp_path / "ck.npz") as z:
        ctx2.restore({k: z[k] for k in z.files})
    lb = ctx2.learn_steps(5, 3, gm); wb = ctx2.get_weights()
    assert np.array_equal(wa, wb) and np.array_equal(la, lb)
    ck2 = ctx2.checkpoint()
    assert ck2["steps_done"] == 7
    ctx2.close()


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_randomised_configurations_bit_exact(capi, seed):
    """self-play -> get_batch -> one learner step under randomly drawn hyper-parameters (network depths and widths of the state,
    stacking depth, unroll / td steps, PUCT constants, discount, noise, temperature): every GameHistory, the batch and the updated
    weights are bit-identical to the oracle."""
    rng = np.random.default_rng(100 + seed)
    kw = dict(num_iters=int(rng.integers(3, 41)), stacked_observations=int(rng.integers(0, 3)), num_unroll_steps=int(rng.integers(1, 6)),
              td_steps=int(rng.integers(1, 8)), batch_size=int(rng.integers(1, 70)), pb_c_base=int(rng.integers(50, 30000)),
              pb_c_init=float(np.float32(rng.uniform(0.5, 2.5))), discount=float(np.float32(rng.uniform(0.8, 1.0))),
              dirichlet_alpha=float(np.float32(rng.uniform(0.1, 1.0))), exploration_eps=float(np.float32(rng.choice([0.0, 0.25, 0.5]))),
              depth_representation=int(rng.integers(0, 4)), depth_prediction=int(rng.integers(0, 4)), depth_dynamics=int(rng.integers(0, 4)),
              depth_policy=int(rng.integers(0, 3)), depth_value=int(rng.integers(0, 3)), depth_reward=int(rng.integers(0, 3)),
              depth_state_head=int(rng.integers(0, 4)), width_hidden=int(rng.choice([32, 48, 64])),
              intermediate_rewards=int(rng.integers(0, 2)), seed=int(rng.integers(1, 1 << 30)), num_slots=int(rng.integers(8, 100)))
    temperature = float(rng.choice([0.0, 0.5, 1.0]))
    games = int(rng.integers(20, 120))
    ctx, ocfg = make_ctx(capi, replay_buffer_size=512, **kw)
    ctx.init_weights(seed + 40); blob = ctx.get_weights()
    sims, moves = ctx.self_play(10, games, temperature)
    o = O.self_play(ocfg, blob, 10, games, temperature, 4)
    assert sims == o["sims"] and moves == int(o["T"].sum()), kw
    h = ctx.history_export()
    for j in range(games):
        i = int(h["game_id"][j]) - 10
        for k in common.HIST_KEYS:
            assert np.array_equal(h[k][j], o[k][i]), (k, i, kw)
    # rebuild the buffer in game-id order (the ring's key order is the save order) and compare batch + learner step
    ctx.replay_clear()
    hist = {k: o[k] for k in common.HIST_KEYS}
    ctx.history_import(hist, game_id=np.arange(10, 10 + games))
    b = ctx.get_batch(3); ob = O.get_batch(ocfg, hist, 3, first_key=1)
    for k in ("index", "obs", "actions", "values", "rewards", "policies", "gscale"):
        assert np.array_equal(b[k], ob[k]), (k, kw)
    pv, pr, pp, losses = ctx.learn_forward(b)                      # unroll forward + loss (Learning.jl:347-374, 261-288)
    opv, opr, opp, ol = O.learn_forward(ocfg, blob, ob)
    assert np.array_equal(pv, opv) and np.array_equal(pr, opr) and np.array_equal(pp, opp), kw
    assert np.allclose(losses, ol, rtol=LOSS_RTOL, atol=0), kw
    try:                                                           # gradient through the unroll vs the oracle's Float64 backward
        g, _ = ctx.learn_gradients(b, capi.GRAD_BPTT)
    except capi.MuZeroB200Error as e:                              # the saved activations of very deep / wide draws exceed one CTA's shared memory
        assert e.code == capi.E_UNSUPPORTED and "shared memory" in str(e), kw
    else:
        _, og = O.learn_gradients(ocfg, blob, ob, fwd64=False)
        assert np.max(np.abs(g - og)) <= 2e-5 * max(np.max(np.abs(og)), 1e-30), kw
    ctx.close()
