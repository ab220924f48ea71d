"""mz_k_search_lat (mz_kernels_lat.cuh): the low-latency search kernel for few roots -- one tree per two-CTA cluster, networks, tree and
PUCT table resident in shared memory -- against the Float32 oracle.  It computes in the exact arithmetic contract, so everything is
bit-exact: visit counts, priors, root values (run_mcts, SelfPlay.jl:230-285) and whole GameHistories (play_game, :330-382)."""
import os

import numpy as np
import pytest

import common
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from muzero_jl_b200 import capi
    return capi


def make_ctx(capi, lat=None, **kw):
    kw.setdefault("num_slots", 256); kw.setdefault("replay_buffer_size", 1024)
    cfg = capi.default_config(**kw)
    old = os.environ.get("MUZERO_B200_LAT")
    if lat is not None:
        os.environ["MUZERO_B200_LAT"] = str(lat)
    try:
        ctx = capi.Context(cfg)
    finally:
        if lat is not None:
            if old is None:
                del os.environ["MUZERO_B200_LAT"]
            else:
                os.environ["MUZERO_B200_LAT"] = old
    return ctx, common.oracle_config(cfg)


@pytest.mark.parametrize("S,eps,tie,n,kw", [(50, 0.25, 0, 1, {}), (50, 0.0, 0, 7, {}), (25, 0.25, 1, 148, {}), (10, 0.25, 0, 40, {"stacked_observations": 0}),
                                            (30, 0.25, 0, 20, {"stacked_observations": 2}), (50, 0.25, 0, 12, {"nn_mode": "split"})])
def test_lat_run_mcts_bit_exact(capi, S, eps, tie, n, kw):
    kw = dict(kw)
    if kw.get("nn_mode") == "split":
        kw["nn_mode"] = capi.NN_SPLIT_MMA          # small calls of a split-precision context run the exact kernel
    ctx, ocfg = make_ctx(capi, num_iters=S, exploration_eps=eps, tie_mode=tie, **kw)
    ctx.init_weights(7); blob = ctx.get_weights()
    st, legal, tp = common.random_stacked(ocfg, n, seed=S + n)
    game = np.arange(n, dtype=np.uint64) + 100; move = (np.arange(n) % 9 + 1).astype(np.int32)
    before = ctx.launch_count()
    vc, rv, pri = ctx.run_mcts(st, legal, tp, True, game, move, priors=True)
    assert ctx.launch_count() <= before + 3            # the search (+ the weight-image rebuilds after init_weights)
    for i in range(n):
        ovc, orv, opri = O.run_mcts(ocfg, blob, st[i], int(legal[i]), int(tp[i]), True, int(game[i]), int(move[i]))
        assert vc[i].tolist() == ovc.tolist(), i
        assert rv[i] == orv and np.array_equal(pri[i], opri), i
    assert np.all(vc.sum(1) == S)
    ctx.close()


def test_lat_equals_batched_kernel(capi):
    """the same roots through mz_k_search (MUZERO_B200_LAT=0) and mz_k_search_lat"""
    res = []
    for lat in (0, 200):
        ctx, ocfg = make_ctx(capi, lat=lat, num_iters=50)
        ctx.init_weights(21)
        st, legal, tp = common.random_stacked(ocfg, 100, seed=9)
        game = np.arange(100, dtype=np.uint64); move = np.ones(100, np.int32)
        res.append(ctx.run_mcts(st, legal, tp, True, game, move, priors=True)); ctx.close()
    for x, y in zip(res[0], res[1]):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("temperature,eps,slots,games,kw", [(1.0, 0.25, 1, 5, {}), (0.5, 0.25, 16, 100, {}), (0.0, 0.0, 7, 30, {"temperature_threshold": 3}),
                                                            (1.0, 0.25, 64, 200, {"stacked_observations": 2, "num_iters": 20})])
def test_lat_self_play_bit_exact(capi, temperature, eps, slots, games, kw):
    ctx, ocfg = make_ctx(capi, lat=64, exploration_eps=eps, num_slots=slots, replay_buffer_size=256, **kw)
    ctx.init_weights(3); blob = ctx.get_weights()
    sims, moves = ctx.self_play(1000, games, temperature)
    o = O.self_play(ocfg, blob, 1000, games, temperature, 4)
    assert sims == o["sims"] and moves == int(o["T"].sum())
    h = ctx.history_export()
    assert sorted(h["game_id"].tolist()) == list(range(1000, 1000 + games))
    for j in range(games):
        i = int(h["game_id"][j]) - 1000
        for k in common.HIST_KEYS:
            assert np.array_equal(h[k][j], o[k][i]), (k, i)
    ctx.close()


def test_lat_arena_bit_exact(capi):
    ctx, ocfg = make_ctx(capi, lat=64, num_slots=8, replay_buffer_size=64, num_iters=20)
    ctx.init_weights(5); blob = ctx.get_weights()
    ar = ctx.arena(50, 24, capi.OPP_EXPERT, 2, 0.0)
    oa = O.arena(ocfg, blob, 50, 24, O.OPP_EXPERT, 2, 0.0, 4)["outcome"]
    assert (ar["wins"], ar["draws"], ar["losses"]) == (int((oa == 1).sum()), int((oa == 0).sum()), int((oa == -1).sum()))
    ctx.close()


@pytest.mark.parametrize("mode", ["exact", "split"])
def test_lat_few_games_on_a_large_context(capi, mode):
    """play_game one game at a time on an engine built for 4096 concurrent games: the call goes to mz_k_search_lat (exact fp32 in both modes)"""
    import time
    ctx, ocfg = make_ctx(capi, num_slots=4096, replay_buffer_size=8192, num_iters=50, exploration_eps=0.25,
                         nn_mode=capi.NN_SPLIT_MMA if mode == "split" else capi.NN_FP32_EXACT)
    ctx.init_weights(12); blob = ctx.get_weights()
    ctx.self_play(0, 1, 1.0)                                   # warm-up (weight image)
    t0 = time.perf_counter(); sims, moves = ctx.self_play(500, 3, 1.0); dt = time.perf_counter() - t0
    o = O.self_play(ocfg, blob, 500, 3, 1.0, 1)
    assert sims == o["sims"] and moves == int(o["T"].sum())
    h = ctx.history_export()
    sel = [int(np.flatnonzero(h["game_id"] == 500 + i)[0]) for i in range(3)]
    for k in common.HIST_KEYS:
        assert np.array_equal(h[k][sel], o[k]), k
    print("3 games on a 4096-slot %s context: %.2f ms (%d plies)" % (mode, dt * 1e3, int(o["T"].max())))
    ctx.close()


@pytest.mark.parametrize("seed", range(6))
def test_lat_randomised_configurations_bit_exact(capi, seed):
    """randomly drawn network shapes (depth 0 heads and trunks, widths 32 / 48 / 64, stacking depth 0..2), PUCT constants and simulation
    counts: run_mcts with a few roots and a few whole games through mz_k_search_lat, bit-identical to the oracle"""
    rng = np.random.default_rng(700 + seed)
    kw = dict(num_iters=int(rng.integers(3, 60)), stacked_observations=int(rng.integers(0, 3)), pb_c_base=int(rng.integers(50, 30000)),
              pb_c_init=float(np.float32(rng.uniform(0.5, 2.5))), discount=float(np.float32(rng.uniform(0.8, 1.0))),
              dirichlet_alpha=float(np.float32(rng.uniform(0.1, 1.0))), exploration_eps=float(np.float32(rng.choice([0.0, 0.25, 0.5]))),
              depth_representation=int(rng.integers(0, 4)), depth_prediction=int(rng.integers(0, 4)), depth_dynamics=int(rng.integers(0, 4)),
              depth_policy=int(rng.integers(0, 3)), depth_value=int(rng.integers(0, 3)), depth_reward=int(rng.integers(0, 3)),
              depth_state_head=int(rng.integers(0, 4)), width_hidden=int(rng.choice([32, 48, 64])), seed=int(rng.integers(1, 1 << 30)),
              tie_mode=int(rng.integers(0, 2)), num_slots=int(rng.integers(1, 40)))
    ctx, ocfg = make_ctx(capi, replay_buffer_size=256, **kw)
    ctx.init_weights(seed + 90); blob = ctx.get_weights()
    n = int(rng.integers(1, 30))
    st, legal, tp = common.random_stacked(ocfg, n, seed=seed)
    game = np.arange(n, dtype=np.uint64) + 3; move = (np.arange(n) % 9 + 1).astype(np.int32)
    before = ctx.launch_count()
    vc, rv, pri = ctx.run_mcts(st, legal, tp, True, game, move, priors=True)
    assert ctx.launch_count() <= before + 2, kw                          # one search launch (+ the weight image): not chunks of num_slots roots
    for i in range(n):
        ovc, orv, opri = O.run_mcts(ocfg, blob, st[i], int(legal[i]), int(tp[i]), True, int(game[i]), int(move[i]))
        assert vc[i].tolist() == ovc.tolist() and rv[i] == orv and np.array_equal(pri[i], opri), (i, kw)
    games = int(rng.integers(1, 17)); temperature = float(rng.choice([0.0, 0.5, 1.0]))
    sims, moves = ctx.self_play(50, games, temperature)
    o = O.self_play(ocfg, blob, 50, games, temperature, 2)
    assert sims == o["sims"] and moves == int(o["T"].sum()), kw
    h = ctx.history_export()
    for j in range(games):
        i = int(h["game_id"][j]) - 50
        for k in common.HIST_KEYS:
            assert np.array_equal(h[k][j], o[k][i]), (k, i, kw)
    ctx.close()


def test_lat_weight_image_follows_the_learner(capi):
    """the kernel's own weight image is rebuilt on the device after an ADAM step / set_weights before the next small call"""
    ctx, ocfg = make_ctx(capi, num_slots=64, replay_buffer_size=128, num_iters=20, exploration_eps=0.25)
    ctx.init_weights(31)
    st, legal, tp = common.random_stacked(ocfg, 6, seed=2)
    game = np.arange(6, dtype=np.uint64); move = np.ones(6, np.int32)
    ctx.run_mcts(st, legal, tp, True, game, move)                  # builds the image for the initial weights
    ctx.self_play(0, 64, 1.0); ctx.learn_step(1); ctx.learn_step(2)
    blob = ctx.get_weights()
    vc, rv = ctx.run_mcts(st, legal, tp, True, game, move)
    for i in range(6):
        ovc, orv, _ = O.run_mcts(ocfg, blob, st[i], int(legal[i]), int(tp[i]), True, int(game[i]), 1)
        assert vc[i].tolist() == ovc.tolist() and rv[i] == orv, i
    blob2 = O.init_weights(ocfg, 77); ctx.set_weights(blob2)
    vc, rv = ctx.run_mcts(st, legal, tp, True, game, move)
    for i in range(6):
        ovc, orv, _ = O.run_mcts(ocfg, blob2, st[i], int(legal[i]), int(tp[i]), True, int(game[i]), 1)
        assert vc[i].tolist() == ovc.tolist() and rv[i] == orv, i
    ctx.close()
