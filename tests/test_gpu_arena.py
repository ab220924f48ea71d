"""GPU: competitive play (mz_arena, mz_opponent_action; src/SelfPlay.jl:421-435, 311-325) through the C ABI against the oracle.
Exact fp32 path: wins / draws / losses and every exported GameHistory are bit-identical to oracle.arena."""
import ctypes as C

import numpy as np
import pytest

import common
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from muzero_jl_b200 import capi
    return capi


def make_ctx(capi, **kw):
    kw.setdefault("num_slots", 64); kw.setdefault("replay_buffer_size", 1024); kw.setdefault("num_iters", 20)
    cfg = capi.default_config(**kw)
    return capi.Context(cfg), common.oracle_config(cfg)


@pytest.mark.parametrize("opponent", [1, 2])
@pytest.mark.parametrize("muzero_player", [1, 2])
@pytest.mark.parametrize("temperature", [0.0, 1.0])
def test_arena_bit_exact(capi, opponent, muzero_player, temperature):
    ctx, ocfg = make_ctx(capi)
    ctx.init_weights(21); blob = ctx.get_weights()
    n = 150                                                   # more games than slots: refills interleave opponent and MuZero plies
    r = ctx.arena(500, n, opponent, muzero_player, temperature)
    o = O.arena(ocfg, blob, 500, n, opponent, muzero_player, temperature, 4)
    oc = o["outcome"]
    assert (r["wins"], r["draws"], r["losses"]) == (int((oc == 1).sum()), int((oc == 0).sum()), int((oc == -1).sum()))
    assert r["wins"] + r["draws"] + r["losses"] == n and r["simulations"] == o["sims"]
    h = ctx.history_export()
    assert sorted(h["game_id"].tolist()) == list(range(500, 500 + n))
    for j in range(n):
        i = int(h["game_id"][j]) - 500
        for k in common.HIST_KEYS:
            assert np.array_equal(h[k][j], o[k][i]), (k, i)
    ctx.close()


def test_self_play_unchanged_after_an_arena_call(capi):
    ctx, ocfg = make_ctx(capi)
    ctx.init_weights(4); blob = ctx.get_weights()
    ctx.arena(0, 40, capi.OPP_RANDOM, 2, 0.0)
    ctx.replay_clear()
    ctx.self_play(1000, 40, 1.0)
    o = O.self_play(ocfg, blob, 1000, 40, 1.0, 4)
    h = ctx.history_export()
    for j in range(40):
        i = int(h["game_id"][j]) - 1000
        for k in common.HIST_KEYS:
            assert np.array_equal(h[k][j], o[k][i]), (k, i)
    ctx.close()


def test_opponent_action_matches_oracle(capi):
    ctx, ocfg = make_ctx(capi)
    rng = np.random.default_rng(3)
    p1s, p2s, pls, gids, mvs, envs = [], [], [], [], [], []
    while len(p1s) < 400:
        e = O.Env(); O.lib().mzo_env_reset(C.byref(ocfg), C.byref(e)); acts = []
        for ply in range(int(rng.integers(0, 8))):
            legal = O.lib().mzo_env_legal_mask(C.byref(ocfg), C.byref(e))
            if legal == 0 or O.lib().mzo_env_is_terminated(C.byref(ocfg), C.byref(e)):
                break
            a = int(rng.choice([i + 1 for i in range(9) if (legal >> i) & 1]))
            O.lib().mzo_env_step(C.byref(ocfg), C.byref(e), a); acts.append(a)
        if O.lib().mzo_env_is_terminated(C.byref(ocfg), C.byref(e)):
            continue
        p1s.append(e.p1); p2s.append(e.p2); pls.append(e.player); gids.append(len(p1s)); mvs.append(len(acts) + 1); envs.append(e)
    for opp in (capi.OPP_RANDOM, capi.OPP_EXPERT):
        got = ctx.opponent_action(p1s, p2s, pls, opp, gids, mvs)
        want = [O.lib().mzo_opponent_action(C.byref(ocfg), C.byref(e), opp, C.c_uint64(g), m) for e, g, m in zip(envs, gids, mvs)]
        assert got.tolist() == want
    ctx.close()


def test_odd_slot_count(capi):
    """an odd number of slots (the save/refill scratch keeps a 64-bit key behind the slot list)"""
    ctx, ocfg = make_ctx(capi, num_slots=37)
    ctx.init_weights(8); blob = ctx.get_weights()
    sims, _ = ctx.self_play(0, 90, 1.0)
    o = O.self_play(ocfg, blob, 0, 90, 1.0, 4)
    h = ctx.history_export()
    assert sims == o["sims"]
    for j in range(90):
        for k in common.HIST_KEYS:
            assert np.array_equal(h[k][j], o[k][int(h["game_id"][j])]), k
    ctx.close()


def test_arena_argument_errors(capi):
    ctx, _ = make_ctx(capi)
    ctx.init_weights(1)
    assert ctx.arena(0, 0, capi.OPP_RANDOM, 1, 0.0) == dict(wins=0, draws=0, losses=0, simulations=0)      # empty input
    assert ctx.opponent_action([], [], [], capi.OPP_RANDOM, [], []).size == 0
    for args in ((0, 4, 7, 1, 0.0), (0, 4, capi.OPP_RANDOM, 3, 0.0), (0, 4, capi.OPP_RANDOM, 0, 0.0)):
        with pytest.raises(capi.MuZeroB200Error) as e:
            ctx.arena(*args)
        assert e.value.code == capi.E_ARG
    ctx.close()


def test_competitive_play_mirror_and_networks(capi):
    """api.competitive_play on the three network paths: tallies are complete, and the tensor-core FC path agrees with itself across slot counts."""
    import muzero_jl_b200 as mz
    eng = mz.Engine(mz.Config(num_iters=10), mz.FeedForwardHP(), num_slots=64)
    mz.init_networks(eng)
    r = mz.competitive_play(eng, 50, opponent="random", muzero_player=1)
    assert r["wins"] + r["draws"] + r["losses"] == 50 and r["simulations"] > 0
    g = mz.play_game(eng, 0.0, False, "expert", 2)
    assert len(g.action_history) >= 5 and g.to_play_history[0] == 1
    with pytest.raises(ValueError):
        mz.competitive_play(eng, 1, opponent="nobody")
    reng = mz.Engine(mz.Config(num_iters=6), mz.ResNetHP(), num_slots=28)                 # the reference's ResNetHP through the mirror
    mz.init_networks(reng)
    sims, moves = mz.self_play(reng, 40, 1.0)
    assert sims == 6 * moves
    rr = mz.competitive_play(reng, 20, opponent="expert", muzero_player=2)
    assert rr["wins"] + rr["draws"] + rr["losses"] == 20
    res = []
    for slots in (32, 96):
        ctx = capi.Context(capi.default_config(num_slots=slots, num_iters=10, nn_mode=capi.NN_BF16_TC)); ctx.init_weights(2)
        res.append(ctx.arena(0, 80, capi.OPP_EXPERT, 1, 0.0)); ctx.close()
    assert res[0] == res[1]
    ctx = capi.Context(capi.connect_config(num_slots=24, num_iters=8)); ctx.init_weights(2)
    r = ctx.arena(0, 30, capi.OPP_RANDOM, 2, 0.0)
    assert r["wins"] + r["draws"] + r["losses"] == 30
    ctx.close()
