"""CPU: competitive play (src/SelfPlay.jl:421-435, select_opponent_action :311-325) in the oracle, and the product's scalar
opponent / outcome code (tests/host_harness.cpp) against it.  The reference's "random" branch reads an undefined variable and
expert_agent() does not exist, so these pin the repaired specification (oracle/mz_oracle.c: mzo_opponent_action)."""
import ctypes as C

import numpy as np

import common
from oracle import oracle as O


def bits(cells):
    return sum(1 << (c - 1) for c in cells)


def test_expert_wins_then_blocks_then_random():
    cfg = O.default_config()
    # player 2 to move; player 1 holds 1,2 -> block at 3
    assert O.opponent_action(cfg, bits([1, 2]), bits([4]), 2, O.OPP_EXPERT, 5, 3) == 3
    # player 2 holds 4,5 and can win at 6: winning comes before blocking 3
    assert O.opponent_action(cfg, bits([1, 2, 7]), bits([4, 5]), 2, O.OPP_EXPERT, 5, 4) == 6
    # nothing to win or block: the same action as the random opponent
    for g in range(20):
        a = O.opponent_action(cfg, bits([5]), 0, 2, O.OPP_EXPERT, g, 2)
        assert a == O.opponent_action(cfg, bits([5]), 0, 2, O.OPP_RANDOM, g, 2) and a != 5


def test_random_opponent_is_uniform_over_legal_and_keyed():
    cfg = O.default_config()
    p1, p2 = bits([1, 5]), bits([2])
    counts = np.zeros(10, int)
    for g in range(6000):
        counts[O.opponent_action(cfg, p1, p2, 2, O.OPP_RANDOM, g, 4)] += 1
    legal = [3, 4, 6, 7, 8, 9]
    assert counts[[0, 1, 2, 5]].sum() == 0
    assert np.all(np.abs(counts[legal] - 1000) < 150)                    # rand(rng, las), :320
    assert O.opponent_action(cfg, p1, p2, 2, O.OPP_RANDOM, 7, 4) == O.opponent_action(cfg, p1, p2, 2, O.OPP_RANDOM, 7, 4)
    cfg2 = O.default_config(); cfg2.seed = 99
    assert any(O.opponent_action(cfg, p1, p2, 2, O.OPP_RANDOM, g, 4) != O.opponent_action(cfg2, p1, p2, 2, O.OPP_RANDOM, g, 4) for g in range(32))


def test_outcome_is_the_first_completed_line():
    cfg = O.default_config()
    f = lambda acts, mp: O.lib().mzo_arena_outcome(C.byref(cfg), len(acts), np.asarray(acts, np.int32).ctypes.data_as(C.POINTER(C.c_int32)), mp)
    # SURVEY KAT-env-1: player 1 completes 1,4,7 on ply 5, the game runs one more ply (Q14)
    assert f([1, 2, 4, 5, 7, 3], 1) == 1 and f([1, 2, 4, 5, 7, 3], 2) == -1
    # ... even if that extra ply completes a line for player 2 (2,5,8): the first line wins
    assert f([1, 2, 4, 5, 7, 8], 1) == 1
    assert f([1, 2, 3, 5, 4, 6, 8, 7, 9], 1) == 0                        # full board, no line: draw
    assert f([], 1) == 0


def test_arena_histories_follow_play_game():
    cfg = O.default_config(); cfg.num_iters = 12
    blob = O.init_weights(cfg, 5)
    for opp in (O.OPP_RANDOM, O.OPP_EXPERT):
        for mp in (1, 2):
            r = O.arena(cfg, blob, 100, 40, opp, mp, 0.0, 4)
            muzero_plies = 0
            for g in range(40):
                T = int(r["T"][g]); tp = r["to_play"][g, :T]
                assert tp.tolist() == [1 + (i % 2) for i in range(T)]
                assert r["outcome"][g] == O.lib().mzo_arena_outcome(C.byref(cfg), T, r["actions"][g].ctypes.data_as(C.POINTER(C.c_int32)), mp)
                for i in range(T):
                    if tp[i] == mp:
                        muzero_plies += 1
                        assert abs(r["child_visits"][g, i].sum() - 1.0) < 1e-5
                    elif i == 0:                                          # the reference's `root` is still the Int 0 here
                        assert not r["child_visits"][g, 0].any() and r["root_values"][g, 0] == 0
                    else:                                                 # store_search_stats! with the stale root (:374)
                        assert np.array_equal(r["child_visits"][g, i], r["child_visits"][g, i - 1]) and r["root_values"][g, i] == r["root_values"][g, i - 1]
            assert r["sims"] == muzero_plies * cfg.num_iters
    # the opponent "self" is plain self-play
    a = O.arena(cfg, blob, 7, 6, O.OPP_SELF, 1, 1.0, 2); b = O.self_play(cfg, blob, 7, 6, 1.0, 2)
    for k in common.HIST_KEYS:
        assert np.array_equal(a[k], b[k])


def test_expert_beats_random_play_more_often_than_random_does():
    # sanity of the repaired "expert": against the same (random-weights) MuZero it must lose less often than the random opponent
    cfg = O.default_config(); cfg.num_iters = 8
    blob = O.init_weights(cfg, 1)
    lost = {}
    for opp in (O.OPP_RANDOM, O.OPP_EXPERT):
        r = O.arena(cfg, blob, 0, 300, opp, 1, 0.0, 8)
        lost[opp] = int((r["outcome"] == 1).sum())                       # MuZero wins = opponent losses
    assert lost[O.OPP_EXPERT] < lost[O.OPP_RANDOM]


def test_product_scalar_code_matches_oracle():
    hh = common.harness()
    from muzero_jl_b200 import capi
    rng = np.random.default_rng(0)
    for pc, oc in ((capi.default_config(), O.default_config()), (capi.connect_config(), O.connect_config())):
        A = pc.A
        for trial in range(300):
            # a random legal position by random play in the oracle's env
            e = O.Env(); O.lib().mzo_env_reset(C.byref(oc), C.byref(e)); acts = []
            for ply in range(int(rng.integers(0, 3 * A))):
                legal = O.lib().mzo_env_legal_mask(C.byref(oc), C.byref(e))
                if legal == 0 or O.lib().mzo_env_is_terminated(C.byref(oc), C.byref(e)):
                    break
                a = int(rng.choice([i + 1 for i in range(A) if (legal >> i) & 1]))
                O.lib().mzo_env_step(C.byref(oc), C.byref(e), a); acts.append(a)
            if O.lib().mzo_env_is_terminated(C.byref(oc), C.byref(e)):
                for mp in (1, 2):
                    arr = np.asarray(acts, np.int32)
                    assert hh.hh_arena_outcome(C.byref(pc), len(acts), arr.ctypes.data_as(C.POINTER(C.c_int32)), mp) == \
                        O.lib().mzo_arena_outcome(C.byref(oc), len(acts), arr.ctypes.data_as(C.POINTER(C.c_int32)), mp)
                continue
            # the product keeps its own board encoding: replay the actions through its env (same action sequence)
            p1, p2, player = common.product_board(pc, acts)
            for opp in (O.OPP_RANDOM, O.OPP_EXPERT):
                want = O.lib().mzo_opponent_action(C.byref(oc), C.byref(e), opp, C.c_uint64(trial), len(acts) + 1)
                assert hh.hh_opponent_action(C.byref(pc), p1, p2, player, opp, trial, len(acts) + 1) == want, (acts, opp)
