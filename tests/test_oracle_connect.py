"""Oracle tests of the synthetic Connect game (BASELINE.json configs[3]: 6x7 board, 7 actions, column drop, four in a row,
clean termination, reward to the last mover) -- the reference ships no such game, so the rules are pinned by known answers and
by an independent numpy checker over random playouts."""
import ctypes as C

import numpy as np

from oracle import oracle as O

L = O.lib()


def play(cfg, actions):
    e = O.Env(); L.mzo_env_reset(C.byref(cfg), C.byref(e))
    for a in actions:
        assert L.mzo_env_legal_mask(C.byref(cfg), C.byref(e)) >> (a - 1) & 1
        L.mzo_env_step(C.byref(cfg), C.byref(e), a)
    return e


def status(cfg, e):
    return (L.mzo_env_is_terminated(C.byref(cfg), C.byref(e)), L.mzo_env_reward(C.byref(cfg), C.byref(e), 1),
            L.mzo_env_reward(C.byref(cfg), C.byref(e), 2), L.mzo_env_legal_mask(C.byref(cfg), C.byref(e)))


def test_connect_known_answers():
    cfg = O.connect_config()
    assert list(cfg.child_order[:7]) == sorted(cfg.child_order[:7], key=list(cfg.child_order[:7]).index) and sorted(cfg.child_order[:7]) == list(range(1, 8))
    # vertical: P1 plays column 1 four times, P2 column 2
    e = play(cfg, [1, 2, 1, 2, 1, 2, 1])
    assert status(cfg, e) == (1, 1, -1, 0)                      # terminal, +1 for the last mover (P1), -1 for P2, no legal moves
    e = play(cfg, [1, 2, 1, 2, 1, 2])
    assert status(cfg, e) == (0, 0, 0, 0b1111111)
    # horizontal: P2 wins on the bottom row (columns 4..7) while P1 stacks column 1
    e = play(cfg, [1, 4, 1, 5, 1, 6, 2, 7])
    assert status(cfg, e) == (1, -1, 1, 0)
    # diagonal /: P1 at (0,c1) (1,c2) (2,c3) (3,c4)
    e = play(cfg, [1, 2, 2, 3, 3, 4, 3, 4, 4, 6, 4])
    assert status(cfg, e)[0] == 1 and status(cfg, e)[1] == 1
    # a full column is illegal
    e = play(cfg, [3, 3, 3, 3, 3, 3])
    assert status(cfg, e)[3] == 0b1111011 and status(cfg, e)[0] == 0
    obs = np.zeros(6 * 7 * 3, np.float32); L.mzo_env_observation(C.byref(cfg), C.byref(e), O._p(obs))
    o = obs.reshape(3, 7, 6)                                    # Julia (W,H,C) = (rows, columns, planes)
    assert o[0, 2].tolist() == [1, 0, 1, 0, 1, 0] and o[1, 2].tolist() == [0, 1, 0, 1, 0, 1] and o[2].sum() == 36


def _has4(board, who):
    b = board == who
    R, Cn = b.shape
    for r in range(R):
        for c in range(Cn):
            for dr, dc in ((1, 0), (0, 1), (1, 1), (1, -1)):
                if all(0 <= r + k * dr < R and 0 <= c + k * dc < Cn and b[r + k * dr, c + k * dc] for k in range(4)):
                    return True
    return False


def test_connect_random_playouts_match_numpy_rules():
    cfg = O.connect_config()
    rng = np.random.default_rng(5)
    ends = {"win": 0, "draw": 0}
    for g in range(300):
        e = O.Env(); L.mzo_env_reset(C.byref(cfg), C.byref(e))
        board = np.zeros((6, 7), np.int32); player = 1
        for ply in range(43):
            m = L.mzo_env_legal_mask(C.byref(cfg), C.byref(e))
            won = ply > 0 and _has4(board, 3 - player)
            full = bool((board != 0).all())
            assert bool(L.mzo_env_is_terminated(C.byref(cfg), C.byref(e))) == (won or full)
            expect = 0 if won else sum(1 << c for c in range(7) if board[5, c] == 0)
            assert m == expect
            if won or full:
                last = 3 - player
                assert L.mzo_env_reward(C.byref(cfg), C.byref(e), last) == (1 if won else 0)
                assert L.mzo_env_reward(C.byref(cfg), C.byref(e), player) == (-1 if won else 0)
                ends["win" if won else "draw"] += 1
                break
            a = int(rng.choice([c for c in range(7) if m >> c & 1])) + 1
            r = int((board[:, a - 1] != 0).sum()); board[r, a - 1] = player
            L.mzo_env_step(C.byref(cfg), C.byref(e), a); player = 3 - player
            assert e.player == player
    assert ends["win"] > 200


def test_connect_resnet_self_play_runs():
    cfg = O.connect_config(num_iters=8, rn_num_filters=16, rn_num_blocks=1, exploration_eps=0.0)
    blob = O.init_weights(cfg, 3)
    h = O.self_play(cfg, blob, 0, 3, 1.0, 1)
    assert h["sims"] == int(h["T"].sum()) * 8 and h["T"].min() >= 7 and h["T"].max() <= 42
    for g in range(3):
        T = h["T"][g]
        assert np.allclose(h["child_visits"][g, :T].sum(1), 1.0, atol=1e-6)
        assert set(np.unique(h["rewards"][g, :T - 1])) <= {0.0} and h["rewards"][g, T - 1] in (0.0, 1.0)
