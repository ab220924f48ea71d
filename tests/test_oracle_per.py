"""Oracle tests of prioritised replay (conf.PER = true) in its repaired form (the reference's update_priorities! indexes out
of bounds, ReplayBuffer.jl:176-178): known answers for the priorities, the sampling law, the importance weights, the weighted
loss and its gradient (finite differences), and the priority update."""
import numpy as np
import pytest

from oracle import oracle as O

f32 = np.float32


def _setup(B=16, alpha=1, games=12, seed=2):
    c = O.default_config(batch_size=B, per=1, per_alpha=alpha, exploration_eps=0.0)
    blob = O.init_weights(c, seed)
    hist = O.self_play(c, blob, 0, games, 1.0, 1)
    return c, blob, hist


@pytest.mark.parametrize("alpha", [1, 2])
def test_per_initial_priorities(alpha):
    c, blob, hist = _setup(alpha=alpha)
    q_pos, q_game = O.per_priorities(c, hist)
    for g in range(len(hist["T"])):
        T = hist["T"][g]
        for i in range(1, T + 1):
            tv = O.lib().mzo_compute_target_value(O.C.byref(c), int(T), O._p(np.ascontiguousarray(hist["rewards"][g])),
                                                  O._p(np.ascontiguousarray(hist["to_play"][g], np.int32), O.C.c_int32),
                                                  O._p(np.ascontiguousarray(hist["root_values"][g])), i)
            p = f32(abs(f32(hist["root_values"][g, i - 1]) - f32(tv)))
            p = p if alpha == 1 else f32(p * p)
            assert q_pos[g, i - 1] == max(1, int(np.rint(np.float64(f32(p * f32(65536.0))))))
        assert q_game[g] == q_pos[g, :T].max() and not q_pos[g, T:].any()
    assert O.lib().mzo_per_quantise(0.0) == 1 and O.lib().mzo_per_quantise(0.5) == 32768


def test_per_sampling_follows_the_priorities():
    c, blob, hist = _setup(B=64)
    n = len(hist["T"])
    q_pos = np.ones((n, c.max_moves + 1), np.uint32); q_game = np.ones(n, np.uint32)
    q_game[3] = 1000; q_pos[3, 2] = 500                       # game key 4 dominates; inside it position 3
    counts = np.zeros(n); pos3 = 0; tot3 = 0
    for step in range(1, 41):
        b = O.get_batch_per(c, hist, q_pos, q_game, step)
        for key, pos in b["index"]:
            counts[key - 1] += 1
            if key == 4:
                tot3 += 1; pos3 += int(pos == 3)
        assert b["weights"].max() == 1.0 and b["weights"].min() > 0
    frac = counts[3] / counts.sum()
    assert abs(frac - 1000 / (1000 + n - 1)) < 0.02
    T3 = hist["T"][3]
    assert abs(pos3 / tot3 - 500 / (500 + T3 - 1)) < 0.03
    # weights: 1 / (total_samples * game_prob * pos_prob) normalised by the maximum
    b = O.get_batch_per(c, hist, q_pos, q_game, 7)
    Q = int(q_game.sum()); total = int(hist["T"].sum())
    w = []
    for key, pos in b["index"]:
        g = key - 1; Qg = int(q_pos[g, :hist["T"][g]].sum())
        gp = f32(q_game[g] / Q); pp = f32(q_pos[g, pos - 1] / Qg)
        w.append(f32(1.0) / f32(f32(f32(total) * gp) * pp))
    w = np.array(w, f32)
    assert np.array_equal(b["weights"], w / w.max())
    # uniform priorities reproduce uniform sampling statistics
    q_pos[:] = 1; q_game[:] = 1
    b = O.get_batch_per(c, hist, q_pos, q_game, 3)
    assert len(set(b["index"][:, 0])) > 5


def test_per_weighted_loss_and_gradient():
    c, blob, hist = _setup(B=6)
    rng = np.random.default_rng(1)
    blob = blob + (rng.standard_normal(blob.shape[0]) * 0.02).astype(f32)
    q_pos, q_game = O.per_priorities(c, hist)
    batch = O.get_batch_per(c, hist, q_pos, q_game, 4)
    _, _, _, l_w = O.learn_forward_w(c, blob, batch)
    ones = dict(batch); ones["weights"] = np.ones_like(batch["weights"])
    _, _, _, l_1 = O.learn_forward_w(c, blob, ones)
    _, _, _, l_0 = O.learn_forward(c, blob, batch)
    assert np.array_equal(l_1, l_0) and not np.array_equal(l_w, l_0)
    loss, grad = O.learn_gradients_w(c, blob, batch, fwd64=True)
    worst = 0.0
    for i in rng.integers(0, blob.shape[0], 30):
        lp, _ = O.learn_gradients_w(c, blob, batch, fwd64=True, perturb=(int(i), 1e-6), want_grad=False)
        lm, _ = O.learn_gradients_w(c, blob, batch, fwd64=True, perturb=(int(i), -1e-6), want_grad=False)
        fd = (lp - lm) / 2e-6 + 2.0 * float(blob[i])
        worst = max(worst, abs(fd - grad[i]) / (abs(grad[i]) + 1e-6))
    assert worst < 2e-4, worst


def test_per_update_priorities_repaired_bounds_and_order():
    c, blob, hist = _setup(B=4)
    n = len(hist["T"]); K1 = c.num_unroll_steps + 1
    q_pos, q_game = O.per_priorities(c, hist)
    T0 = int(hist["T"][0])
    index = np.array([[1, T0 - 1], [1, 1], [2, 2], [1, 1]], np.int32)    # game 1 three times: the LAST batch element wins at overlapping positions
    pv = np.zeros((4, K1), f32); tv = np.zeros((4, K1), f32)
    pv[0] = 0.25; pv[1] = 0.5; pv[2] = 1.0; pv[3] = 0.125
    before = q_pos.copy()
    O.per_update(c, hist, q_pos, q_game, index, pv, tv)
    q = lambda x: O.lib().mzo_per_quantise(x)
    nk = min(K1, T0)                                                      # rows k with 1 + k <= T
    assert np.all(q_pos[0, :nk] == q(0.125))                              # element 3 overwrote element 1
    if T0 - 1 > nk:
        assert q_pos[0, T0 - 2] == q(0.25)
    assert q_pos[0, T0 - 1] == (q(0.25) if T0 > nk else q(0.125))         # element 0 wrote positions T-1, T only (clipped at T)
    assert not q_pos[0, T0:].any() and q_game[0] == q_pos[0, :T0].max()
    T1 = int(hist["T"][1])
    assert np.all(q_pos[1, 1:min(1 + K1, T1)] == q(1.0)) and q_pos[1, 0] == before[1, 0]
    assert np.array_equal(q_pos[2:], before[2:])
