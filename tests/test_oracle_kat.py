"""Pins the CPU oracle against the hand-derived known-answer vectors of SURVEY.md section 8c.
(The reference ships no tests or golden vectors and cannot be executed here: parity is unpinned
beyond these.)  CPU-only."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import oracle as O
import pyref

f32 = np.float32


@pytest.fixture(scope="module")
def cfg():
    return O.default_config()


def play(cfg, actions):
    L = O.lib(); e = O.Env(); L.mzo_env_reset(C.byref(cfg), C.byref(e))
    for a in actions:
        L.mzo_env_step(C.byref(cfg), C.byref(e), a)
    return e


def legal_list(cfg, e):
    m = O.lib().mzo_env_legal_mask(C.byref(cfg), C.byref(e))
    return [a for a in range(1, cfg.A + 1) if m >> (a - 1) & 1]


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert O.philox(0, 0, 0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    seed = 0xffffffffffffffff
    assert O.philox(seed, 0, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    seed = (0x299f31d0 << 32) | 0xa4093822
    assert O.philox(seed, 0, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_julia_dict_order():
    order = (C.c_int32 * 16)()
    O.lib().mzo_julia_dict_order(9, order); assert list(order)[:9] == [7, 4, 9, 2, 3, 5, 8, 6, 1]
    O.lib().mzo_julia_dict_order(7, order); assert list(order)[:7] == [7, 4, 2, 3, 5, 6, 1]
    O.lib().mzo_julia_dict_order(3, order); assert list(order)[:3] == [2, 3, 1]      # Dict(1,2,3) -> 2,3,1
    O.lib().mzo_julia_dict_order(4, order); assert list(order)[:4] == [4, 2, 3, 1]   # Set(1:4) -> 4,2,3,1


def test_kat_env_1(cfg):  # late termination (Q14): P1 completes a line on ply 5, game ends on ply 6
    L = O.lib()
    e = play(cfg, [1, 2, 4, 5, 7])
    assert L.mzo_env_is_terminated(C.byref(cfg), C.byref(e)) == 0
    assert L.mzo_env_reward(C.byref(cfg), C.byref(e), 1) == 0
    assert legal_list(cfg, e) == [3, 6, 8, 9]
    L.mzo_env_step(C.byref(cfg), C.byref(e), 3)
    assert L.mzo_env_is_terminated(C.byref(cfg), C.byref(e)) == 1
    assert L.mzo_env_reward(C.byref(cfg), C.byref(e), 2) == -1
    assert legal_list(cfg, e) == []


def test_kat_env_2(cfg):  # winner mislabel (Q15): P2 actually won, last mover P1 gets +1
    L = O.lib()
    e = play(cfg, [1, 2, 3, 5, 4, 8])
    assert L.mzo_env_is_terminated(C.byref(cfg), C.byref(e)) == 0
    L.mzo_env_step(C.byref(cfg), C.byref(e), 6)
    assert L.mzo_env_is_terminated(C.byref(cfg), C.byref(e)) == 1
    assert L.mzo_env_reward(C.byref(cfg), C.byref(e), 1) == 1


def test_kat_env_3_census(cfg):  # Q16
    out = (C.c_int64 * (32 + 64 * 6))()
    O.lib().mzo_env_census(C.byref(cfg), out)
    assert (out[0], out[1], out[2]) == (6046, 646, 318096)
    assert [out[3 + l] for l in range(10)] == [0, 0, 0, 0, 0, 0, 5760, 15984, 95904, 200448]
    hist = {(l, mv, r): out[32 + l * 6 + (mv - 1) * 3 + (r + 1)] for l in range(10) for mv in (1, 2) for r in (-1, 0, 1)}
    nz = {k: v for k, v in hist.items() if v}
    assert nz == {(6, 2, -1): 5760, (7, 1, 1): 15984, (8, 2, -1): 95904, (9, 1, 0): 127872, (9, 1, 1): 72576}


def test_kat_obs(cfg):  # Q13 stacking layout, raw action index
    L = O.lib(); s = O.sizes(cfg)
    obs = np.zeros((2, s["obs"]), f32); acts = np.array([5, 0], np.int32)
    e = play(cfg, []); L.mzo_env_observation(C.byref(cfg), C.byref(e), O._p(obs[0]))
    e = play(cfg, [5]); L.mzo_env_observation(C.byref(cfg), C.byref(e), O._p(obs[1]))
    st = np.zeros(s["stack"], f32)
    L.mzo_stack_observations(C.byref(cfg), O._p(obs), O._p(acts, C.c_int32), 1, O._p(st))
    assert st.tolist() == [0] * 18 + [1] * 9 + [0] * 36
    L.mzo_stack_observations(C.byref(cfg), O._p(obs), O._p(acts, C.c_int32), 2, O._p(st))
    cur = [0, 0, 0, 0, 1, 0, 0, 0, 0] + [0] * 9 + [1, 1, 1, 1, 0, 1, 1, 1, 1]
    assert st.tolist() == cur + [5.0] * 9 + [0] * 18 + [1] * 9


def test_kat_sizes(cfg):
    assert [O.num_params(cfg, i) for i in range(4)] == [18331, 23242, 33308, 74881]


class NN:
    def __init__(self, cfg, blob):
        self.cfg, self.blob = cfg, blob
    def expf(self, x): return O.lib().mzo_expf(x)
    def representation(self, s): return O.representation(self.cfg, self.blob, s)
    def prediction(self, h): return O.prediction(self.cfg, self.blob, h)
    def dynamics(self, sa): return O.dynamics(self.cfg, self.blob, sa)


def test_kat_dyn_in_and_doubling(cfg):
    """KAT-dyn-in + Q5/Q6: the oracle's trace must equal what the reference does: prediction on the parent
    state scaled 2^(k-1), dynamics on 2^k * h with the action plane float(a/9)."""
    assert f32(3.0 / 9.0).view(np.uint32) == 0x3EAAAAAB
    blob = O.init_weights(cfg)
    c2 = O.default_config(num_iters=6, exploration_eps=0.0, tie_mode=O.TIE_FIRST)
    s = O.sizes(c2); st = np.zeros(s["stack"], f32); st[18:27] = 1
    vc, rv, pri, tr = O.run_mcts(c2, blob, st, 0x1ff, 1, True, 0, 1, trace=True)
    h0 = O.representation(c2, blob, st)
    k = 0
    for depth, action, exp_id, value, reward in tr:
        if exp_id != 0:
            continue
        v, _ = O.prediction(c2, blob, h0 * f32(2.0 ** k))
        sa = np.concatenate([h0 * f32(2.0 ** (k + 1)), np.full(9, f32(action / 9.0), f32)])
        _, r = O.dynamics(c2, blob, sa)
        assert v == value and r == reward
        k += 1
    assert k >= 2


def test_kat_ucb_first_sim_is_tie(cfg):  # Q2: N=0 zeroes every score -> first simulation is a pure tie-break
    blob = O.init_weights(cfg)
    s = O.sizes(cfg); st = np.zeros(s["stack"], f32); st[18:27] = 1
    c1 = O.default_config(num_iters=1, exploration_eps=0.0, tie_mode=O.TIE_FIRST)
    vc, _, _ = O.run_mcts(c1, blob, st, 0x1ff, 1, True, 0, 1)
    assert vc.tolist() == [0, 0, 0, 0, 0, 0, 1, 0, 0]  # first key in Dict order is 7
    vc, _, _ = O.run_mcts(c1, blob, st, 0x1ff & ~(1 << 6), 1, True, 0, 1)
    assert vc.tolist() == [0, 0, 0, 1, 0, 0, 0, 0, 0]  # then 4
    c1 = O.default_config(num_iters=1, exploration_eps=0.0)
    r = O.philox(c1.seed, 1, 3, 2, 1, 1)[0]
    vc, _, _ = O.run_mcts(c1, blob, st, 0x1ff, 1, True, 3, 2)
    assert vc[[7, 4, 9, 2, 3, 5, 8, 6, 1][(r * 9) >> 32] - 1] == 1


@pytest.mark.parametrize("S,legal,to_play,seed", [(10, 0x1ff, 1, 0), (25, 0x1ff, 2, 1), (50, 0b101101011, 1, 2), (50, 0b000011010, 2, 3)])
def test_mcts_matches_independent_python_restatement(cfg, S, legal, to_play, seed):
    """KAT-backup and the whole tree logic: the C oracle vs an independent dict-based Python restatement."""
    c = O.default_config(num_iters=S, exploration_eps=0.0)
    blob = O.init_weights(c, seed=100 + seed)
    rng = np.random.default_rng(seed)
    s = O.sizes(c); st = rng.integers(0, 2, s["stack"]).astype(f32)
    game, move = 11 + seed, 3
    vc, rv, pri = O.run_mcts(c, blob, st, legal, to_play, True, game, move)
    order = list(c.child_order)[:c.A]
    legal_l = [a for a in range(1, c.A + 1) if legal >> (a - 1) & 1]
    pick = lambda sim, depth, n: (O.philox(c.seed, 1, game, move, sim, depth)[0] * n) >> 32
    vc2, rv2, root = pyref.run_mcts(c, NN(c, blob), st, legal_l, to_play, order, pick)
    assert vc.tolist() == vc2.tolist()
    assert rv == rv2
    assert vc.sum() == S
    for a in legal_l:
        assert pri[a - 1] == root.children[a].prior


def test_kat_backup_hand_derived(cfg):
    """KAT-backup: S=2 with forced ties -> path root->c->leaf; check value sums against the closed form
    (two players, Q8): leaf +v2; c: +v1 then +r_leaf; root: -v1... derived from SelfPlay.jl:199-212."""
    c = O.default_config(num_iters=2, exploration_eps=0.0, tie_mode=O.TIE_FIRST)
    blob = O.init_weights(c)
    s = O.sizes(c); st = np.zeros(s["stack"], f32); st[18:27] = 1
    order = list(c.child_order)[:9]
    vc2, rv2, root = pyref.run_mcts(c, NN(c, blob), st, list(range(1, 10)), 1, order, lambda *a: 0)
    vc, rv, _ = O.run_mcts(c, blob, st, 0x1ff, 1, True, 0, 1)
    assert vc.tolist() == vc2.tolist() and rv == rv2
    # sim 1: leaf = child 7 (tie, first).  v1 = prediction(h0).value, r1 = dynamics reward.
    h0 = O.representation(c, blob, st)
    v1, _ = O.prediction(c, blob, h0)
    _, r1 = O.dynamics(c, blob, np.concatenate([h0 * f32(2), np.full(9, f32(7 / 9.0), f32)]))
    g = f32(c.discount)
    # after sim 1: child7.value_sum = +v1 (to_play 2 == vtp 2); then value = -r1; root.to_play(1) != 2:
    # root.value_sum = -(-r1) = r1
    root_vs = f32(f32(0) - f32(-r1))
    if vc[6] == 1:  # sim 2 went elsewhere at depth 1 (n=1 child has value term): closed form for a depth-1 leaf
        a2 = int(np.argmax(vc * (np.arange(9) != 6))) + 1
        v2, _ = O.prediction(c, blob, h0 * f32(2))
        _, r2 = O.dynamics(c, blob, np.concatenate([h0 * f32(4), np.full(9, f32(a2 / 9.0), f32)]))
        root_vs = f32(root_vs - f32(-r2))
        assert root.children[a2].value_sum == v2
    assert root.value_sum == root_vs
    assert rv == f32(root_vs / f32(2))


def test_kat_target(cfg):  # Q17
    T = 9
    rng = np.random.default_rng(0)
    rew = rng.uniform(-1, 1, 10).astype(f32); rv = rng.uniform(-1, 1, 10).astype(f32)
    tp = np.array([1, 2, 1, 2, 1, 2, 1, 2, 1, 0], np.int32)
    L = O.lib()
    tv = lambda idx: L.mzo_compute_target_value(C.byref(cfg), T, O._p(rew), O._p(tp, C.c_int32), O._p(rv), idx)
    for idx in range(4, 10):
        assert tv(idx) == 0.0
    g = f32(0.997)
    gp = [f32(1), g, f32(g * g), f32(f32(g * g) * g)] + [f32(np.float32(math.pow(float(g), i))) for i in range(4, 8)]
    for idx in (1, 2, 3):
        b = idx + 5
        last = rv[b - 1] if tp[b - 1] == tp[idx - 1] else f32(-rv[b - 1])
        v = f32(last * f32(np.float32(0.997) ** 5 if False else gp[5]))
        for i in range(1, 7):
            r = rew[idx + i - 2]
            v = f32(v + f32((r if tp[idx - 1] == tp[idx + i - 1] else f32(-r)) * gp[i]))
        assert abs(tv(idx) - v) <= 1e-7


def test_kat_loss_uniform(cfg):  # Q21: uniform P and uniform target -> s_j = 6*log(9); loss = mean_j(s_j)*mean_i(1/g_i)
    B, K1, A = 4, 6, 9
    c = O.default_config(batch_size=B)
    s = O.sizes(c)
    blob = np.zeros(O.num_params(c), f32)  # all-zero nets: logits 0 -> P uniform, value tanh(0) = 0
    g = np.array([5, 4, 2, 1], f32)
    batch = dict(obs=np.zeros((B, s["stack"]), f32), actions=np.ones((B, K1), f32), values=np.zeros((B, K1), f32),
                 rewards=np.zeros((B, K1), f32), policies=np.full((B, K1, A), f32(1) / f32(9), f32), gscale=g)
    pv, pr, pp, losses = O.learn_forward(c, blob, batch)
    assert np.all(pp == f32(1) / f32(9)) and np.all(pv == 0) and np.all(pr == 0)
    expect = 6 * math.log(9) * float(np.mean(1.0 / g))
    assert losses[0] == losses[1] == losses[2]
    assert abs(float(losses[0]) - expect) < 1e-5 * expect
    # value term: v = 0 vs target t: mean_b(sum_k t^2 / g_b)
    batch["values"][:] = 0.5
    _, _, _, l2 = O.learn_forward(c, blob, batch)
    assert abs(float(l2[0]) - (expect + float(np.mean(6 * 0.25 / g)))) < 1e-5 * expect


def test_kat_grad_reference_l2(cfg):  # Q20 + Q22
    blob = O.init_weights(cfg); theta0 = blob.copy()
    m = np.zeros_like(blob); v = np.zeros_like(blob)
    hist = O.self_play(O.default_config(exploration_eps=0.0), blob, 0, 8, 1.0, 1)
    batch = O.get_batch(cfg, hist, step=1)
    losses = O.learn_step(cfg, blob, m, v, 1, batch)
    assert O.lib().mzo_cos_schedule(1) == pytest.approx(0.1, abs=1e-15)
    assert O.lib().mzo_cos_schedule(6) == pytest.approx(1e-4, abs=1e-12)
    nz = theta0 != 0
    step = theta0 - blob
    # first Adam step: eta * g/(|g| + eps') ~= eta*sign(theta) for g = 2*theta
    expect = 0.1 * theta0[nz] / (np.abs(theta0[nz]) + 5e-9)
    assert np.allclose(step[nz], expect, rtol=1e-4, atol=1e-7)
    assert np.all(blob[~nz] == 0)  # zero-initialised biases never move
    assert np.all(np.isfinite(losses)) and losses[0] != losses[1]


def test_self_play_shapes_and_invariants(cfg):
    c = O.default_config(exploration_eps=0.0)
    blob = O.init_weights(c)
    h = O.self_play(c, blob, 0, 32, 1.0, 2)
    h1 = O.self_play(c, blob, 0, 32, 1.0, 1)
    for k in ("T", "actions", "rewards", "child_visits", "root_values", "obs"):
        assert np.array_equal(h[k], h1[k])  # threading does not change results
    assert h["sims"] == int(h["T"].sum()) * c.num_iters
    assert h["T"].min() >= 6 and h["T"].max() <= 9  # Q14: never 5 plies
    for g in range(32):
        T = h["T"][g]
        assert np.allclose(h["child_visits"][g, :T].sum(1), 1.0, atol=1e-6)
        assert np.all(h["rewards"][g, :T - 1] == 0)
        assert list(h["to_play"][g, :T]) == [1 + (i % 2) for i in range(T)]


# ---- grad_mode = BPTT: the Float64 backward of the oracle is pinned by finite differences of its own Float64 loss ----
def _bptt_case(intermediate_rewards, B=4, seed=3):
    c = O.default_config(batch_size=B, intermediate_rewards=int(intermediate_rewards), exploration_eps=0.0)
    blob = O.init_weights(c, seed)
    rng = np.random.default_rng(seed)
    blob = blob + (rng.standard_normal(blob.shape[0]) * 0.02).astype(f32)      # non-zero biases
    hist = O.self_play(c, blob, 0, 8, 1.0, 1)
    batch = O.get_batch(c, hist, step=seed)
    batch["rewards"] = batch["rewards"] + (rng.standard_normal(batch["rewards"].shape) * 0.3).astype(f32)   # exercise the reward head
    return c, blob, batch


@pytest.mark.parametrize("ir", [0, 1])
def test_bptt_gradient_matches_finite_differences(ir):
    c, blob, batch = _bptt_case(ir)
    loss, grad = O.learn_gradients(c, blob, batch, fwd64=True)
    assert np.isfinite(loss) and np.all(np.isfinite(grad))
    rng = np.random.default_rng(11)
    n = blob.shape[0]
    nr, npred = O.num_params(c, 0), O.num_params(c, 1)
    picks = list(rng.integers(0, nr, 12)) + list(nr + rng.integers(0, npred, 12)) + list(nr + npred + rng.integers(0, n - nr - npred, 16))
    h = 1e-6
    worst = 0.0
    for i in picks:
        lp, _ = O.learn_gradients(c, blob, batch, fwd64=True, perturb=(int(i), +h), want_grad=False)
        lm, _ = O.learn_gradients(c, blob, batch, fwd64=True, perturb=(int(i), -h), want_grad=False)
        fd = (lp - lm) / (2 * h) + 2.0 * float(blob[i])          # the returned loss is the data loss; grad adds 2*theta
        worst = max(worst, abs(fd - grad[i]) / (abs(grad[i]) + 1e-6))
    assert worst < 2e-4, worst
    if not ir:   # without intermediate rewards the reward head only sees the L2 term (Learning.jl:277-279)
        net = O.lib().mzo_num_params
        nd = O.num_params(c, 2)
        # reward head = last (depth_reward + 1) layers of the dynamics net: 64*64+64 + 64*1+1 parameters
        tail = 64 * 64 + 64 + 64 + 1
        assert np.allclose(grad[n - tail:], 2.0 * blob[n - tail:].astype(np.float64), rtol=0, atol=1e-12)


def test_bptt_gradient_with_batchnorm_matches_finite_differences():
    """FeedForwardHP.use_batch_norm (Learning.jl:70-79), test mode: the Float64 backward through gamma * (W x + b - mu) / sqrt(sigma2 + 1f-5) + beta
    against finite differences of the oracle's own Float64 loss, for W, b, beta and gamma entries; mu and sigma2 are not Flux parameters: 0."""
    from test_host_harness import _randomise_batchnorm
    c = O.default_config(batch_size=4, intermediate_rewards=1, exploration_eps=0.0, use_batch_norm=1)
    blob = O.init_weights(c, 3)
    rng = np.random.default_rng(3)
    blob = _randomise_batchnorm(c, blob + (rng.standard_normal(blob.shape[0]) * 0.02).astype(f32), 5)
    hist = O.self_play(c, blob, 0, 8, 1.0, 1)
    batch = O.get_batch(c, hist, step=3)
    batch["rewards"] = batch["rewards"] + (rng.standard_normal(batch["rewards"].shape) * 0.3).astype(f32)
    loss, grad = O.learn_gradients(c, blob, batch, fwd64=True)
    assert np.isfinite(loss) and np.all(np.isfinite(grad))
    mask = O.trainable_mask(c)
    assert np.all(grad[mask == 0] == 0.0)
    stats = np.flatnonzero(mask == 0)
    runs = np.split(stats, np.flatnonzero(np.diff(stats) > 1) + 1)
    picks = []
    for r in runs:                                          # per BatchNorm: two beta, two gamma, two bias and two weight entries in front of it
        n = len(r) // 2
        picks += [r[0] - 2 * n + 1, r[0] - n - 1, r[0] - n, r[0] - 1, r[0] - 3 * n, r[0] - 2 * n - 1, r[0] - 3 * n - 1, r[0] - 3 * n - 7]
    picks += list(rng.integers(0, blob.shape[0], 20))
    h = 1e-6
    worst = 0.0
    for i in picks:
        if not mask[i]:
            continue
        lp, _ = O.learn_gradients(c, blob, batch, fwd64=True, perturb=(int(i), +h), want_grad=False)
        lm, _ = O.learn_gradients(c, blob, batch, fwd64=True, perturb=(int(i), -h), want_grad=False)
        fd = (lp - lm) / (2 * h) + 2.0 * float(blob[i])
        worst = max(worst, abs(fd - grad[i]) / (abs(grad[i]) + 1e-6))
    assert worst < 2e-4, worst
    # the Float32-forward variant (what the CUDA path is compared with) linearises around almost the same point
    _, g32 = O.learn_gradients(c, blob, batch, fwd64=False)
    assert np.max(np.abs(g32 - grad)) <= 1e-3 * np.max(np.abs(grad))


def test_bptt_float32_forward_is_close_to_float64_forward():
    c, blob, batch = _bptt_case(1)
    l64, g64 = O.learn_gradients(c, blob, batch, fwd64=True)
    l32, g32 = O.learn_gradients(c, blob, batch, fwd64=False)
    _, _, _, losses = O.learn_forward(c, blob, batch)
    assert abs(l64 - l32) < 1e-5 * abs(l64)
    assert np.max(np.abs(g64 - g32)) < 1e-4 * np.max(np.abs(g64))
    # the Float32 loss scalar of learn_forward = data loss + that net's own sum(theta^2)
    nr = O.num_params(c, 0)
    assert abs(float(losses[0]) - (l32 + float(np.sum(blob[:nr].astype(np.float64) ** 2)))) < 1e-4 * float(losses[0])


def test_bptt_learn_step_uses_the_gradient():
    c, blob, batch = _bptt_case(0)
    _, g = O.learn_gradients(c, blob, batch, fwd64=False)
    b1 = blob.copy(); m1 = np.zeros_like(blob); v1 = np.zeros_like(blob)
    O.learn_step(c, b1, m1, v1, 1, batch, grad_mode=O.GRAD_BPTT)
    b2 = blob.copy(); m2 = np.zeros_like(blob); v2 = np.zeros_like(blob)
    O.adam_apply(b2, m2, v2, g.astype(f32), 1)
    assert np.array_equal(b1, b2) and np.array_equal(m1, m2) and np.array_equal(v1, v2)


def test_golden_bptt_per_resnet_fixture_is_reproduced():
    """tests/golden/bptt_per_resnet.npz (generator committed) freezes the oracle's BPTT gradient, prioritised batch, priority update
    and ResNet outputs: any drift of the oracle shows up here, on the CPU."""
    import os
    import common
    g = np.load(os.path.join(common.ROOT, "tests", "golden", "bptt_per_resnet.npz"))
    cfg = O.default_config(batch_size=24, per=1, intermediate_rewards=1)
    blob = O.init_weights(cfg, 77)
    hist = {k: g["hist_" + k] for k in common.HIST_KEYS}
    q_pos, q_game = O.per_priorities(cfg, hist)
    assert np.array_equal(q_pos, g["q_pos"]) and np.array_equal(q_game, g["q_game"])
    pb = O.get_batch_per(cfg, hist, q_pos, q_game, 3)
    for k in common.BATCH_KEYS + ("weights",):
        assert np.array_equal(pb[k], g["pb_" + k]), k
    _, grad = O.learn_gradients_w(cfg, blob, pb, fwd64=False)
    assert np.array_equal(grad.astype(np.float32), g["grad"])
    rcfg = O.resnet_config(num_iters=20, exploration_eps=0.0)
    rblob = O.init_weights(rcfg, 5)
    assert np.array_equal(np.stack([O.representation(rcfg, rblob, x) for x in g["rn_stacked"]]), g["rn_hidden_f32"])
    O.set_bf16(True)
    try:
        assert np.array_equal(np.stack([O.representation(rcfg, rblob, x) for x in g["rn_stacked"][:4]]), g["rn_hidden_bf16"][:4])
    finally:
        O.set_bf16(False)


def test_temperature_threshold_switches_to_greedy_play():
    """conf.temperature_threshold (SelfPlay.jl:344-346): once length(action_history) >= threshold the rest of the game is played at
    temperature 0.  Moves are independent given the position (every draw is keyed by (game, move)), so: threshold 0 == a game at
    temperature 0; threshold 3 == the temperature-1 game for its first 3 plies, and from then on the argmax of the visit counts."""
    cfg = O.default_config(num_iters=12, exploration_eps=0.25)
    blob = O.init_weights(cfg, 21)
    n = 40
    warm = O.self_play(cfg, blob, 500, n, 1.0, 2)
    greedy = O.self_play(cfg, blob, 500, n, 0.0, 2)
    cfg.temperature_threshold = 0
    t0 = O.self_play(cfg, blob, 500, n, 1.0, 2)
    for k in ("T", "actions", "child_visits", "root_values"):
        assert np.array_equal(t0[k], greedy[k]), k
    cfg.temperature_threshold = 3
    t3 = O.self_play(cfg, blob, 500, n, 1.0, 2)
    assert np.array_equal(t3["actions"][:, :3], warm["actions"][:, :3])
    assert not np.array_equal(t3["actions"], warm["actions"])            # some later ply differs from the sampled game
    order = cfg.child_order[:9]
    for g in range(n):
        for t in range(3, int(t3["T"][g])):
            cv = t3["child_visits"][g, t]
            best = max(order, key=lambda a: (cv[a - 1], -order.index(a)))   # argmax over the children in Dict order, first maximum wins
            assert int(t3["actions"][g, t]) == best, (g, t)
    cfg.temperature_threshold = 64                                        # never reached: identical to nothing
    assert np.array_equal(O.self_play(cfg, blob, 500, n, 1.0, 2)["actions"], warm["actions"])
