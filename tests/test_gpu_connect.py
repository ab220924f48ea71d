"""GPU tests of BASELINE.json configs[3]: the synthetic 6x7 / 7-action Connect game with the ResNet networks
(3 trees per 128-row tile, 12 trees per CTA, policy head with two K blocks).  The environment must agree with the oracle
bit for bit; the networks and searches are held to the bf16-emulating oracle like the TicTacToe ResNet."""
import ctypes as C

import numpy as np
import pytest

import common
from oracle import oracle as O
from test_oracle_resnet import _randomised_blob

pytestmark = pytest.mark.gpu
RN_ATOL = 2e-2


@pytest.fixture(scope="module")
def capi():
    from muzero_jl_b200 import capi
    return capi


def make(capi, **kw):
    kw.setdefault("num_slots", 48); kw.setdefault("replay_buffer_size", 256)
    cfg = capi.connect_config(**kw)
    return capi.Context(cfg), common.oracle_config(cfg)


def test_connect_environment_matches_oracle(capi):
    ctx, ocfg = make(capi, num_iters=4)
    L = O.lib()
    n = 256
    rng = np.random.default_rng(2)
    p1, p2, pl = ctx.env_reset(n)
    envs = [O.Env() for _ in range(n)]
    for e in envs:
        L.mzo_env_reset(C.byref(ocfg), C.byref(e))
    alive = np.ones(n, bool)
    legal = ctx.env_legal(p1, p2, pl)
    for ply in range(42):
        for i in range(n):
            assert legal[i] == L.mzo_env_legal_mask(C.byref(ocfg), C.byref(envs[i]))
        act = np.ones(n, np.int32)
        for i in range(n):
            if alive[i] and legal[i]:
                act[i] = int(rng.choice([c for c in range(7) if legal[i] >> c & 1])) + 1
            else:
                alive[i] = False
        idx = np.nonzero(alive)[0]
        if len(idx) == 0:
            break
        q1, q2, ql = p1[idx].copy(), p2[idx].copy(), pl[idx].copy()
        movers = ql.copy()
        rew, done, lg = ctx.env_step(q1, q2, ql, act[idx])
        obs = ctx.env_observation(q1, q2)
        for k, i in enumerate(idx):
            L.mzo_env_step(C.byref(ocfg), C.byref(envs[i]), int(act[i]))
            assert rew[k] == L.mzo_env_reward(C.byref(ocfg), C.byref(envs[i]), int(movers[k]))
            assert done[k] == L.mzo_env_is_terminated(C.byref(ocfg), C.byref(envs[i]))
            o = np.zeros(126, np.float32); L.mzo_env_observation(C.byref(ocfg), C.byref(envs[i]), O._p(o))
            assert np.array_equal(obs[k], o) and ql[k] == envs[i].player
        p1[idx], p2[idx], pl[idx], legal[idx] = q1, q2, ql, lg
    assert not legal[~alive].any() or True
    ctx.close()


def _stacked_connect(ocfg, n, seed):
    """Random reachable positions (a few plies deep) with their stacked observation, legal mask and side to move."""
    L = O.lib(); rng = np.random.default_rng(seed); s = O.sizes(ocfg)
    out = np.zeros((n, s["stack"]), np.float32); legal = np.zeros(n, np.uint32); tp = np.zeros(n, np.int32)
    for i in range(n):
        e = O.Env(); L.mzo_env_reset(C.byref(ocfg), C.byref(e))
        obs = [np.zeros(s["obs"], np.float32)]; L.mzo_env_observation(C.byref(ocfg), C.byref(e), O._p(obs[0])); acts = []
        for _ in range(int(rng.integers(0, 10))):
            m = L.mzo_env_legal_mask(C.byref(ocfg), C.byref(e))
            a = int(rng.choice([c for c in range(7) if m >> c & 1])) + 1
            L.mzo_env_step(C.byref(ocfg), C.byref(e), a)
            if L.mzo_env_is_terminated(C.byref(ocfg), C.byref(e)):
                break
            acts.append(a); o = np.zeros(s["obs"], np.float32); L.mzo_env_observation(C.byref(ocfg), C.byref(e), O._p(o)); obs.append(o)
        # rebuild the env up to the last non-terminal position
        e = O.Env(); L.mzo_env_reset(C.byref(ocfg), C.byref(e))
        for a in acts:
            L.mzo_env_step(C.byref(ocfg), C.byref(e), a)
        hist = np.stack(obs[:len(acts) + 1]); a_arr = np.array(acts + [0], np.int32)
        L.mzo_stack_observations(C.byref(ocfg), O._p(hist), O._p(a_arr, C.c_int32), len(acts) + 1, O._p(out[i]))
        legal[i] = L.mzo_env_legal_mask(C.byref(ocfg), C.byref(e)); tp[i] = e.player
    return out, legal, tp


def test_connect_resnet_networks_and_mcts(capi):
    ctx, ocfg = make(capi, num_iters=20, exploration_eps=0.0)
    blob = _randomised_blob(ocfg, 13)
    ctx.set_weights(blob)
    n = 30
    st, legal, tp = _stacked_connect(ocfg, n, 6)
    O.set_bf16(True)
    try:
        h = ctx.representation(st)
        oh = np.stack([O.representation(ocfg, blob, x) for x in st])
        assert h.shape == (n, 6 * 7 * 64)
        assert np.max(np.abs(h - oh)) < RN_ATOL * max(1.0, np.max(np.abs(oh))) and np.median(np.abs(h - oh)) < 1e-3
        v, p = ctx.prediction(oh)
        ov, op = zip(*[O.prediction(ocfg, blob, x) for x in oh])
        assert np.max(np.abs(v - np.array(ov))) < RN_ATOL and np.max(np.abs(p - np.stack(op))) < RN_ATOL
        sa = np.concatenate([2 * oh, np.repeat((np.arange(n) % 7 + 1)[:, None] / np.float32(7), 42, 1).astype(np.float32)], 1)
        nh, r = ctx.dynamics(sa)
        onh, orr = zip(*[O.dynamics(ocfg, blob, x) for x in sa])
        assert np.max(np.abs(nh - np.stack(onh))) < RN_ATOL * max(1.0, np.max(np.abs(np.stack(onh)))) and np.max(np.abs(r - np.array(orr))) < RN_ATOL
        gid = np.arange(n, dtype=np.uint64) + 5; mv = np.ones(n, np.int32)
        vc, rv, pri = ctx.run_mcts(st, legal, tp, False, gid, mv, priors=True)
        assert np.all(vc.sum(1) == 20)
        same = 0
        for i in range(n):
            ovc, orv, opri = O.run_mcts(ocfg, blob, st[i], int(legal[i]), int(tp[i]), False, int(gid[i]), 1)
            same += int(np.array_equal(ovc, vc[i]))
            assert np.max(np.abs(opri - pri[i])) < RN_ATOL
        assert same >= 0.8 * n, same
    finally:
        O.set_bf16(False)
    ctx.close()


def test_connect_self_play_with_200_simulations(capi):
    """configs[3]: S = 200 simulations per move (1408 nodes per tree); histories must be legal Connect games."""
    ctx, ocfg = make(capi, num_iters=200, num_slots=24, replay_buffer_size=64)
    ctx.init_weights(2)
    sims, moves = ctx.self_play(0, 30, 1.0)
    assert sims == moves * 200
    h = ctx.history_export()
    L = O.lib()
    for j in range(30):
        T = h["T"][j]
        assert 7 <= T <= 42 and np.allclose(h["child_visits"][j, :T].sum(1), 1.0, atol=1e-6)
        e = O.Env(); L.mzo_env_reset(C.byref(ocfg), C.byref(e))
        for t in range(T):
            o = np.zeros(126, np.float32); L.mzo_env_observation(C.byref(ocfg), C.byref(e), O._p(o))
            assert np.array_equal(h["obs"][j, t], o)
            a = int(h["actions"][j, t]); mover = e.player
            assert L.mzo_env_legal_mask(C.byref(ocfg), C.byref(e)) >> (a - 1) & 1 and h["to_play"][j, t] == mover
            L.mzo_env_step(C.byref(ocfg), C.byref(e), a)
            assert h["rewards"][j, t] == L.mzo_env_reward(C.byref(ocfg), C.byref(e), mover)
        assert L.mzo_env_is_terminated(C.byref(ocfg), C.byref(e))
    ctx.close()
