"""CPU-only: the product's scalar device code (csrc/mz_common.h + mz_host.h), compiled by g++ and replayed one
kernel-thread at a time by tests/host_harness.cpp, must agree bit-for-bit with the oracle.  This validates the
tree logic, weight packing, UCB tables, RNG keying and target construction before any GPU time is spent."""
import ctypes as C

import numpy as np
import pytest

import common
from oracle import oracle as O

f32p = C.POINTER(C.c_float)


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


@pytest.fixture(scope="module")
def hh():
    return common.harness()


def test_config_layouts_agree(hh):
    from muzero_jl_b200 import capi
    cfg = common.product_config()
    ocfg = common.oracle_config(cfg)
    ref = O.default_config()
    for name, _ in O.Config._fields_:
        a, b = getattr(ocfg, name), getattr(ref, name)
        assert (list(a) == list(b)) if name == "child_order" else (a == b), name
    assert hh.hh_num_params(C.byref(cfg)) == O.num_params(ref) == 74881
    assert C.sizeof(capi.MzConfig) == C.sizeof(O.Config) + 8


def test_init_weights_and_networks_bit_exact(hh):
    cfg = common.product_config(); ocfg = common.oracle_config(cfg)
    blob = np.zeros(74881, np.float32); hh.hh_init_weights(C.byref(cfg), 1337, _p(blob))
    assert np.array_equal(blob, O.init_weights(ocfg, 1337))
    rng = np.random.default_rng(0)
    for _ in range(8):
        st = rng.normal(size=63).astype(np.float32)
        h = np.zeros(27, np.float32); hh.hh_nn(C.byref(cfg), _p(blob), 0, _p(st), _p(h), None)
        assert np.array_equal(h, O.representation(ocfg, blob, st))
        v = np.zeros(1, np.float32); p = np.zeros(9, np.float32); hh.hh_nn(C.byref(cfg), _p(blob), 1, _p(h), _p(v), _p(p))
        ov, op = O.prediction(ocfg, blob, h)
        assert v[0] == ov and np.array_equal(p, op)
        sa = rng.normal(size=36).astype(np.float32)
        nh = np.zeros(27, np.float32); r = np.zeros(1, np.float32); hh.hh_nn(C.byref(cfg), _p(blob), 2, _p(sa), _p(nh), _p(r))
        oh, orr = O.dynamics(ocfg, blob, sa)
        assert np.array_equal(nh, oh) and r[0] == orr


@pytest.mark.parametrize("S,eps,tie", [(10, 0.0, 0), (50, 0.0, 0), (50, 0.25, 0), (25, 0.0, 1)])
def test_run_mcts_bit_exact(hh, S, eps, tie):
    cfg = common.product_config(num_iters=S, exploration_eps=eps, tie_mode=tie); ocfg = common.oracle_config(cfg)
    blob = O.init_weights(ocfg, 7)
    st, legal, tp = common.random_stacked(ocfg, 24, seed=S)
    for i in range(len(st)):
        vc = np.zeros(9, np.int32); rv = np.zeros(1, np.float32); pri = np.zeros(9, np.float32)
        assert hh.hh_run_mcts(C.byref(cfg), _p(blob), _p(st[i]), int(legal[i]), int(tp[i]), 1, 100 + i, 1 + i % 9, _p(vc, C.c_int32), _p(rv), _p(pri)) == 0
        ovc, orv, opri = O.run_mcts(ocfg, blob, st[i], int(legal[i]), int(tp[i]), True, 100 + i, 1 + i % 9)
        assert vc.tolist() == ovc.tolist()
        assert rv[0] == orv
        assert np.array_equal(pri, opri)


@pytest.mark.parametrize("temperature,eps,thr", [(1.0, 0.0, -1), (0.0, 0.0, -1), (1.0, 0.25, -1), (0.5, 0.25, -1), (1.0, 0.25, 3), (1.0, 0.0, 0)])
def test_self_play_bit_exact(hh, temperature, eps, thr):
    cfg = common.product_config(num_iters=10, exploration_eps=eps, temperature_threshold=thr); ocfg = common.oracle_config(cfg)
    blob = O.init_weights(ocfg, 3)
    n = 24; s = O.sizes(ocfg)
    T = np.zeros(n, np.int32); obs = np.zeros((n, s["Tmax"], s["obs"]), np.float32); act = np.zeros((n, s["Tmax"]), np.int32)
    rew = np.zeros((n, s["Tmax"]), np.float32); tp = np.zeros((n, s["Tmax"]), np.int32)
    cv = np.zeros((n, s["Tmax"], 9), np.float32); rv = np.zeros((n, s["Tmax"]), np.float32)
    sims = hh.hh_self_play(C.byref(cfg), _p(blob), 40, n, temperature, _p(T, C.c_int32), _p(obs), _p(act, C.c_int32), _p(rew), _p(tp, C.c_int32), _p(cv), _p(rv))
    o = O.self_play(ocfg, blob, 40, n, temperature, 2)
    assert sims == o["sims"]
    got = dict(T=T, obs=obs, actions=act, rewards=rew, to_play=tp, child_visits=cv, root_values=rv)
    for k in common.HIST_KEYS:
        assert np.array_equal(got[k], o[k]), k


def test_get_batch_bit_exact(hh):
    cfg = common.product_config(exploration_eps=0.25); ocfg = common.oracle_config(cfg)
    blob = O.init_weights(ocfg, 5)
    h = O.self_play(ocfg, blob, 0, 40, 1.0, 2)
    s = O.sizes(ocfg); B = cfg.batch_size
    for step in (1, 2, 77):
        ob = O.get_batch(ocfg, h, step, first_key=1)
        idx = np.zeros((B, 2), np.int32); obs = np.zeros((B, s["stack"]), np.float32); act = np.zeros((B, s["K1"]), np.float32)
        val = np.zeros((B, s["K1"]), np.float32); rew = np.zeros((B, s["K1"]), np.float32); pol = np.zeros((B, s["K1"], 9), np.float32)
        gs = np.zeros(B, np.float32)
        assert hh.hh_get_batch(C.byref(cfg), 40, 1, _p(h["T"], C.c_int32), _p(h["obs"]), _p(h["actions"], C.c_int32), _p(h["rewards"]),
                               _p(h["to_play"], C.c_int32), _p(h["child_visits"]), _p(h["root_values"]), step, _p(idx, C.c_int32),
                               _p(obs), _p(act), _p(val), _p(rew), _p(pol), _p(gs)) == 0
        got = dict(index=idx, obs=obs, actions=act, values=val, rewards=rew, policies=pol, gscale=gs)
        for k in common.BATCH_KEYS:
            assert np.array_equal(got[k], ob[k]), k


# ---- ResNet: the host-built step program + weight image, replayed by a scalar interpreter, vs the bf16-emulating oracle ----
@pytest.mark.parametrize("kw", [dict(), dict(rn_kernel=1, rn_num_blocks=1, rn_num_filters=32), dict(game=1, W=6, H=7, A=7, max_moves=42, rn_num_blocks=1)])
def test_resnet_program_replay_matches_bf16_oracle(hh, kw):
    from muzero_jl_b200 import capi
    from test_oracle_resnet import _randomised_blob
    cfg = capi.resnet_config(**kw)
    if kw.get("game") == 1:
        cfg = capi.connect_config(**{k: v for k, v in kw.items() if k not in ("game", "W", "H", "A", "max_moves")})
    ocfg = common.oracle_config(cfg)
    info = np.zeros(8, np.int32)
    assert hh.hh_rn_program_info(C.byref(cfg), _p(info, C.c_int32)) == 0
    assert info[0] == info[1] + info[2] + info[3] and info[6] == 4 * (128 // (cfg.W * cfg.H)) and info[5] >= 8192 + 768
    assert hh.hh_rn_num_params(C.byref(cfg)) == O.num_params(ocfg)
    blob = _randomised_blob(ocfg, 17)
    s = O.sizes(ocfg); n = 5
    rng = np.random.default_rng(3)
    st = rng.integers(0, 2, (n, s["stack"])).astype(np.float32)
    O.set_bf16(True)
    try:
        oh = np.stack([O.representation(ocfg, blob, x) for x in st])
        ov, op = zip(*[O.prediction(ocfg, blob, x) for x in oh])
        sa = np.concatenate([2 * oh, np.repeat(((np.arange(n) % cfg.A + 1) / np.float32(cfg.A)).astype(np.float32)[:, None], cfg.W * cfg.H, 1)], 1).astype(np.float32)
        onh, orr = zip(*[O.dynamics(ocfg, blob, x) for x in sa])
    finally:
        O.set_bf16(False)
    h = np.zeros((n, s["hidden"]), np.float32); dummy = np.zeros(n, np.float32)
    assert hh.hh_rn_forward(C.byref(cfg), _p(blob), 0, n, _p(st), _p(h), _p(dummy)) == 0
    scale = max(1.0, float(np.max(np.abs(oh))))
    assert np.max(np.abs(h - oh)) < 2e-2 * scale and np.median(np.abs(h - oh)) < 1e-3
    v = np.zeros(n, np.float32); p = np.zeros((n, cfg.A), np.float32)
    assert hh.hh_rn_forward(C.byref(cfg), _p(blob), 1, n, _p(np.ascontiguousarray(oh)), _p(v), _p(p)) == 0
    assert np.max(np.abs(v - np.array(ov))) < 2e-2 and np.max(np.abs(p - np.stack(op))) < 2e-2
    nh = np.zeros((n, s["hidden"]), np.float32); r = np.zeros(n, np.float32)
    assert hh.hh_rn_forward(C.byref(cfg), _p(blob), 2, n, _p(sa), _p(nh), _p(r)) == 0
    onh = np.stack(onh)
    assert np.max(np.abs(nh - onh)) < 2e-2 * max(1.0, float(np.max(np.abs(onh)))) and np.max(np.abs(r - np.array(orr))) < 2e-2


def _randomise_batchnorm(ocfg, blob, seed):
    """give every BatchNorm non-trivial beta / gamma / mu / sigma2 (a fresh one is the identity up to 1/sqrt(1 + 1f-5))"""
    rng = np.random.default_rng(seed)
    mask = O.trainable_mask(ocfg)
    stats = np.flatnonzero(mask == 0)                      # mu, sigma2 blocks: [mu(out) | sigma2(out)] per BatchNorm
    assert len(stats) > 0 and len(stats) % 2 == 0
    out = blob.copy()
    # beta, gamma sit right before mu in the blob: walk the runs of statistics
    runs = np.split(stats, np.flatnonzero(np.diff(stats) > 1) + 1)
    for r in runs:
        n = len(r) // 2
        out[r[0] - 2 * n:r[0] - n] = rng.normal(0, 0.3, n)            # beta
        out[r[0] - n:r[0]] = rng.uniform(0.5, 1.5, n)                 # gamma
        out[r[:n]] = rng.normal(0, 0.3, n)                            # mu
        out[r[n:]] = rng.uniform(0.3, 2.0, n)                         # sigma2
    return out.astype(np.float32)


def test_use_batch_norm_networks_and_search_bit_exact(hh):
    """FeedForwardHP.use_batch_norm (Learning.jl:70-79): Dense + BatchNorm(relu) in test mode, blob W, b, beta, gamma, mu, sigma2"""
    cfg = common.product_config(use_batch_norm=1, num_iters=20, exploration_eps=0.25); ocfg = common.oracle_config(cfg)
    n = hh.hh_num_params(C.byref(cfg))
    assert n == O.num_params(ocfg) == 74881 + 4 * 64 * 18          # 18 make_dense layers of width 64
    blob = np.zeros(n, np.float32); hh.hh_init_weights(C.byref(cfg), 5, _p(blob))
    assert np.array_equal(blob, O.init_weights(ocfg, 5))
    blob = _randomise_batchnorm(ocfg, blob, 1)
    rng = np.random.default_rng(2)
    for _ in range(4):
        st = rng.normal(size=63).astype(np.float32)
        h = np.zeros(27, np.float32); hh.hh_nn(C.byref(cfg), _p(blob), 0, _p(st), _p(h), None)
        assert np.array_equal(h, O.representation(ocfg, blob, st))
        v = np.zeros(1, np.float32); p = np.zeros(9, np.float32); hh.hh_nn(C.byref(cfg), _p(blob), 1, _p(h), _p(v), _p(p))
        ov, op = O.prediction(ocfg, blob, h)
        assert v[0] == ov and np.array_equal(p, op)
    # the BatchNorm must matter: the same weights without it give another hidden state
    cfg0 = common.product_config(); ocfg0 = common.oracle_config(cfg0)
    assert not np.array_equal(O.representation(ocfg0, O.init_weights(ocfg0, 5), st), h)
    st, legal, tp = common.random_stacked(ocfg, 6, seed=3)
    for i in range(len(st)):
        vc = np.zeros(9, np.int32); rv = np.zeros(1, np.float32); pri = np.zeros(9, np.float32)
        hh.hh_run_mcts(C.byref(cfg), _p(blob), _p(st[i]), int(legal[i]), int(tp[i]), 1, 40 + i, 1, _p(vc, C.c_int32), _p(rv), _p(pri))
        ovc, orv, opri = O.run_mcts(ocfg, blob, st[i], int(legal[i]), int(tp[i]), True, 40 + i, 1)
        assert vc.tolist() == ovc.tolist() and rv[0] == orv and np.array_equal(pri, opri)
