"""FeedForwardHP.use_batch_norm (src/Constructors.jl:71, src/Learning.jl:70-79): every make_dense layer is Dense + BatchNorm(relu).  The
reference never differentiates a forward pass (Q20) and never calls trainmode!, so BatchNorm runs in test mode everywhere; its beta / gamma
are Flux parameters (they take part in sum(abs2, theta) and in the ADAM update), the running statistics are not.  Exact fp32 path, bit-exact
against the oracle; split-precision tensor-core path (BatchNorm folded into the weight image): within that path's tolerance."""
import numpy as np
import pytest

import common
from oracle import oracle as O
from test_host_harness import _randomise_batchnorm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from muzero_jl_b200 import capi
    return capi


def make_ctx(capi, **kw):
    kw.setdefault("num_slots", 256); kw.setdefault("replay_buffer_size", 1024); kw.setdefault("use_batch_norm", 1)
    cfg = capi.default_config(**kw)
    return capi.Context(cfg), common.oracle_config(cfg)


def test_bn_networks_and_run_mcts_bit_exact(capi):
    ctx, ocfg = make_ctx(capi, num_iters=25, exploration_eps=0.25)
    ctx.init_weights(9)
    assert np.array_equal(ctx.get_weights(), O.init_weights(ocfg, 9))
    blob = _randomise_batchnorm(ocfg, ctx.get_weights(), 4); ctx.set_weights(blob)
    assert np.array_equal(ctx.get_weights(), blob)
    st, legal, tp = common.random_stacked(ocfg, 300, seed=6)
    h = ctx.representation(st); v, p = ctx.prediction(h)
    oh = np.stack([O.representation(ocfg, blob, x) for x in st])
    assert np.array_equal(h, oh)
    for i in range(0, 300, 37):
        ov, op = O.prediction(ocfg, blob, oh[i])
        assert v[i] == ov and np.array_equal(p[i], op)
    game = np.arange(300, dtype=np.uint64) + 7; move = np.ones(300, np.int32)
    vc, rv, pri = ctx.run_mcts(st, legal, tp, True, game, move, priors=True)     # 300 roots: the batched kernel mz_k_search
    for i in range(0, 300, 7):
        ovc, orv, opri = O.run_mcts(ocfg, blob, st[i], int(legal[i]), int(tp[i]), True, int(game[i]), 1)
        assert vc[i].tolist() == ovc.tolist() and rv[i] == orv and np.array_equal(pri[i], opri), i
    vc1, rv1, _ = ctx.run_mcts(st[:5], legal[:5], tp[:5], True, game[:5], move[:5], priors=True)   # few roots: still the batched kernel (the resident-weights kernel has no BatchNorm)
    assert np.array_equal(vc1, vc[:5]) and np.array_equal(rv1, rv[:5])
    ctx.close()


def test_bn_self_play_and_learner_bit_exact(capi):
    ctx, ocfg = make_ctx(capi, num_slots=64, replay_buffer_size=256, num_iters=12)
    ctx.init_weights(3)
    blob = _randomise_batchnorm(ocfg, ctx.get_weights(), 8); ctx.set_weights(blob)
    sims, moves = ctx.self_play(100, 150, 1.0)
    o = O.self_play(ocfg, blob, 100, 150, 1.0, 4)
    assert sims == o["sims"]
    h = ctx.history_export()
    for j in range(150):
        i = int(h["game_id"][j]) - 100
        for k in common.HIST_KEYS:
            assert np.array_equal(h[k][j], o[k][i]), (k, i)
    # learning! (Learning.jl:327-397): losses, then ADAM on 2 * theta over Flux.params: beta and gamma move, mu and sigma2 do not
    mask = O.trainable_mask(ocfg)
    w = blob.copy(); m = np.zeros_like(w); v = np.zeros_like(w)
    for t in (1, 2, 3):
        batch = ctx.get_batch(t)
        losses = ctx.learn_step(t)
        ol = O.learn_step(ocfg, w, m, v, t, batch)
        np.testing.assert_allclose(losses, ol, rtol=2e-6)
        got = ctx.get_weights()
        assert np.array_equal(got, w), t
        assert np.array_equal(got[mask == 0], blob[mask == 0])
    assert not np.array_equal(got[mask == 1], blob[mask == 1])
    ctx.close()


def test_bn_on_the_split_precision_path(capi):
    """MZ_NN_SPLIT_MMA: mz_k_pack_images folds gamma / sqrt(sigma2 + 1f-5) into W and (b - mu, beta) into the bias, so the tcgen05 kernels run
    BatchNorm networks unchanged: network outputs within 2e-5 * max(1, |y|) of the Float32 oracle, run_mcts visit counts identical on >= 99 %
    of the roots, the learner's update (2 * theta over Flux.params) bit-exact, and the image follows the update."""
    S, n = 50, 512
    ctx, ocfg = make_ctx(capi, nn_mode=capi.NN_SPLIT_MMA, num_iters=S, exploration_eps=0.25, num_slots=512)
    ctx.init_weights(9)
    blob = _randomise_batchnorm(ocfg, ctx.get_weights(), 4); ctx.set_weights(blob)
    st, legal, tp = common.random_stacked(ocfg, n, seed=6)
    h = ctx.representation(st)
    oh = np.stack([O.representation(ocfg, blob, x) for x in st])
    assert np.max(np.abs(h - oh)) <= 2e-5 * max(1.0, float(np.max(np.abs(oh))))
    v, p = ctx.prediction(oh)
    ov, op = zip(*[O.prediction(ocfg, blob, x) for x in oh])
    assert np.max(np.abs(v - np.array(ov))) <= 2e-5 and np.max(np.abs(p - np.stack(op))) <= 2e-5
    sa = np.concatenate([oh * 2, np.full((n, 9), np.float32(4.0 / 9.0), np.float32)], axis=1)
    nh, r = ctx.dynamics(sa)
    onh, orr = zip(*[O.dynamics(ocfg, blob, x) for x in sa])
    assert np.max(np.abs(nh - np.stack(onh))) <= 2e-5 * max(1.0, float(np.max(np.abs(np.stack(onh))))) and np.max(np.abs(r - np.array(orr))) <= 2e-5
    game = np.arange(n, dtype=np.uint64) + 7; move = (np.arange(n) % 9 + 1).astype(np.int32)
    vc, rv, _ = ctx.run_mcts(st, legal, tp, True, game, move, priors=True)
    exact = [O.run_mcts(ocfg, blob, st[i], int(legal[i]), int(tp[i]), True, int(game[i]), int(move[i])) for i in range(n)]
    same = np.mean([vc[i].tolist() == exact[i][0].tolist() for i in range(n)])
    print("BatchNorm networks on the split-precision path: visit counts identical to the Float32 oracle on %.4f of %d roots" % (same, n))
    assert np.all(vc.sum(1) == S) and same >= 0.99
    # self-play + learning!: the update does not depend on the forward pass (Q20), so the weights stay bit-exact; the losses carry the path's 1e-6
    ctx.self_play(100, 600, 1.0)
    w = blob.copy(); m = np.zeros_like(w); vv = np.zeros_like(w)
    for t in (1, 2):
        batch = ctx.get_batch(t)
        losses = ctx.learn_step(t)
        ol = O.learn_step(ocfg, w, m, vv, t, batch)
        np.testing.assert_allclose(losses, ol, rtol=1e-4)
        assert np.array_equal(ctx.get_weights(), w), t
    h2 = ctx.representation(st[:64])                         # the image was rebuilt from the updated beta / gamma / W / b
    oh2 = np.stack([O.representation(ocfg, w, x) for x in st[:64]])
    assert np.max(np.abs(h2 - oh2)) <= 2e-5 * max(1.0, float(np.max(np.abs(oh2)))) and not np.array_equal(oh2, oh[:64])
    # MZ_GRAD_BPTT in such a context: forward + backward on the tensor cores (the images ARE the folded layers) + the BatchNorm chain rule in the
    # reduce kernel: the tensor-core backward's tolerance (bf16 operands: 1e-2 of the largest entry per network, tests/test_gpu_mma.py)
    assert ctx.learner_path(capi.GRAD_BPTT) == 2
    batch = ctx.get_batch(3)
    g, _ = ctx.learn_gradients(batch, capi.GRAD_BPTT)
    _, og = O.learn_gradients(ocfg, w, batch, fwd64=False)
    mask = O.trainable_mask(ocfg)
    assert np.all(g[mask == 0] == 0)
    nr, npred = O.num_params(ocfg, 0), O.num_params(ocfg, 1)
    worst = 0.0
    for lo, hi in ((0, nr), (nr, nr + npred), (nr + npred, g.shape[0])):
        worst = max(worst, float(np.max(np.abs(g[lo:hi] - og[lo:hi])) / np.max(np.abs(og[lo:hi]))))
    print("tensor-core BPTT through BatchNorm layers: worst per-network gradient error %.2e of the network's largest entry" % worst)
    assert worst <= 1e-2
    ctx.close()


@pytest.mark.parametrize("kw,B", [({}, 32), ({"intermediate_rewards": 1}, 77), ({"num_unroll_steps": 2, "intermediate_rewards": 1, "depth_value": 2, "depth_reward": 0}, 40),
                                  ({"stacked_observations": 2, "intermediate_rewards": 1}, 33)])
def test_bn_bptt_gradients_match_oracle(capi, kw, B):
    """MZ_GRAD_BPTT with use_batch_norm: mz_k_learn_bptt on the folded weights (mz_k_bn_fold) + the chain rule back to W, b, beta, gamma
    (mz_k_grad_reduce_bn) against the oracle's Float64 backward through the BatchNorm layers (pinned by finite differences, test_oracle_kat.py)"""
    ctx, ocfg = make_ctx(capi, batch_size=B, **kw)
    ctx.init_weights(5)
    rng = np.random.default_rng(8)
    blob = _randomise_batchnorm(ocfg, ctx.get_weights() + (rng.standard_normal(ctx.num_params()) * 0.02).astype(np.float32), 4)
    ctx.set_weights(blob)
    hist = O.self_play(ocfg, blob, 0, 16, 1.0, 2)
    c2 = O.Config.from_buffer_copy(ocfg); c2.batch_size = B
    batch = O.get_batch(c2, hist, step=3)
    batch["rewards"] = batch["rewards"] + (rng.standard_normal(batch["rewards"].shape) * 0.3).astype(np.float32)
    g, losses = ctx.learn_gradients(batch, capi.GRAD_BPTT)
    _, og = O.learn_gradients(ocfg, blob, batch, fwd64=False)
    mask = O.trainable_mask(ocfg)
    assert np.all(g[mask == 0] == 0) and np.all(og[mask == 0] == 0)
    nr, npred = O.num_params(ocfg, 0), O.num_params(ocfg, 1)
    for lo, hi in ((0, nr), (nr, nr + npred), (nr + npred, g.shape[0])):
        scale = np.max(np.abs(og[lo:hi])); err = np.max(np.abs(g[lo:hi].astype(np.float64) - og[lo:hi]))
        assert err <= 2e-5 * scale, (lo, hi, err, scale)
    _, _, _, ol = O.learn_forward(ocfg, blob, batch)
    assert np.allclose(losses, ol, rtol=1e-5)                       # the folded forward is the BatchNorm forward up to rounding
    g2, _ = ctx.learn_gradients(batch, capi.GRAD_BPTT)
    assert np.array_equal(g, g2)
    # the update = Flux.ADAM on that gradient, bit for bit; mu and sigma2 stay put
    m = np.zeros_like(blob); v = np.zeros_like(blob); ob = blob.copy()
    ctx.learn_step(1, capi.GRAD_BPTT, batch)
    O.adam_apply(ob, m, v, g, 1)
    got = ctx.get_weights()
    assert np.array_equal(got, ob) and np.array_equal(got[mask == 0], blob[mask == 0])
    ctx.close()


def test_bn_unsupported_combinations_say_so(capi):
    with pytest.raises(capi.MuZeroB200Error):
        capi.Context(capi.default_config(use_batch_norm=1, nn_mode=capi.NN_BF16_TC))


def test_bn_through_the_reference_level_api(capi):
    from muzero_jl_b200 import api
    conf = api.Config(); hyper = api.FeedForwardHP(use_batch_norm=True)
    cfg = api.to_mz_config(conf, hyper, num_slots=32)
    assert cfg.use_batch_norm == 1
    assert api.to_mz_config(conf, hyper, num_slots=32, nn_mode=capi.NN_SPLIT_MMA).use_batch_norm == 1
    with pytest.raises(NotImplementedError):
        api.to_mz_config(conf, hyper, num_slots=32, nn_mode=capi.NN_BF16_TC)
