"""GPU tests of the ResNet networks on the tensor cores (net_type = MZ_NET_RESNET, nn_mode = MZ_NN_BF16_TC).
bf16 operands cannot be bit-exact against the Float32 oracle; the kernels are held to the oracle's bf16 emulation
(same rounding points: every stored activation and every weight is bfloat16, sums are Float32) with a stated tolerance,
and the searches to a visit-count agreement rate."""
import numpy as np
import pytest

import common
from oracle import oracle as O
from test_oracle_resnet import _randomised_blob

pytestmark = pytest.mark.gpu
RN_ATOL = 2e-2      # one bf16 ulp of an O(1) activation is 4e-3; a few layers of re-rounding on top of tensor-core summation order


@pytest.fixture(scope="module")
def capi():
    from muzero_jl_b200 import capi
    return capi


def make(capi, **kw):
    kw.setdefault("num_slots", 128); kw.setdefault("replay_buffer_size", 512)
    cfg = capi.resnet_config(**kw)
    return capi.Context(cfg), common.oracle_config(cfg)


@pytest.mark.parametrize("kw", [dict(rn_kernel=1, rn_num_blocks=0), dict(rn_kernel=1, rn_num_blocks=1), dict(rn_kernel=1), dict(),
                                dict(rn_num_filters=32, depth_value=0, rn_second_head_filters=3)])
def test_resnet_networks_match_bf16_oracle(capi, kw):
    ctx, ocfg = make(capi, **kw)
    blob = _randomised_blob(ocfg, 7)
    assert ctx.num_params() == blob.shape[0]
    ctx.set_weights(blob)
    assert np.array_equal(ctx.get_weights(), blob)
    st, legal, tp = common.random_stacked(ocfg, 75, seed=4)
    O.set_bf16(True)
    try:
        h = ctx.representation(st)
        oh = np.stack([O.representation(ocfg, blob, x) for x in st])
        err = np.abs(h - oh)
        assert np.max(err) < RN_ATOL * max(1.0, np.max(np.abs(oh))) and np.median(err) < 1e-3, (np.max(err), np.median(err))
        v, p = ctx.prediction(oh)
        ov, op = zip(*[O.prediction(ocfg, blob, x) for x in oh])
        assert np.max(np.abs(v - np.array(ov))) < RN_ATOL and np.max(np.abs(p - np.stack(op))) < RN_ATOL
        sa = np.concatenate([2 * oh, np.repeat((np.arange(75) % 9 + 1)[:, None] / np.float32(9), 9, 1).astype(np.float32)], 1)
        nh, r = ctx.dynamics(sa)
        onh, orr = zip(*[O.dynamics(ocfg, blob, x) for x in sa])
        onh = np.stack(onh)
        assert np.max(np.abs(nh - onh)) < RN_ATOL * max(1.0, np.max(np.abs(onh))) and np.max(np.abs(r - np.array(orr))) < RN_ATOL
    finally:
        O.set_bf16(False)
    ctx.close()


def test_resnet_mcts_agrees_with_bf16_oracle(capi):
    """Visit counts of whole searches: identical to the bf16-emulating oracle for most roots (the remaining ones differ by a
    near-tie in PUCT that the summation order inside the tensor core decides)."""
    ctx, ocfg = make(capi, num_iters=30, exploration_eps=0.0)
    blob = _randomised_blob(ocfg, 11)
    ctx.set_weights(blob)
    n = 120
    st, legal, tp = common.random_stacked(ocfg, n, seed=21)
    gid = np.arange(n, dtype=np.uint64) + 1000; mv = np.ones(n, np.int32)
    vc, rv, pri = ctx.run_mcts(st, legal, tp, False, gid, mv, priors=True)
    assert np.all(vc.sum(1) == 30)
    for i in range(n):
        assert not np.any(vc[i][[(legal[i] >> a) & 1 == 0 for a in range(9)]])
    O.set_bf16(True)
    try:
        same = 0; top = 0; perr = 0.0
        for i in range(n):
            ovc, orv, opri = O.run_mcts(ocfg, blob, st[i], int(legal[i]), int(tp[i]), False, int(gid[i]), 1)
            same += int(np.array_equal(ovc, vc[i])); top += int(np.argmax(ovc) == np.argmax(vc[i]))
            perr = max(perr, float(np.max(np.abs(opri - pri[i]))))
    finally:
        O.set_bf16(False)
    assert perr < RN_ATOL, perr
    assert same >= 0.85 * n and top >= 0.9 * n, (same, top, n)
    ctx.close()


def test_resnet_self_play_is_well_formed_and_deterministic(capi):
    hs = []
    for _ in range(2):
        ctx, ocfg = make(capi, num_iters=16, num_slots=100)
        ctx.init_weights(3)
        sims, moves = ctx.self_play(0, 150, 1.0)
        info = ctx.replay_info()
        assert info["n_games"] == 150 and sims == moves * 16
        h = ctx.history_export()
        hs.append(h)
        for j in range(150):
            T = h["T"][j]
            assert 5 <= T <= 9 and np.allclose(h["child_visits"][j, :T].sum(1), 1.0, atol=1e-6) and np.all(h["rewards"][j, :T - 1] == 0)
        ctx.close()
    o0, o1 = np.argsort(hs[0]["game_id"]), np.argsort(hs[1]["game_id"])
    for k in common.HIST_KEYS:
        assert np.array_equal(hs[0][k][o0], hs[1][k][o1]), k


def test_resnet_learner_reference_l2(capi):
    """learning! with the ResNet networks in the reference's own semantics (Q20: the gradient of every Flux parameter is 2 * theta; the BatchNorm
    statistics are not parameters): the K-step unroll runs as bf16 inference on the tensor cores, so predictions and losses are held to the
    bf16-emulating oracle with the networks' tolerance, while the update does not depend on the forward pass and must be BIT-exact."""
    ctx, ocfg = make(capi, batch_size=48)
    blob = _randomised_blob(ocfg, 11)
    ctx.set_weights(blob)
    ctx.self_play(0, 150, 1.0)
    batch = ctx.get_batch(3)
    batch.pop("index", None)
    pv, pr, pp, losses = ctx.learn_forward(batch)
    O.set_bf16(True)
    try:
        opv, opr, opp, ol = O.learn_forward(ocfg, blob, batch)
    finally:
        O.set_bf16(False)
    assert np.max(np.abs(pv - opv)) <= RN_ATOL and np.max(np.abs(pr - opr)) <= RN_ATOL and np.max(np.abs(pp - opp)) <= RN_ATOL
    assert np.array_equal(pv[:, 0], pv[:, 1]) and np.array_equal(pp[:, 0], pp[:, 1]) and np.all(pr[:, 0] == 0)        # Q19, Learning.jl:352
    assert np.allclose(losses, ol, rtol=2e-2)
    # gradient = 2 * theta on Flux.params, 0 on the running statistics
    mask = O.trainable_mask(ocfg).astype(bool)
    assert 0 < (~mask).sum() < mask.size
    g, _ = ctx.learn_gradients(batch, capi.GRAD_REFERENCE_L2)
    assert np.array_equal(g[mask], (blob + blob)[mask]) and np.all(g[~mask] == 0)
    # three learning! iterations: weights bit-identical to the oracle's ADAM on that gradient; statistics untouched
    w = blob.copy(); m = np.zeros_like(w); v = np.zeros_like(w)
    for t in (1, 2, 3):
        ctx.learn_step(t, capi.GRAD_REFERENCE_L2, batch)
        O.learn_step(ocfg, w, m, v, t, batch)
    got = ctx.get_weights()
    assert np.array_equal(got, w) and np.array_equal(got[~mask], blob[~mask]) and not np.array_equal(got[mask], blob[mask])
    # the networks (and the search) now run on the updated weights: the bf16 image was rebuilt
    st, legal, tp = common.random_stacked(ocfg, 20, seed=2)
    O.set_bf16(True)
    try:
        oh = np.stack([O.representation(ocfg, w, x) for x in st])
    finally:
        O.set_bf16(False)
    h = ctx.representation(st)
    assert np.max(np.abs(h - oh)) < RN_ATOL * max(1.0, float(np.max(np.abs(oh))))
    # the library's own sampling path + optimiser checkpoint
    ctx.learn_steps(4, 2)
    ck = ctx.checkpoint()
    assert int(ck["steps_done"]) == 5 and np.all(ck["adam_m"][~mask] == 0) and np.any(ck["adam_m"][mask] != 0)
    wa = ctx.learn_steps(6, 2); after = ctx.get_weights()
    ctx.restore(ck)                                                       # weights, moments, step count, replay: the next steps repeat bit for bit
    wb = ctx.learn_steps(6, 2)
    assert np.array_equal(ctx.get_weights(), after) and np.array_equal(wa, wb)
    with pytest.raises(capi.MuZeroB200Error) as e:
        ctx.learn_step(8, capi.GRAD_BPTT, batch)                         # the backward through the convolution towers is not built
    assert e.value.code == capi.E_UNSUPPORTED
    ctx.close()


def test_golden_resnet_and_per_fixture(capi):
    """The committed fixture tests/golden/bptt_per_resnet.npz against the CUDA path: prioritised batch and priority update bit-exact,
    BPTT gradient and ResNet outputs within their tolerances, ResNet visit counts."""
    import os
    g = np.load(os.path.join(common.ROOT, "tests", "golden", "bptt_per_resnet.npz"))
    ctx = capi.Context(capi.default_config(num_slots=64, replay_buffer_size=64, batch_size=24, per=1, intermediate_rewards=1))
    ctx.init_weights(77)
    ctx.history_import({k: g["hist_" + k] for k in common.HIST_KEYS})
    q_pos, q_game = ctx.replay_priorities()
    assert np.array_equal(q_pos, g["q_pos"]) and np.array_equal(q_game, g["q_game"])
    pb = ctx.get_batch_per(3)
    for k in common.BATCH_KEYS + ("weights",):
        assert np.array_equal(pb[k], g["pb_" + k]), k
    grad, losses = ctx.learn_gradients(pb, capi.GRAD_BPTT)
    assert np.max(np.abs(grad - g["grad"])) <= 2e-5 * np.max(np.abs(g["grad"])) and np.allclose(losses, g["losses"], rtol=2e-6)
    ctx.learn_step(3, capi.GRAD_REFERENCE_L2)
    q2, g2 = ctx.replay_priorities()
    assert np.array_equal(q2, g["q_pos_after"]) and np.array_equal(g2, g["q_game_after"])
    ctx.close()
    rctx, ocfg = make(capi, num_iters=20, exploration_eps=0.0)
    rctx.init_weights(5)
    h = rctx.representation(g["rn_stacked"])
    assert np.max(np.abs(h - g["rn_hidden_bf16"])) < RN_ATOL * max(1.0, float(np.max(np.abs(g["rn_hidden_bf16"]))))
    v, p = rctx.prediction(g["rn_hidden_bf16"])
    assert np.max(np.abs(v - g["rn_value_bf16"])) < RN_ATOL and np.max(np.abs(p - g["rn_policy_bf16"])) < RN_ATOL
    vc, rv = rctx.run_mcts(g["rn_stacked"], g["rn_legal"], g["rn_to_play"], False, np.arange(12, dtype=np.uint64) + 500, np.ones(12, np.int32))
    assert (vc == g["rn_visit_counts_bf16"]).all(1).sum() >= 10
    rctx.close()
