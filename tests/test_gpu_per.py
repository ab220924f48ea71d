"""GPU tests of prioritised replay (conf.PER = true, repaired specification): priorities at save, prioritised sampling and
importance weights, the weighted loss and its gradient, and the priority update -- all against the oracle.  Integer work
(priorities, sampled indices) and the batch are bit-exact; loss scalars / gradients use the learner's tolerances."""
import numpy as np
import pytest

import common
from oracle import oracle as O

pytestmark = pytest.mark.gpu
LOSS_RTOL = 2e-6
BPTT_RTOL = 2e-5


@pytest.fixture(scope="module")
def capi():
    from muzero_jl_b200 import capi
    return capi


def make(capi, **kw):
    kw.setdefault("num_slots", 64); kw.setdefault("replay_buffer_size", 160); kw.setdefault("per", 1)
    cfg = capi.default_config(**kw)
    return capi.Context(cfg), common.oracle_config(cfg)


@pytest.mark.parametrize("alpha", [1, 2])
def test_per_priorities_sampling_and_weights_bit_exact(capi, alpha):
    ctx, ocfg = make(capi, batch_size=96, per_alpha=alpha)
    ctx.init_weights(5)
    ctx.self_play(0, 150, 1.0)
    h = ctx.history_export()
    q_pos, q_game = ctx.replay_priorities()
    oq_pos, oq_game = O.per_priorities(ocfg, h)
    assert np.array_equal(q_pos, oq_pos) and np.array_equal(q_game, oq_game)
    for step in (1, 2, 9):
        b = ctx.get_batch_per(step)
        ob = O.get_batch_per(ocfg, h, oq_pos, oq_game, step)
        for k in common.BATCH_KEYS + ("weights",):
            assert np.array_equal(b[k], ob[k]), (step, k)
        assert b["weights"].max() == 1.0
    # eviction: the ring wraps (capacity 160), priorities follow their games
    ctx.self_play(150, 64, 1.0)
    h2 = ctx.history_export()
    q2, g2 = ctx.replay_priorities()
    o2, og2 = O.per_priorities(ocfg, h2)
    assert np.array_equal(q2, o2) and np.array_equal(g2, og2)
    info = ctx.replay_info()
    b = ctx.get_batch_per(3); ob = O.get_batch_per(ocfg, h2, o2, og2, 3, first_key=info["first_key"])
    for k in common.BATCH_KEYS + ("weights",):
        assert np.array_equal(b[k], ob[k]), k
    ctx.close()


def test_per_weighted_loss_and_gradients(capi):
    ctx, ocfg = make(capi, batch_size=40, intermediate_rewards=1)
    ctx.init_weights(6)
    rng = np.random.default_rng(4)
    blob = ctx.get_weights() + (rng.standard_normal(ctx.num_params()) * 0.02).astype(np.float32)
    ctx.set_weights(blob)
    ctx.self_play(0, 100, 1.0)
    b = ctx.get_batch_per(2)
    b["rewards"] = b["rewards"] + (rng.standard_normal(b["rewards"].shape) * 0.3).astype(np.float32)
    g, losses = ctx.learn_gradients(b, capi.GRAD_BPTT)
    _, _, _, ol = O.learn_forward_w(ocfg, blob, b)
    assert np.allclose(losses, ol, rtol=LOSS_RTOL)
    _, og = O.learn_gradients_w(ocfg, blob, b, fwd64=False)
    nr, npred = O.num_params(ocfg, 0), O.num_params(ocfg, 1)
    for lo, hi in ((0, nr), (nr, nr + npred), (nr + npred, g.shape[0])):
        assert np.max(np.abs(g[lo:hi] - og[lo:hi])) <= BPTT_RTOL * np.max(np.abs(og[lo:hi]))
    # the weights matter
    b1 = dict(b); b1["weights"] = np.ones_like(b["weights"])
    g1, l1 = ctx.learn_gradients(b1, capi.GRAD_BPTT)
    assert not np.array_equal(l1, losses)
    ctx.close()


@pytest.mark.parametrize("mode", ["l2", "bptt"])
def test_per_training_loop_updates_priorities_like_the_oracle(capi, mode):
    """learning! with PER: batch (prioritised) -> update -> update_priorities!, several steps; priorities tracked by the oracle."""
    gm = capi.GRAD_BPTT if mode == "bptt" else capi.GRAD_REFERENCE_L2
    ctx, ocfg = make(capi, batch_size=48)
    ctx.init_weights(8)
    ctx.self_play(0, 120, 1.0)
    h = ctx.history_export()
    oq_pos, oq_game = O.per_priorities(ocfg, h)
    for t in range(1, 6):
        ob = O.get_batch_per(ocfg, h, oq_pos, oq_game, t)
        b = ctx.get_batch_per(t)
        for k in common.BATCH_KEYS + ("weights",):
            assert np.array_equal(b[k], ob[k]), (t, k)
        pv, _, _, _ = ctx.learn_forward(b)                        # predictions with the weights BEFORE the update (Learning.jl:347-374, 401)
        ctx.learn_step(t, gm)
        O.per_update(ocfg, h, oq_pos, oq_game, ob["index"], pv, ob["values"])
        q_pos, q_game = ctx.replay_priorities()
        assert np.array_equal(q_pos, oq_pos) and np.array_equal(q_game, oq_game), t
    ctx.close()


def test_per_off_is_unchanged(capi):
    ctx, ocfg = make(capi, per=0, batch_size=32)
    ctx.init_weights(9); ctx.self_play(0, 64, 1.0)
    with pytest.raises(capi.MuZeroB200Error):
        ctx.get_batch_per(1)
    b = ctx.get_batch(1)
    ob = O.get_batch(ocfg, ctx.history_export(), step=1)
    for k in common.BATCH_KEYS:
        assert np.array_equal(b[k], ob[k])
    ctx.close()
