"""GPU parity at the configurations bench.py measures (BASELINE.json configs[1], [2]) -- the save / refill scan, the host loop that
runs one iteration ahead of the device and the idle-CTA exit only show their bugs at scale -- plus conf.temperature_threshold and the
checkpoint round trip of the replay state (priorities, reanalysed values, counters)."""
import numpy as np
import pytest

import common
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from muzero_jl_b200 import capi
    return capi


def _histories_equal(h, o, first_game):
    bad = 0
    for j in range(len(h["T"])):
        i = int(h["game_id"][j]) - first_game
        if not all(np.array_equal(h[k][j], o[k][i]) for k in common.HIST_KEYS):
            bad += 1
    return bad


@pytest.mark.parametrize("eps,temperature", [(0.25, 1.0), (0.0, 0.0)])
def test_configs1_4096_slots_50_sims_every_history_bit_exact(capi, eps, temperature):
    """BASELINE.json configs[1] exactly: 4096 concurrent games x 50 simulations / move; 10000 games = 2.4 waves, so slots are refilled
    and the 8192-entry replay ring evicts.  EVERY GameHistory (and the counters) must equal the oracle's."""
    G, S, games, first = 4096, 50, 10000, 7000
    cfg = capi.default_config(num_slots=G, num_iters=S, replay_buffer_size=8192, exploration_eps=eps)
    ctx = capi.Context(cfg); ocfg = common.oracle_config(ctx.cfg)
    ctx.init_weights(1337); blob = ctx.get_weights()
    sims, moves = ctx.self_play(first, games, temperature)
    o = O.self_play(ocfg, blob, first, games, temperature, 16)
    assert sims == o["sims"] and moves == int(o["T"].sum())
    info = ctx.replay_info()
    assert info["n_games"] == 8192 and info["first_key"] == games - 8192 + 1
    h = ctx.history_export()
    assert len(set(h["game_id"].tolist())) == 8192 and first <= h["game_id"].min() and h["game_id"].max() < first + games
    assert info["total_samples"] == int(h["T"].sum())
    assert _histories_equal(h, o, first) == 0
    c = ctx.replay_counters()
    assert c[0] == games and c[1] == int(o["T"].sum())
    # a second call continues on the same context (idle launch of the previous wave, slots handed out again)
    sims2, _ = ctx.self_play(first + games, 4096, temperature)
    o2 = O.self_play(ocfg, blob, first + games, 4096, temperature, 16)
    assert sims2 == o2["sims"]
    h2 = ctx.history_export(key0=games + 1, n=4096)
    assert _histories_equal(h2, o2, first + games) == 0
    ctx.close()


def test_configs1_split_precision_headline_agrees_with_float32_oracle(capi):
    """The bench's headline path at the bench's size: 4096 slots x 50 simulations with the networks on tcgen05 (bf16 hi + lo operands),
    10000 games = 2.4 waves.  Not bit-exact by construction; asserted: every game is played exactly once and is well formed, the counters
    are right, >= 99.5 % of first plies and >= 98 % of whole games are identical to the FLOAT32 oracle's (measured 99.95 % / 99.3 %)."""
    G, S, games, first = 4096, 50, 10000, 7000
    ctx = capi.Context(capi.default_config(num_slots=G, num_iters=S, replay_buffer_size=16384, nn_mode=capi.NN_SPLIT_MMA))
    ocfg = common.oracle_config(ctx.cfg)
    ctx.init_weights(1337); blob = ctx.get_weights()
    sims, moves = ctx.self_play(first, games, 1.0)
    o = O.self_play(ocfg, blob, first, games, 1.0, 16)
    h = ctx.history_export()
    assert sorted(h["game_id"].tolist()) == list(range(first, first + games)) and sims == moves * S and moves == int(h["T"].sum())
    same_first = same_game = 0
    for j in range(games):
        i = int(h["game_id"][j]) - first
        same_first += int(np.array_equal(h["child_visits"][j, 0], o["child_visits"][i, 0]))
        same_game += int(all(np.array_equal(h[k][j], o[k][i]) for k in ("T", "obs", "actions", "rewards", "to_play", "child_visits")))
    print("split-precision headline config: %.4f of %d first plies, %.4f of whole games identical to the Float32 oracle" % (same_first / games, games, same_game / games))
    assert same_first >= 0.995 * games and same_game >= 0.98 * games
    c = ctx.replay_counters()
    assert c[0] == games and c[1] == moves
    ctx.close()


def test_configs2_resnet_16384_slots_sampled_games_agree_with_bf16_oracle(capi):
    """BASELINE.json configs[2]: 16384 concurrent ResNet games (bf16 tcgen05).  A sample of the games is replayed by the bf16-emulating
    oracle; the executor accumulates in a different order, so agreement is a rate: stated and asserted."""
    G, S, first = 16384, 50, 300000
    ctx = capi.Context(capi.resnet_config(num_slots=G, num_iters=S, replay_buffer_size=G))
    ocfg = common.oracle_config(ctx.cfg)
    ctx.init_weights(1337); blob = ctx.get_weights()
    sims, moves = ctx.self_play(first, G, 1.0)
    assert sims == moves * S
    h = ctx.history_export()
    assert sorted(h["game_id"].tolist()) == list(range(first, first + G))
    by_id = {int(g): j for j, g in enumerate(h["game_id"])}
    assert h["T"].min() >= 5 and h["T"].max() <= 9
    same_game = same_first = n = 0
    O.set_bf16(True)
    try:
        for lo in (0, 5000, 11111, 16384 - 16):                      # 4 blocks of 16 games spread over the slots
            o = O.self_play(ocfg, blob, first + lo, 16, 1.0, 16)
            for i in range(16):
                j = by_id[first + lo + i]; n += 1
                same_first += int(np.array_equal(h["child_visits"][j, 0], o["child_visits"][i, 0]) and h["actions"][j, 0] == o["actions"][i, 0])
                same_game += int(all(np.array_equal(h[k][j], o[k][i]) for k in ("T", "actions", "child_visits")))
    finally:
        O.set_bf16(False)
    print("ResNet 16384-slot wave: %d / %d sampled games identical to the bf16-emulating oracle (all plies), %d / %d first plies" % (same_game, n, same_first, n))
    assert same_first >= 0.9 * n and same_game >= 0.6 * n
    ctx.close()


@pytest.mark.parametrize("thr,temperature", [(0, 1.0), (3, 1.0), (5, 0.5)])
def test_temperature_threshold_bit_exact(capi, thr, temperature):
    """conf.temperature_threshold (SelfPlay.jl:344-346) in self-play and competitive play."""
    cfg = capi.default_config(num_slots=96, replay_buffer_size=512, num_iters=20, temperature_threshold=thr)
    ctx = capi.Context(cfg); ocfg = common.oracle_config(ctx.cfg)
    assert ocfg.temperature_threshold == thr
    ctx.init_weights(5); blob = ctx.get_weights()
    ctx.self_play(100, 250, temperature)
    o = O.self_play(ocfg, blob, 100, 250, temperature, 8)
    h = ctx.history_export()
    assert _histories_equal(h, o, 100) == 0
    ocfg.temperature_threshold = -1
    assert not np.array_equal(O.self_play(ocfg, blob, 100, 250, temperature, 8)["actions"], o["actions"])   # the threshold matters
    ocfg.temperature_threshold = thr
    ctx.replay_clear()
    r = ctx.arena(900, 120, capi.OPP_RANDOM, 2, temperature)
    oa = O.arena(ocfg, blob, 900, 120, O.OPP_RANDOM, 2, temperature, 8)
    ha = ctx.history_export()
    assert _histories_equal(ha, oa, 900) == 0
    assert (r["wins"], r["draws"], r["losses"]) == tuple(int((oa["outcome"] == v).sum()) for v in (1, 0, -1))
    ctx.close()


def test_temperature_threshold_tensor_core_paths_follow_it(capi):
    """the tcgen05 FC and ResNet search kernels share the rule: with threshold 0 every ply is the argmax of its visit counts"""
    for cfg in (capi.default_config(num_slots=64, replay_buffer_size=128, num_iters=16, temperature_threshold=0, nn_mode=capi.NN_BF16_TC),
                capi.resnet_config(num_slots=56, replay_buffer_size=128, num_iters=12, temperature_threshold=0)):
        ctx = capi.Context(cfg); ctx.init_weights(9)
        ctx.self_play(0, 100, 1.0)
        h = ctx.history_export()
        order = list(cfg.child_order)[:9]
        for j in range(100):
            for t in range(int(h["T"][j])):
                cv = h["child_visits"][j, t]
                best = max(order, key=lambda a: (cv[a - 1], -order.index(a)))
                assert int(h["actions"][j, t]) == best
        ctx.close()


def test_invalid_temperature_threshold_is_rejected(capi):
    with pytest.raises(capi.MuZeroB200Error) as e:
        capi.Context(capi.default_config(num_slots=32, temperature_threshold=-2))
    assert e.value.code == capi.E_ARG


@pytest.mark.parametrize("mode", ["l2", "bptt"])
def test_checkpoint_resume_with_per_and_reanalyse_is_bit_identical(capi, mode, tmp_path):
    """the replay STATE must survive a checkpoint -- priorities as update_priorities! left them, reanalysed values +
    flags, the original keys (ring positions, eviction order) and the save_game counters."""
    gm = capi.GRAD_BPTT if mode == "bptt" else capi.GRAD_REFERENCE_L2
    kw = dict(num_slots=64, replay_buffer_size=128, batch_size=48, per=1)
    ctx = capi.Context(capi.default_config(**kw))
    ctx.init_weights(41); ctx.self_play(0, 200, 1.0)                 # 200 games into a 128-entry ring: first key 73
    ctx.learn_steps(1, 4, gm)                                        # update_priorities! has changed q_pos / q_game
    ctx.reanalyse(key0=100, n=60)                                    # some games carry reanalysed values
    q_before = ctx.replay_priorities(); info_before = ctx.replay_info(); cnt_before = ctx.replay_counters()
    ck = ctx.checkpoint()
    np.savez(tmp_path / "ck.npz", **ck)
    la = ctx.learn_steps(5, 3, gm); wa = ctx.get_weights(); qa = ctx.replay_priorities()
    ctx.self_play(200, 40, 1.0); ha = ctx.history_export(); ia = ctx.replay_info()
    ctx.close()
    ctx2 = capi.Context(capi.default_config(**kw))
    with np.load(tmp_path / "ck.npz") as z:
        ctx2.restore({k: z[k] for k in z.files})
    assert ctx2.replay_info() == info_before and info_before["first_key"] == 73
    assert np.array_equal(ctx2.replay_counters(), cnt_before)
    q2 = ctx2.replay_priorities()
    assert np.array_equal(q2[0], q_before[0]) and np.array_equal(q2[1], q_before[1])
    v2, f2 = ctx2.reanalysed_export()
    assert np.array_equal(v2, ck["reanalysed_values"]) and np.array_equal(f2, ck["reanalysed_set"]) and f2.sum() == 60
    lb = ctx2.learn_steps(5, 3, gm); wb = ctx2.get_weights(); qb = ctx2.replay_priorities()
    assert np.array_equal(wa, wb) and np.array_equal(la, lb)
    assert np.array_equal(qa[0], qb[0]) and np.array_equal(qa[1], qb[1])
    ctx2.self_play(200, 40, 1.0); hb = ctx2.history_export()
    assert ctx2.replay_info() == ia
    for k in ("game_id",) + common.HIST_KEYS:
        assert np.array_equal(ha[k], hb[k]), k
    ctx2.close()


def test_learner_restart_at_step_one_is_a_fresh_optimiser(capi):
    """learning! builds a new ADAMW per call (Learning.jl:318): t = 1 on a used context must equal t = 1 on a new one."""
    kw = dict(num_slots=32, replay_buffer_size=64, batch_size=16)
    ctx = capi.Context(capi.default_config(**kw)); ctx.init_weights(3); blob = ctx.get_weights()
    ctx.self_play(0, 40, 1.0)
    batch = ctx.get_batch(1)
    ctx.learn_step(1, capi.GRAD_BPTT, batch); ctx.learn_step(2, capi.GRAD_BPTT, batch)
    ctx.set_weights(blob)
    ctx.learn_step(1, capi.GRAD_BPTT, batch); w_again = ctx.get_weights()
    ctx.close()
    ctx2 = capi.Context(capi.default_config(**kw)); ctx2.set_weights(blob)
    ctx2.learn_step(1, capi.GRAD_BPTT, batch); w_fresh = ctx2.get_weights()
    ctx2.close()
    assert np.array_equal(w_again, w_fresh)
