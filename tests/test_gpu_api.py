"""GPU tests of the reference-facing Python mirror (muzero.jl_b200/api.py): the call sequence a user of the reference
would write -- Config / FeedForwardHP -> init networks -> run_mcts / play_game / self_play! -> get_batch -> learning! --
checked against the oracle.  Also error behaviour of the C ABI."""
import numpy as np
import pytest

import common
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mz():
    import muzero_jl_b200 as mz
    return mz


def test_reference_call_sequence(mz):
    conf = mz.Config(num_iters=10, exploration_ϵ=0.0)
    hyper = mz.FeedForwardHP()
    eng = mz.Engine(conf, hyper, num_slots=64)
    NNs = mz.init_networks(eng, seed=1337)
    ocfg = common.oracle_config(eng.ctx.cfg)
    blob = eng.ctx.get_weights()
    # environment verbs (games/tictactoe/game.jl) and KAT-env-1
    env = mz.TicTacToe(eng)
    obs = env.reset()
    assert obs.shape == (3, 3, 3) and obs[2].all() and not obs[:2].any() and env.current_player() == 1
    for a in (1, 2, 4, 5, 7):
        env(a)
    assert env.legal_action_space() == [3, 6, 8, 9] and not env.is_terminated()
    env(3)
    assert env.is_terminated() and env.legal_action_space() == []
    # the same through the names of games/AbstractGame.jl (:20-123): KAT-env-2, the mislabelled winner (Q15)
    g = mz.TicTacToe(eng)
    assert g.reset_game().shape == (3, 3, 3) and g.to_play() == 1 and g.legal_actions() == list(range(1, 10))
    for a_ in (1, 2, 3, 5, 4, 8):
        ob, rew, done = g.execute_step(a_)
        assert rew == 0.0 and not done
    assert g.expert_agent() in g.legal_actions()
    ob, rew, done = g.execute_step(6)
    assert done and rew == 1.0 and ob.shape == (3, 3, 3) and g.action_to_string(6) == "cell (3, 2)" and g.close_game() is None
    g.reset_game(); g.execute_step(1); g.execute_step(4); g.execute_step(2); g.execute_step(5)
    assert g.expert_agent() == 3                              # player 1 completes 1, 2, 3
    # NNs callables: shapes of the Flux chains
    st = np.zeros((1, 63), np.float32); st[0, 18:27] = 1
    h = NNs["representation"](st); v, p = NNs["prediction"](h)
    assert h.shape == (1, 27) and p.shape == (1, 9) and abs(p.sum() - 1) < 1e-6
    # run_mcts + select_action
    root = mz.run_mcts(eng, st[0], list(range(1, 10)), 1, True, game_id=5, move_idx=1)
    ovc, orv, opri = O.run_mcts(ocfg, blob, st[0], 0x1ff, 1, True, 5, 1)
    assert root.visit_counts.tolist() == ovc.tolist() and root.value == orv and np.array_equal(root.priors, opri)
    assert mz.select_action(eng, root, 0.0, 5, 1) == O.select_action(ocfg, ovc, 0x1ff, 0.0, 5, 1)
    with pytest.raises(AssertionError):
        mz.run_mcts(eng, st[0], [], 1)                        # SelfPlay.jl:243
    # play_game / self_play! / save_game / get_batch / learning!
    hist = mz.play_game(eng, 1.0, False, "self", 1)
    o = O.self_play(ocfg, blob, 0, 1, 1.0, 1)
    T = int(o["T"][0])
    assert hist.action_history.tolist() == o["actions"][0, :T].tolist() and np.array_equal(hist.child_visits, o["child_visits"][0, :T])
    assert hist.observation_history.shape == (T, 3, 3, 3)
    sims, moves = mz.self_play(eng, 40, temperature=1.0)
    assert sims == moves * 10 and len(mz.ReplayBuffer(eng)) == 41
    mz.save_game(eng, hist, game_id=999)                       # a host-built history goes through the same ring
    assert len(mz.ReplayBuffer(eng)) == 42 and mz.ReplayBuffer(eng)[42].action_history.tolist() == hist.action_history.tolist()
    index_batch, (obs_b, act_b, val_b, rew_b, pol_b, w_b, gs_b) = mz.get_batch(eng)
    assert len(index_batch) == 32 and obs_b.shape == (32, 63) and pol_b.shape == (32, 6, 9) and w_b is None
    losses = mz.learning(eng, 3)
    assert eng.training_step == 3 and np.all(np.isfinite(losses))
    eng.close()


def test_error_behaviour(mz):
    capi = mz.capi
    ctx = capi.Context(capi.default_config(num_slots=32, replay_buffer_size=64))
    with pytest.raises(capi.MuZeroB200Error) as e:
        ctx.learn_step(1)                                      # empty buffer (Learning.jl:311 waits; the ABI reports)
    assert e.value.code == capi.E_STATE
    with pytest.raises(capi.MuZeroB200Error) as e:
        ctx.set_weights(np.zeros(10, np.float32))
    assert e.value.code == capi.E_ARG
    with pytest.raises(capi.MuZeroB200Error):
        ctx.env_step(*ctx.env_reset(2), [0, 10])               # actions out of range
    with pytest.raises(capi.MuZeroB200Error) as e:
        ctx.learn_step(1, grad_mode=capi.GRAD_BPTT, batch=None)
    with pytest.raises(capi.MuZeroB200Error):
        capi.Context(capi.default_config(num_slots=64, replay_buffer_size=32))   # ring smaller than the slot count
    ctx.close()


def test_two_contexts_are_independent(mz):
    capi = mz.capi
    a = capi.Context(capi.default_config(num_slots=32, replay_buffer_size=64, num_iters=8))
    b = capi.Context(capi.default_config(num_slots=64, replay_buffer_size=64, num_iters=20, exploration_eps=0.0))
    a.init_weights(1); b.init_weights(2)
    a.self_play(0, 40, 1.0); b.self_play(0, 40, 1.0)
    ha, hb = a.history_export(), b.history_export()
    oa = O.self_play(common.oracle_config(a.cfg), a.get_weights(), 0, 40, 1.0, 2)
    ob = O.self_play(common.oracle_config(b.cfg), b.get_weights(), 0, 40, 1.0, 2)
    for h, o in ((ha, oa), (hb, ob)):
        order = np.argsort(h["game_id"])
        assert np.array_equal(h["actions"][order], o["actions"]) and np.array_equal(h["root_values"][order], o["root_values"])
    assert a.launch_count() > 0 and a.kernel_time(0)[1] == 0   # timers are off by default
    a.close(); b.close()


def test_contexts_with_different_sizes_coexist(mz):
    capi = mz.capi
    """The shared-memory limit of a kernel is a process-wide attribute: creating a context with a smaller tree / network
    must not break an existing larger one (regression: the second mz_create used to lower the limit)."""
    import numpy as np
    big = capi.Context(capi.default_config(num_slots=64, replay_buffer_size=64, num_iters=50))
    small = capi.Context(capi.default_config(num_slots=64, replay_buffer_size=64, num_iters=5))
    big.init_weights(1); small.init_weights(1)
    s1, _ = big.self_play(0, 64, 1.0)
    s2, _ = small.self_play(0, 64, 1.0)
    assert s1 > 0 and s2 > 0
    l = big.learn_steps(1, 2, capi.GRAD_BPTT)
    assert np.all(np.isfinite(l))
    big.close(); small.close()


def test_new_entry_points_error_behaviour(mz):
    """ABI error classes of the ResNet / PER / reanalyse / checkpoint entry points: bad configurations are refused at mz_create
    with a message, misuse returns MZ_E_* and never aborts."""
    import numpy as np
    capi = mz.capi
    for bad, frag in ((capi.resnet_config(nn_mode=capi.NN_FP32_EXACT), "tensor cores"), (capi.resnet_config(rn_num_filters=20), "rn_num_filters"),
                      (capi.resnet_config(rn_kernel=5), "rn_kernel"), (capi.default_config(per=1, per_alpha=7), "PER_alpha"),
                      (capi.resnet_config(per=1), "PER"), (capi.default_config(net_type=9), "net_type")):
        with pytest.raises(capi.MuZeroB200Error) as e:
            capi.Context(bad)
        assert frag in str(e.value), str(e.value)
    ctx = capi.Context(capi.default_config(num_slots=32, replay_buffer_size=64))
    ctx.init_weights(1)
    with pytest.raises(capi.MuZeroB200Error):
        ctx.reanalyse(key0=1, n=4)                    # empty buffer
    ctx.self_play(0, 40, 1.0)
    with pytest.raises(capi.MuZeroB200Error):
        ctx.reanalyse(key0=30, n=40)                  # keys 30..69 but only 1..40 exist
    with pytest.raises(capi.MuZeroB200Error):
        ctx.get_batch_per(1)                          # conf.PER is false
    ck = ctx.checkpoint()
    ck["adam_m"] = ck["adam_m"][:-1]
    with pytest.raises(capi.MuZeroB200Error):
        ctx.restore(ck)
    vc, rv = ctx.run_mcts(np.zeros((0, 63), np.float32), np.zeros(0, np.uint32), np.zeros(0, np.int32), False, np.zeros(0, np.uint64), np.zeros(0, np.int32))
    assert vc.shape == (0, 9)
    ctx.close()
    rn = capi.Context(capi.resnet_config(num_slots=8, replay_buffer_size=8, num_iters=4))
    rn.init_weights(1)
    assert rn.representation(np.zeros((0, 63), np.float32)).shape == (0, 576)
    with pytest.raises(capi.MuZeroB200Error):
        rn.reanalyse(key0=1, n=1)
    ck = rn.checkpoint()                              # the ResNet learner keeps an optimiser state too (reference_l2 semantics)
    assert int(ck["steps_done"]) == 0 and not ck["adam_m"].any()
    assert rn.learner_path(capi.GRAD_REFERENCE_L2) == 3 and rn.learner_path(capi.GRAD_BPTT) == -1
    rn.close()


def test_dropin_main_jl_call_sequence(mz):
    """games/tictactoe/main.jl:14-41 line for line through the reference-signature layer (muzero.jl_b200/dropin.py)."""
    d = mz.dropin
    conf = mz.Config(num_iters=8, training_steps=40, checkpoint_interval=10, replay_buffer_size=512)
    hyper = mz.FeedForwardHP()
    eng = d.bind(mz.Engine(conf, hyper, num_slots=64))
    env = d.TicTacToe()                                                       # main.jl:14
    training_step = d.RemoteChannel(lambda: d.Channel(1))                     # :15-19
    num_played_games = d.RemoteChannel(lambda: d.Channel(1))
    num_played_steps = d.RemoteChannel(lambda: d.Channel(1))
    num_reanalysed_games = d.RemoteChannel(lambda: d.Channel(1))
    total_samples = d.RemoteChannel(lambda: d.Channel(1))
    remote_NNs = d.RemoteChannel(lambda: d.Channel(1))                        # :20
    remote_buffer = d.RemoteChannel(lambda: d.BufferChannel())                # :21
    d.put(remote_NNs, d.NNs(representation=d.init_representation(d.hyper), prediction=d.init_prediction(d.hyper), dynamics=d.init_dynamics(d.hyper)))   # :23
    for ch in (training_step, num_played_games, num_played_steps, num_reanalysed_games, total_samples):
        d.put(ch, 0)                                                          # :24-28
    w0 = eng.ctx.get_weights()
    # the pieces, with the reference's own argument lists
    nns = d.fetch(remote_NNs)
    obs = d.reset_(env)
    stacked = np.concatenate([obs.reshape(-1), np.zeros(36, np.float32)])
    root = d.run_mcts(stacked.reshape(7, 3, 3), d.legal_action_space(env, d.current_player(env)), d.current_player(env), True, nns)   # SelfPlay.jl:359
    assert root.visit_counts.sum() == 8 and 1 <= d.select_action(root, 1.0) <= 9
    hist = d.play_game(env, 1.0, False, "self", conf.muzero_player, nns)       # SelfPlay.jl:405
    assert 6 <= len(hist.action_history) <= 9 and len(remote_buffer) == 0      # returned, not saved
    d.save_game(hist, remote_buffer, num_played_games, num_played_steps, total_samples)   # :414
    assert len(remote_buffer) == 1 and d.fetch(num_played_games) == 1 and d.fetch(total_samples) == len(hist.action_history)
    assert remote_buffer[1].action_history.tolist() == hist.action_history.tolist()
    index_batch, batch = d.get_batch(remote_buffer)                            # Learning.jl:331
    assert len(index_batch) == 32 and len(batch) == 7 and batch[5] is None
    hist_r = d.play_game(env, 0.0, False, "random", 2, nns)
    assert set(hist_r.to_play_history.tolist()) == {1, 2}
    # the two actors (main.jl:30-41)
    sp = d.spawnat(d.self_play_, env, training_step, num_played_games, num_played_steps, total_samples, remote_NNs, remote_buffer)
    learn = d.spawnat(d.learning_, num_played_games, training_step, remote_NNs, remote_buffer)
    learn.join(timeout=120); sp.join(timeout=120)
    assert not learn.is_alive() and not sp.is_alive() and learn.error is None and sp.error is None
    assert learn.result is True and sp.result is True
    assert d.fetch(training_step) == 41 and eng.training_step == 41            # while training_step_ <= conf.training_steps
    n = d.fetch(num_played_games)
    assert n >= 65 and (n - 1) % 64 == 0 and len(remote_buffer) == min(n, 512)
    assert d.fetch(total_samples) == eng.ctx.replay_info()["total_samples"] and d.fetch(num_played_steps) >= d.fetch(total_samples)
    assert not np.array_equal(eng.ctx.get_weights(), w0) and np.all(np.isfinite(list(eng.last_losses.values())))
    eng.close()
