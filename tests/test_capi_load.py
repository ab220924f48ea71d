"""CPU-only: the C-ABI library loads, exports every symbol include/muzero_b200.h declares, and refuses to run
without a CUDA device (no CPU fallback).  No compute calls."""
import ctypes as C
import os
import re
import unicodedata

import numpy as np
import pytest

import common
from oracle import oracle as O


def header_symbols():
    text = open(os.path.join(common.ROOT, "include", "muzero_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mz_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from muzero_jl_b200 import capi
    L = capi.lib()
    names = header_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), "libmuzero_b200.so does not export %s" % n
    assert set(names) == set(L._mz_symbols), "capi.py binding and header disagree"
    assert L.mz_abi_version() == 7


def test_config_defaults_follow_params_jl():
    from muzero_jl_b200 import capi
    c = capi.default_config()
    assert (c.W, c.H, c.C, c.A, c.num_players) == (3, 3, 3, 9, 2)
    assert (c.stacked_observations, c.max_moves, c.num_iters, c.num_unroll_steps, c.td_steps, c.batch_size) == (1, 9, 10, 5, 5, 32)
    assert (c.pb_c_base, c.replay_buffer_size, c.seed) == (19652, 10000, 1337)
    assert abs(c.discount - 0.997) < 1e-7 and c.pb_c_init == 1.25 and c.dirichlet_alpha == 0.25 and c.exploration_eps == 0.25
    assert list(c.child_order)[:9] == [7, 4, 9, 2, 3, 5, 8, 6, 1]
    assert [capi.lib().mz_num_params(C.byref(c), i) for i in range(4)] == [18331, 23242, 33308, 74881]


def test_invalid_config_is_rejected_with_message():
    from muzero_jl_b200 import capi
    c = capi.default_config(hidden_state_size=28)
    assert capi.lib().mz_num_params(C.byref(c), 3) == capi.E_ARG
    assert b"hidden_state_size" in capi.lib().mz_last_error(None)


def test_slot_count_limit_is_reported():
    from muzero_jl_b200 import capi
    c = capi.default_config(num_slots=65537)
    assert capi.lib().mz_num_params(C.byref(c), 3) == capi.E_ARG
    assert b"num_slots" in capi.lib().mz_last_error(None)
    assert capi.lib().mz_num_params(C.byref(capi.default_config(num_slots=65536)), 3) == 74881


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from muzero_jl_b200 import capi
    with pytest.raises(capi.MuZeroB200Error) as e:
        capi.Context()
    assert e.value.code == capi.E_CUDA and "no CPU fallback" in str(e.value)


def test_python_interface_mirrors_reference_names():
    import muzero_jl_b200 as mz
    conf = mz.Config()
    for f in ("seed", "observation_shape", "action_space", "players", "stacked_observations", "muzero_player", "opponent",
              "intermediate_rewards", "num_workers", "selfplay_on_gpu", "max_moves", "temperature_threshold", "dirichlet_α",
              "exploration_ϵ", "pb_c_base", "pb_c_init", "discount", "num_iters", "replay_buffer_size", "num_unroll_steps",
              "td_steps", "PER", "PER_alpha", "results_path", "networks_path", "training_steps", "batch_size",
              "checkpoint_interval", "value_loss_weight"):
        # Python NFKC-normalises identifiers (the reference's U+03F5 becomes U+03B5); same spelling at the call site
        assert hasattr(conf, unicodedata.normalize("NFKC", f)), f      # src/Constructors.jl:18-52
    hp = mz.FeedForwardHP()
    for f in ("width_hidden", "depth_representation", "depth_prediction", "depth_dynamics", "depth_policy", "depth_value",
              "depth_reward", "depth_state_head", "use_batch_norm", "batch_norm_momentum", "hidden_state_size", "reward_activation"):
        assert hasattr(hp, f), f        # src/Constructors.jl:62-75
    rhp = mz.ResNetHP()
    for f in ("num_blocks", "depth_representation", "num_filters", "conv_kernel_size", "num_second_head_filters", "num_first_head_filters",
              "batch_norm_momentum", "downsample", "hidden_state_size", "representation_output_size", "depth_policy", "depth_value"):
        assert hasattr(rhp, f), f       # src/Constructors.jl:77-90
    from muzero_jl_b200.api import to_mz_config
    c = to_mz_config(conf, hp, num_slots=64)
    from muzero_jl_b200 import capi
    rc = to_mz_config(conf, rhp, num_slots=64); want = capi.resnet_config(num_slots=64, replay_buffer_size=rc.replay_buffer_size)
    for name, _ in capi.MzConfig._fields_:
        a_, b_ = getattr(rc, name), getattr(want, name)
        assert (list(a_) == list(b_)) if hasattr(a_, "__len__") else (a_ == b_), name
    ref = O.default_config()
    o = common.oracle_config(c)
    for name, _ in O.Config._fields_:
        if name in ("replay_buffer_size", "child_order"):
            continue
        assert getattr(o, name) == getattr(ref, name), name


def test_dropin_entry_points_keep_the_reference_signatures():
    """the signatures the wrapper "must keep exactly" -- argument names and order of the reference's functions
    (SelfPlay.jl:230, 330, 384-390; ReplayBuffer.jl:133-135, 188; Learning.jl:306-309), `!` spelt `_`."""
    import inspect
    import muzero_jl_b200 as mz
    d = mz.dropin
    want = {
        "run_mcts": ["observation", "legal_actions", "to_play", "exploration", "NNs_"],
        "play_game": ["env", "temperature", "render", "opponent", "muzero_player", "NNs_"],
        "self_play_": ["env", "training_step", "num_played_games", "num_played_steps", "total_samples", "remote_NNs", "remote_buffer"],
        "save_game": ["history", "remote_buffer", "num_played_games", "num_played_steps", "total_samples"],
        "get_batch": ["buffer"],
        "learning_": ["num_played_games", "training_step", "remote_NNs", "remote_buffer"],
        "select_action": ["node", "temperature"],
        "init_representation": ["hyper_"], "init_prediction": ["hyper_"], "init_dynamics": ["hyper_"],
    }
    for name, args in want.items():
        params = [p for p in inspect.signature(getattr(d, name)).parameters.values() if p.default is inspect.Parameter.empty]
        assert [p.name for p in params] == args, name
    # the capacity-1 channel of main.jl:15-19
    ch = d.RemoteChannel(lambda: d.Channel(1))
    d.put(ch, 0); assert d.fetch(ch) == 0 and d.take(ch) == 0 and not ch.isready()
    with pytest.raises(RuntimeError):
        d._engine = None; d.run_mcts(None, [1], 1, True, None)
    # the Julia wrapper declares the same methods (source only: no Julia in this image)
    src = open(os.path.join(common.ROOT, "muzero.jl_b200", "julia", "MuZeroB200.jl")).read()
    for sig in ("function run_mcts(observation::Array{Float32,3}, legal_actions::Vector{Int}, to_play::Int, exploration::Bool, NNs)",
                "function play_game(env, temperature, render::Bool, opponent::String, muzero_player::Int, NNs)",
                "function self_play!(env, training_step, num_played_games, num_played_steps, total_samples, remote_NNs, remote_buffer)",
                "function save_game(history, remote_buffer, num_played_games, num_played_steps, total_samples)",
                "function get_batch(buffer)",
                "function learning!(num_played_games, training_step, remote_NNs, remote_buffer)"):
        assert sig in src, sig
