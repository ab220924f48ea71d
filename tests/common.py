"""Shared helpers for the test-suite: oracle <-> product config mapping and the CPU harness loader."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_COMMON = ["game", "W", "H", "C", "A", "num_players", "stacked_observations", "max_moves", "num_iters", "num_unroll_steps", "td_steps",
           "batch_size", "replay_buffer_size", "pb_c_base", "intermediate_rewards", "tie_mode", "pb_c_init", "discount",
           "dirichlet_alpha", "exploration_eps", "seed", "width_hidden", "depth_representation", "depth_prediction",
           "depth_dynamics", "depth_policy", "depth_value", "depth_reward", "depth_state_head", "hidden_state_size",
           "reward_activation_tanh", "net_type", "rn_num_blocks", "rn_num_filters", "rn_kernel", "rn_first_head_filters",
           "rn_second_head_filters", "per", "per_alpha", "temperature_threshold", "use_batch_norm"]


def oracle_config(mzcfg):
    """mz_config -> mzo_config (same field names by construction)."""
    o = O.Config()
    for k in _COMMON:
        setattr(o, k, getattr(mzcfg, k))
    for i in range(16):
        o.child_order[i] = mzcfg.child_order[i]
    return o


def product_config(**kw):
    from muzero_jl_b200 import capi
    return capi.default_config(**kw)


_hh = None


def harness():
    """g++ build of tests/host_harness.cpp (product scalar device code replayed on the CPU)."""
    global _hh
    if _hh is not None:
        return _hh
    from muzero_jl_b200 import capi
    src = os.path.join(ROOT, "tests", "host_harness.cpp")
    so = os.path.join(ROOT, "tests", "libhost_harness.so")
    deps = [src] + [os.path.join(ROOT, "muzero.jl_b200", "csrc", f) for f in ("mz_common.h", "mz_host.h", "mz_rn_host.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        flags = ["-mavx2", "-mfma"] if (O._cpu_has("avx2") and O._cpu_has("fma")) else []
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-math-errno", "-fPIC", "-shared"] + flags + ["-o", so, src, "-lm"])
    L = C.CDLL(so)
    f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    cfgp = C.POINTER(capi.MzConfig)
    L.hh_num_params.argtypes = [cfgp]
    L.hh_init_weights.argtypes = [cfgp, C.c_uint64, f32p]
    L.hh_nn.argtypes = [cfgp, f32p, C.c_int, f32p, f32p, f32p]
    L.hh_run_mcts.argtypes = [cfgp, f32p, f32p, C.c_uint32, C.c_int, C.c_int, C.c_uint64, C.c_int, i32p, f32p, f32p]
    L.hh_self_play.argtypes = [cfgp, f32p, C.c_uint64, C.c_int, C.c_float, i32p, f32p, i32p, f32p, i32p, f32p, f32p]
    L.hh_self_play.restype = C.c_int64
    L.hh_get_batch.argtypes = [cfgp, C.c_int, C.c_int64, i32p, f32p, i32p, f32p, i32p, f32p, f32p, C.c_uint64, i32p] + [f32p] * 6
    L.hh_rn_num_params.argtypes = [cfgp]
    L.hh_rn_program_info.argtypes = [cfgp, i32p]
    L.hh_opponent_action.argtypes = [cfgp, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_int]
    L.hh_arena_outcome.argtypes = [cfgp, C.c_int, i32p, C.c_int]
    L.hh_rn_program_dump.argtypes = [cfgp]
    L.hh_rn_forward.argtypes = [cfgp, f32p, C.c_int, C.c_int, f32p, f32p, f32p]
    _hh = L
    return L


def product_board(pcfg, actions):
    """(p1, p2, player) of the product's board encoding after `actions` from the empty board (host harness)."""
    hh = harness()
    a = np.asarray(actions, np.int32); p1, p2, pl = C.c_uint64(), C.c_uint64(), C.c_int32()
    assert hh.hh_board_after(C.byref(pcfg), len(a), a.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(p1), C.byref(p2), C.byref(pl)) == 0
    return p1.value, p2.value, pl.value


def random_stacked(cfg, n, seed=0):
    """Random but well-formed stacked observations: a random reachable TicTacToe position + its predecessor."""
    rng = np.random.default_rng(seed)
    s = O.sizes(cfg)
    L = O.lib()
    out = np.zeros((n, s["stack"]), np.float32); legal = np.zeros(n, np.uint32); to_play = np.zeros(n, np.int32)
    for i in range(n):
        while True:
            e = O.Env(); L.mzo_env_reset(C.byref(cfg), C.byref(e))
            obs = [np.zeros(s["obs"], np.float32)]; acts = []
            L.mzo_env_observation(C.byref(cfg), C.byref(e), O._p(obs[0]))
            depth = int(rng.integers(0, 6))
            ok = True
            for _ in range(depth):
                m = L.mzo_env_legal_mask(C.byref(cfg), C.byref(e))
                la = [a for a in range(1, cfg.A + 1) if m >> (a - 1) & 1]
                if not la or L.mzo_env_is_terminated(C.byref(cfg), C.byref(e)):
                    ok = False; break
                a = int(rng.choice(la)); acts.append(a)
                L.mzo_env_step(C.byref(cfg), C.byref(e), a)
                o = np.zeros(s["obs"], np.float32); L.mzo_env_observation(C.byref(cfg), C.byref(e), O._p(o)); obs.append(o)
            m = L.mzo_env_legal_mask(C.byref(cfg), C.byref(e))
            if ok and m and not L.mzo_env_is_terminated(C.byref(cfg), C.byref(e)):
                break
        hist = np.stack(obs); a_arr = np.array(acts + [0], np.int32)
        L.mzo_stack_observations(C.byref(cfg), O._p(hist), O._p(a_arr, C.c_int32), len(obs), O._p(out[i]))
        legal[i] = m; to_play[i] = e.player
    return out, legal, to_play


HIST_KEYS = ("T", "obs", "actions", "rewards", "to_play", "child_visits", "root_values")
BATCH_KEYS = ("index", "obs", "actions", "values", "rewards", "policies", "gscale")
