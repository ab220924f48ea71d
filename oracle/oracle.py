"""ctypes loader for the CPU oracle (oracle/mz_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmz_oracle.so")

MAX_A = 16
TIE_PHILOX, TIE_FIRST = 0, 1
GRAD_REFERENCE_L2, GRAD_BPTT = 0, 1


class Config(C.Structure):
    """POD mirror of mzo_config (oracle/mz_oracle.h) = Config + FeedForwardHP (src/Constructors.jl:18-75)."""
    _fields_ = [
        ("game", C.c_int32), ("W", C.c_int32), ("H", C.c_int32), ("C", C.c_int32), ("A", C.c_int32),
        ("num_players", C.c_int32), ("stacked_observations", C.c_int32), ("max_moves", C.c_int32),
        ("num_iters", C.c_int32), ("num_unroll_steps", C.c_int32), ("td_steps", C.c_int32),
        ("batch_size", C.c_int32), ("replay_buffer_size", C.c_int32), ("pb_c_base", C.c_int32),
        ("intermediate_rewards", C.c_int32), ("tie_mode", C.c_int32),
        ("pb_c_init", C.c_float), ("discount", C.c_float), ("dirichlet_alpha", C.c_float),
        ("exploration_eps", C.c_float), ("seed", C.c_uint64), ("child_order", C.c_int32 * MAX_A),
        ("width_hidden", C.c_int32), ("depth_representation", C.c_int32), ("depth_prediction", C.c_int32),
        ("depth_dynamics", C.c_int32), ("depth_policy", C.c_int32), ("depth_value", C.c_int32),
        ("depth_reward", C.c_int32), ("depth_state_head", C.c_int32), ("hidden_state_size", C.c_int32),
        ("reward_activation_tanh", C.c_int32),
        ("net_type", C.c_int32), ("rn_num_blocks", C.c_int32), ("rn_num_filters", C.c_int32), ("rn_kernel", C.c_int32),
        ("rn_first_head_filters", C.c_int32), ("rn_second_head_filters", C.c_int32),
        ("per", C.c_int32), ("per_alpha", C.c_int32), ("temperature_threshold", C.c_int32), ("use_batch_norm", C.c_int32),
    ]


class Env(C.Structure):
    _fields_ = [("p1", C.c_uint64), ("p2", C.c_uint64), ("player", C.c_int32), ("moves", C.c_int32)]


def _cpu_has(flag):
    try:
        with open("/proc/cpuinfo") as f:
            return flag in f.read()
    except OSError:
        return False


def build(force=False):
    """Compile the oracle (gcc).  Falls back to a generic build when the host lacks AVX2/FMA."""
    src = os.path.join(_HERE, "mz_oracle.c")
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "mz_oracle.h"))):
        return _SO
    target = [] if (_cpu_has("avx2") and _cpu_has("fma")) else ["generic"]
    subprocess.check_call(["make", "-C", _HERE, "-B"] + target, stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build(force=not os.path.exists(_SO) or not (_cpu_has("avx2") and _cpu_has("fma")))   # no-op when the library is newer than its sources
    L = C.CDLL(_SO)
    f32p, i32p, i64p, u32p = (C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_uint32))
    cfgp = C.POINTER(Config)
    L.mzo_default_config.argtypes = [cfgp]
    L.mzo_num_params.argtypes = [cfgp, C.c_int]; L.mzo_num_params.restype = C.c_int
    L.mzo_init_weights.argtypes = [cfgp, C.c_uint64, f32p]
    L.mzo_julia_dict_order.argtypes = [C.c_int, i32p]
    for name in ("mzo_expf", "mzo_logf", "mzo_tanhf"):
        getattr(L, name).argtypes = [C.c_float]; getattr(L, name).restype = C.c_float
    L.mzo_set_bf16.argtypes = [C.c_int]
    L.mzo_philox.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u32p]
    envp = C.POINTER(Env)
    L.mzo_env_reset.argtypes = [cfgp, envp]
    L.mzo_env_step.argtypes = [cfgp, envp, C.c_int]
    L.mzo_env_legal_mask.argtypes = [cfgp, envp]; L.mzo_env_legal_mask.restype = C.c_uint32
    L.mzo_env_is_terminated.argtypes = [cfgp, envp]; L.mzo_env_is_terminated.restype = C.c_int
    L.mzo_env_reward.argtypes = [cfgp, envp, C.c_int]; L.mzo_env_reward.restype = C.c_int
    L.mzo_env_observation.argtypes = [cfgp, envp, f32p]
    L.mzo_env_census.argtypes = [cfgp, i64p]
    L.mzo_representation.argtypes = [cfgp, f32p, f32p, f32p]
    L.mzo_prediction.argtypes = [cfgp, f32p, f32p, f32p, f32p]
    L.mzo_dynamics.argtypes = [cfgp, f32p, f32p, f32p, f32p]
    L.mzo_stack_observations.argtypes = [cfgp, f32p, i32p, C.c_int, f32p]
    L.mzo_run_mcts.argtypes = [cfgp, f32p, f32p, C.c_uint32, C.c_int, C.c_int, C.c_uint64, C.c_int, i32p, f32p, f32p, f32p]
    L.mzo_select_action.argtypes = [cfgp, i32p, C.c_uint32, C.c_float, C.c_uint64, C.c_int]; L.mzo_select_action.restype = C.c_int
    L.mzo_self_play.argtypes = [cfgp, f32p, C.c_uint64, C.c_int, C.c_float, C.c_int, i32p, f32p, i32p, f32p, i32p, f32p, f32p]
    L.mzo_self_play.restype = C.c_int64
    L.mzo_compute_target_value.argtypes = [cfgp, C.c_int, f32p, i32p, f32p, C.c_int]; L.mzo_compute_target_value.restype = C.c_float
    L.mzo_get_batch.argtypes = [cfgp, C.c_int, C.c_int64, i32p, f32p, i32p, f32p, i32p, f32p, f32p, C.c_uint64,
                                i32p, f32p, f32p, f32p, f32p, f32p, f32p]
    L.mzo_learn_forward.argtypes = [cfgp, f32p, C.c_int] + [f32p] * 10
    L.mzo_cos_schedule.argtypes = [C.c_int]; L.mzo_cos_schedule.restype = C.c_double
    L.mzo_learn_step.argtypes = [cfgp, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int] + [f32p] * 7
    L.mzo_learn_gradients.argtypes = [cfgp, f32p, C.c_int] + [f32p] * 6 + [C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double)]
    L.mzo_learn_gradients.restype = C.c_double
    L.mzo_adam_apply.argtypes = [f32p, f32p, f32p, f32p, C.c_int, C.c_int]
    L.mzo_per_quantise.argtypes = [C.c_float]; L.mzo_per_quantise.restype = C.c_uint32
    L.mzo_per_priorities.argtypes = [cfgp, C.c_int, f32p, i32p, f32p, u32p, u32p]
    L.mzo_get_batch_per.argtypes = [cfgp, C.c_int, C.c_int64, i32p, f32p, i32p, f32p, i32p, f32p, f32p, u32p, u32p, C.c_uint64,
                                    i32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p]
    L.mzo_per_update.argtypes = [cfgp, C.c_int, i32p, f32p, f32p, C.c_int, C.c_int64, i32p, u32p, u32p]
    L.mzo_learn_forward_w.argtypes = [cfgp, f32p, C.c_int] + [f32p] * 11
    L.mzo_learn_gradients_w.argtypes = [cfgp, f32p, C.c_int] + [f32p] * 7 + [C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double)]
    L.mzo_learn_gradients_w.restype = C.c_double
    _lib = L
    return L


def set_bf16(on):
    """bf16-operand emulation of the networks (for checking the tensor-core path); off = the reference's Float32."""
    lib().mzo_set_bf16(int(on))


def _p(a, t=C.c_float):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def default_config(**kw):
    cfg = Config()
    lib().mzo_default_config(C.byref(cfg))
    for k, v in kw.items():
        if k == "child_order":
            for i, x in enumerate(v):
                cfg.child_order[i] = x
        else:
            setattr(cfg, k, v)
    return cfg


def sizes(cfg):
    planes = cfg.C * (cfg.stacked_observations + 1) + cfg.stacked_observations
    return dict(obs=cfg.W * cfg.H * cfg.C, stack=cfg.W * cfg.H * planes, sa=cfg.hidden_state_size + cfg.W * cfg.H,
                Tmax=cfg.max_moves + 1, K1=cfg.num_unroll_steps + 1, A=cfg.A, hidden=cfg.hidden_state_size)


def resnet_config(**kw):
    """TicTacToe with the repaired ResNetHP networks (hidden state (W,H,num_filters))."""
    base = dict(net_type=1, rn_num_blocks=2, rn_num_filters=64, rn_kernel=3, rn_first_head_filters=1, rn_second_head_filters=2)
    base.update(kw)
    cfg = default_config(**base)
    cfg.hidden_state_size = cfg.W * cfg.H * cfg.rn_num_filters
    return cfg


def connect_config(**kw):
    """The synthetic 6x7, 7-action Connect game of BASELINE.json configs[3] with the ResNet networks."""
    base = dict(game=1, W=6, H=7, A=7, max_moves=42)
    base.update(kw)
    cfg = resnet_config(**base)
    order = (C.c_int32 * MAX_A)()
    lib().mzo_julia_dict_order(cfg.A, order)
    for i in range(MAX_A):
        cfg.child_order[i] = order[i] if i < cfg.A else 0
    return cfg


def num_params(cfg, net=3):
    return lib().mzo_num_params(C.byref(cfg), net)


def init_weights(cfg, seed=None):
    blob = np.zeros(num_params(cfg), np.float32)
    lib().mzo_init_weights(C.byref(cfg), cfg.seed if seed is None else seed, _p(blob))
    return blob


def philox(seed, stream, c0, c1=0, c2=0, c3=0):
    out = (C.c_uint32 * 4)()
    lib().mzo_philox(seed, stream, c0, c1, c2, c3, out)
    return list(out)


def representation(cfg, blob, stacked):
    h = np.zeros(cfg.hidden_state_size, np.float32)
    lib().mzo_representation(C.byref(cfg), _p(blob), _p(np.ascontiguousarray(stacked, np.float32)), _p(h))
    return h


def prediction(cfg, blob, hidden):
    v = np.zeros(1, np.float32); p = np.zeros(cfg.A, np.float32)
    lib().mzo_prediction(C.byref(cfg), _p(blob), _p(np.ascontiguousarray(hidden, np.float32)), _p(v), _p(p))
    return v[0], p


def dynamics(cfg, blob, sa):
    h = np.zeros(cfg.hidden_state_size, np.float32); r = np.zeros(1, np.float32)
    lib().mzo_dynamics(C.byref(cfg), _p(blob), _p(np.ascontiguousarray(sa, np.float32)), _p(h), _p(r))
    return h, r[0]


def run_mcts(cfg, blob, stacked, legal_mask, to_play, exploration, game_id, move_idx, trace=False):
    vc = np.zeros(cfg.A, np.int32); rv = np.zeros(1, np.float32); pri = np.zeros(cfg.A, np.float32)
    tr = np.zeros((cfg.num_iters, 5), np.float32) if trace else None
    lib().mzo_run_mcts(C.byref(cfg), _p(blob), _p(np.ascontiguousarray(stacked, np.float32)), legal_mask, to_play,
                       int(exploration), game_id, move_idx, _p(vc, C.c_int32), _p(rv), _p(pri), _p(tr))
    return (vc, rv[0], pri, tr) if trace else (vc, rv[0], pri)


def select_action(cfg, visit_counts, legal_mask, temperature, game_id, move_idx):
    vc = np.ascontiguousarray(visit_counts, np.int32)
    return lib().mzo_select_action(C.byref(cfg), _p(vc, C.c_int32), legal_mask, temperature, game_id, move_idx)


def self_play(cfg, blob, first_game, n_games, temperature=1.0, nthreads=1):
    """Returns a dict shaped like n GameHistory objects padded to Tmax, plus total simulations."""
    s = sizes(cfg)
    out = dict(T=np.zeros(n_games, np.int32), obs=np.zeros((n_games, s["Tmax"], s["obs"]), np.float32),
               actions=np.zeros((n_games, s["Tmax"]), np.int32), rewards=np.zeros((n_games, s["Tmax"]), np.float32),
               to_play=np.zeros((n_games, s["Tmax"]), np.int32),
               child_visits=np.zeros((n_games, s["Tmax"], s["A"]), np.float32),
               root_values=np.zeros((n_games, s["Tmax"]), np.float32))
    out["sims"] = lib().mzo_self_play(C.byref(cfg), _p(blob), first_game, n_games, temperature, nthreads,
                                      _p(out["T"], C.c_int32), _p(out["obs"]), _p(out["actions"], C.c_int32),
                                      _p(out["rewards"]), _p(out["to_play"], C.c_int32), _p(out["child_visits"]),
                                      _p(out["root_values"]))
    return out


OPP_SELF, OPP_RANDOM, OPP_EXPERT = 0, 1, 2


def arena(cfg, blob, first_game, n_games, opponent=OPP_RANDOM, muzero_player=1, temperature=0.0, nthreads=1):
    """competitive_play! (SelfPlay.jl:421-435) over n games: histories as in self_play plus outcome[n] (+1 / 0 / -1 for MuZero)."""
    s = sizes(cfg)
    out = dict(T=np.zeros(n_games, np.int32), obs=np.zeros((n_games, s["Tmax"], s["obs"]), np.float32),
               actions=np.zeros((n_games, s["Tmax"]), np.int32), rewards=np.zeros((n_games, s["Tmax"]), np.float32),
               to_play=np.zeros((n_games, s["Tmax"]), np.int32),
               child_visits=np.zeros((n_games, s["Tmax"], s["A"]), np.float32),
               root_values=np.zeros((n_games, s["Tmax"]), np.float32), outcome=np.zeros(n_games, np.int32))
    L = lib()
    L.mzo_arena.restype = C.c_int64
    out["sims"] = L.mzo_arena(C.byref(cfg), _p(blob), C.c_uint64(first_game), C.c_int(n_games), C.c_int(opponent), C.c_int(muzero_player),
                              C.c_float(temperature), C.c_int(nthreads), _p(out["T"], C.c_int32), _p(out["obs"]), _p(out["actions"], C.c_int32),
                              _p(out["rewards"]), _p(out["to_play"], C.c_int32), _p(out["child_visits"]), _p(out["root_values"]),
                              _p(out["outcome"], C.c_int32))
    return out


def opponent_action(cfg, p1, p2, player, opponent, game_id, move_idx):
    e = Env(); e.p1, e.p2, e.player, e.moves = p1, p2, player, bin(p1 | p2).count("1")
    return lib().mzo_opponent_action(C.byref(cfg), C.byref(e), opponent, C.c_uint64(game_id), move_idx)


def get_batch(cfg, hist, step, first_key=1):
    s = sizes(cfg); B = cfg.batch_size
    out = dict(index=np.zeros((B, 2), np.int32), obs=np.zeros((B, s["stack"]), np.float32),
               actions=np.zeros((B, s["K1"]), np.float32), values=np.zeros((B, s["K1"]), np.float32),
               rewards=np.zeros((B, s["K1"]), np.float32), policies=np.zeros((B, s["K1"], s["A"]), np.float32),
               gscale=np.zeros(B, np.float32))
    lib().mzo_get_batch(C.byref(cfg), len(hist["T"]), first_key, _p(hist["T"], C.c_int32), _p(hist["obs"]),
                        _p(hist["actions"], C.c_int32), _p(hist["rewards"]), _p(hist["to_play"], C.c_int32),
                        _p(hist["child_visits"]), _p(hist["root_values"]), step, _p(out["index"], C.c_int32),
                        _p(out["obs"]), _p(out["actions"]), _p(out["values"]), _p(out["rewards"]),
                        _p(out["policies"]), _p(out["gscale"]))
    return out


def learn_forward(cfg, blob, batch):
    B = batch["obs"].shape[0]; s = sizes(cfg)
    pv = np.zeros((B, s["K1"]), np.float32); pr = np.zeros((B, s["K1"]), np.float32)
    pp = np.zeros((B, s["K1"], s["A"]), np.float32); losses = np.zeros(3, np.float32)
    lib().mzo_learn_forward(C.byref(cfg), _p(blob), B, _p(batch["obs"]), _p(batch["actions"]), _p(batch["values"]),
                            _p(batch["rewards"]), _p(batch["policies"]), _p(batch["gscale"]), _p(pv), _p(pr), _p(pp),
                            _p(losses))
    return pv, pr, pp, losses


def learn_step(cfg, blob, adam_m, adam_v, t, batch, grad_mode=GRAD_REFERENCE_L2):
    B = batch["obs"].shape[0]; losses = np.zeros(3, np.float32)
    lib().mzo_learn_step(C.byref(cfg), _p(blob), _p(adam_m), _p(adam_v), t, grad_mode, B, _p(batch["obs"]),
                         _p(batch["actions"]), _p(batch["values"]), _p(batch["rewards"]), _p(batch["policies"]),
                         _p(batch["gscale"]), _p(losses))
    return losses


def learn_gradients(cfg, blob, batch, fwd64=False, perturb=None, want_grad=True):
    """grad_mode = BPTT: (Float64 data loss, Float64 gradient of data loss + sum(theta^2) in blob order)."""
    B = batch["obs"].shape[0]
    grad = np.zeros(blob.shape[0], np.float64) if want_grad else None
    idx, delta = perturb if perturb is not None else (-1, 0.0)
    loss = lib().mzo_learn_gradients(C.byref(cfg), _p(blob), B, _p(batch["obs"]), _p(batch["actions"]), _p(batch["values"]),
                                     _p(batch["rewards"]), _p(batch["policies"]), _p(batch["gscale"]), int(fwd64), int(idx),
                                     float(delta), grad.ctypes.data_as(C.POINTER(C.c_double)) if want_grad else None)
    return loss, grad


def adam_apply(blob, adam_m, adam_v, grad, t):
    lib().mzo_adam_apply(_p(blob), _p(adam_m), _p(adam_v), _p(grad), blob.shape[0], t)


# ---- prioritised replay (conf.PER = true; repaired specification, see mz_oracle.c) ----
def per_priorities(cfg, hist):
    """save_game's initial priorities (ReplayBuffer.jl:136-145) for every history: (q_pos [n][Tmax], q_game [n]) fixed point."""
    n = len(hist["T"]); Tmax = cfg.max_moves + 1
    q_pos = np.zeros((n, Tmax), np.uint32); q_game = np.zeros(n, np.uint32)
    for g in range(n):
        lib().mzo_per_priorities(C.byref(cfg), int(hist["T"][g]), _p(np.ascontiguousarray(hist["rewards"][g])),
                                 _p(np.ascontiguousarray(hist["to_play"][g], np.int32), C.c_int32), _p(np.ascontiguousarray(hist["root_values"][g])),
                                 q_pos[g].ctypes.data_as(C.POINTER(C.c_uint32)), q_game[g:].ctypes.data_as(C.POINTER(C.c_uint32)))
    return q_pos, q_game


def get_batch_per(cfg, hist, q_pos, q_game, step, first_key=1):
    s = sizes(cfg); B = cfg.batch_size
    out = dict(index=np.zeros((B, 2), np.int32), obs=np.zeros((B, s["stack"]), np.float32),
               actions=np.zeros((B, s["K1"]), np.float32), values=np.zeros((B, s["K1"]), np.float32),
               rewards=np.zeros((B, s["K1"]), np.float32), policies=np.zeros((B, s["K1"], s["A"]), np.float32),
               gscale=np.zeros(B, np.float32), weights=np.zeros(B, np.float32))
    lib().mzo_get_batch_per(C.byref(cfg), len(hist["T"]), first_key, _p(hist["T"], C.c_int32), _p(hist["obs"]),
                            _p(hist["actions"], C.c_int32), _p(hist["rewards"]), _p(hist["to_play"], C.c_int32),
                            _p(hist["child_visits"]), _p(hist["root_values"]), _p(q_pos, C.c_uint32), _p(q_game, C.c_uint32), step,
                            _p(out["index"], C.c_int32), _p(out["obs"]), _p(out["actions"]), _p(out["values"]), _p(out["rewards"]),
                            _p(out["policies"]), _p(out["gscale"]), _p(out["weights"]))
    return out


def per_update(cfg, hist, q_pos, q_game, index_batch, pred_values, target_values, first_key=1):
    """update_priorities! (ReplayBuffer.jl:168-183, repaired bounds), in place."""
    lib().mzo_per_update(C.byref(cfg), index_batch.shape[0], _p(np.ascontiguousarray(index_batch, np.int32), C.c_int32),
                         _p(np.ascontiguousarray(pred_values, np.float32)), _p(np.ascontiguousarray(target_values, np.float32)),
                         len(hist["T"]), first_key, _p(hist["T"], C.c_int32), _p(q_pos, C.c_uint32), _p(q_game, C.c_uint32))


def learn_forward_w(cfg, blob, batch):
    B = batch["obs"].shape[0]; s = sizes(cfg)
    pv = np.zeros((B, s["K1"]), np.float32); pr = np.zeros((B, s["K1"]), np.float32)
    pp = np.zeros((B, s["K1"], s["A"]), np.float32); losses = np.zeros(3, np.float32)
    lib().mzo_learn_forward_w(C.byref(cfg), _p(blob), B, _p(batch["obs"]), _p(batch["actions"]), _p(batch["values"]),
                              _p(batch["rewards"]), _p(batch["policies"]), _p(batch["gscale"]), _p(batch["weights"]), _p(pv), _p(pr), _p(pp),
                              _p(losses))
    return pv, pr, pp, losses


def learn_gradients_w(cfg, blob, batch, fwd64=False, perturb=None, want_grad=True):
    B = batch["obs"].shape[0]
    grad = np.zeros(blob.shape[0], np.float64) if want_grad else None
    idx, delta = perturb if perturb is not None else (-1, 0.0)
    loss = lib().mzo_learn_gradients_w(C.byref(cfg), _p(blob), B, _p(batch["obs"]), _p(batch["actions"]), _p(batch["values"]),
                                       _p(batch["rewards"]), _p(batch["policies"]), _p(batch["gscale"]), _p(batch["weights"]), int(fwd64),
                                       int(idx), float(delta), grad.ctypes.data_as(C.POINTER(C.c_double)) if want_grad else None)
    return loss, grad


def trainable_mask(cfg):
    """1 per blob entry that is a Flux parameter, 0 for the BatchNorm running statistics of the ResNet networks."""
    n = num_params(cfg, 3)
    m = np.zeros(n, np.uint8)
    lib().mzo_trainable_mask(C.byref(cfg), m.ctypes.data_as(C.POINTER(C.c_uint8)))
    return m
