"""The reference's entry points with the reference's own signatures, so that the driver script
games/tictactoe/main.jl:14-41 reads the same line for line:

    env = TicTacToe()
    training_step = RemoteChannel(lambda: Channel(1)); ...                               # main.jl:15-19
    remote_NNs = RemoteChannel(lambda: Channel(1)); remote_buffer = RemoteChannel(BufferChannel)   # :20-21
    put(remote_NNs, NNs(representation=init_representation(hyper), prediction=init_prediction(hyper), dynamics=init_dynamics(hyper)))
    put(training_step, 0); ...                                                           # :24-28
    sp = spawnat(self_play_, env, training_step, num_played_games, num_played_steps, total_samples, remote_NNs, remote_buffer)   # :30-36
    learn = spawnat(learning_, num_played_games, training_step, remote_NNs, remote_buffer)                                       # :38-41

Like the reference, the functions read module globals: `conf`, `hyper` (games/tictactoe/params.jl) and -- instead of the Flux models
living in `remote_NNs` -- one `Engine` (GPU context) bound with `bind(engine)`.  Julia's `!` suffix is spelt `_` (self_play!, learning!
-> self_play_, learning_).  Everything computes on the GPU through the C ABI; the channels carry counters only ("keep the
RemoteChannel arguments for signature compatibility but use them only for counters").

Differences that are deliberate and stated:
  * self_play! plays one WAVE of `engine.ctx.cfg.num_slots` concurrent games per iteration instead of one game per training step: the
    reference's lock-step take!(training_step) (SelfPlay.jl:396 / Learning.jl:411) is the throughput ceiling this path removes;
  * run_mcts / play_game draw their random numbers from the library's counter-based streams, keyed by a per-engine
    call counter, where the reference uses an unseeded global RNG.
"""
import threading
from collections import namedtuple
from typing import List

import numpy as np

from . import api, capi

NNs = namedtuple("NNs", ["representation", "prediction", "dynamics"])     # NamedTuple{(:representation, :prediction, :dynamics)}

conf: api.Config = None
hyper = None
_engine: api.Engine = None
_lock = threading.RLock()          # a context is used by one host thread at a time (include/muzero_b200.h)


def bind(engine: api.Engine):
    """The module-global engine (as the reference binds the globals conf / hyper, games/tictactoe/params.jl:2,18)."""
    global conf, hyper, _engine
    _engine, conf, hyper = engine, engine.conf, engine.hyper
    engine._mcts_calls = 0
    return engine


def _e() -> api.Engine:
    if _engine is None:
        raise RuntimeError("no engine bound: dropin.bind(Engine(conf, hyper)) first")
    return _engine


# ---- Distributed stand-ins (main.jl:15-21): a capacity-1 channel and the replay buffer channel -------------------------------------
class Channel:
    """Channel{T}(1): put! blocks while full, take! blocks while empty, fetch reads without removing."""

    def __init__(self, capacity=1):
        self._cv = threading.Condition(); self._items: List = []; self.capacity = capacity

    def put(self, v):
        with self._cv:
            while len(self._items) >= self.capacity:
                self._cv.wait()
            self._items.append(v); self._cv.notify_all()
        return v

    def take(self):
        with self._cv:
            while not self._items:
                self._cv.wait()
            v = self._items.pop(0); self._cv.notify_all()
        return v

    def fetch(self):
        with self._cv:
            while not self._items:
                self._cv.wait()
            return self._items[0]

    def isready(self):
        with self._cv:
            return bool(self._items)


def RemoteChannel(f):                      # RemoteChannel(()->Channel{Int}(1)): the channel itself, there is one process
    return f()


def put(ch, v): return ch.put(v)           # put!
def take(ch): return ch.take()             # take!
def fetch(ch): return ch.fetch()


class BufferChannel(api.ReplayBuffer):
    """src/RemoteBufferChannel.jl: Dict{Int,GameHistory} keyed by game number -- here a view of the engine's device replay ring."""

    def __init__(self):
        super().__init__(_e())


def spawnat(f, *args):
    """@spawnat :any f(args...) (main.jl:30,38): a host thread; fetch with .join()/.result."""
    class _Future(threading.Thread):
        def run(self):
            try:
                self.result = f(*args); self.error = None
            except BaseException as ex:     # the reference's Futures swallow worker exceptions; here they are kept
                self.result = None; self.error = ex
    t = _Future(daemon=True); t.start()
    return t


def _bump(ch: Channel, delta):             # Constructors.jl:55-59: take! then put! on a capacity-1 channel
    ch.put(ch.take() + delta)


# ---- networks (Learning.jl:87,100,118) -----------------------------------------------------------------------------------------------
def _callable(net):
    def call(x):
        with _lock:
            return getattr(_e().ctx, net)(x)
    call.net = net
    return call


def init_representation(hyper_):
    """init_representation(hyper): Glorot-uniform weights on the device (all three networks are initialised together, once per engine)."""
    with _lock:
        e = _e()
        if not getattr(e, "_weights_ready", False):
            e.ctx.init_weights(e.conf.seed); e._weights_ready = True
    return _callable("representation")


def init_prediction(hyper_):
    init_representation(hyper_)
    return _callable("prediction")


def init_dynamics(hyper_):
    init_representation(hyper_)
    return _callable("dynamics")


class TicTacToe(api.TicTacToe):
    """games/tictactoe/game.jl: `TicTacToe()` with the RLBase verbs as free functions below."""

    def __init__(self):
        super().__init__(_e())


def reset_(env): return env.reset()                                  # reset!(env)
def current_player(env): return env.current_player()
def legal_action_space(env, p=None): return env.legal_action_space(p)
def is_terminated(env): return env.is_terminated()
def reward(env, p):                                                  # RLBase.reward(env, p), game.jl:87-100 (+1 for p == 1 once the side to move has a line)
    if not hasattr(env, "_last_reward"):
        return 0.0
    r = float(env._last_reward[0])                                   # the kernel's reward(env, mover)
    return r if env._last_mover == p else -r


# ---- SelfPlay.jl ------------------------------------------------------------------------------------------------------------------------
def run_mcts(observation, legal_actions, to_play, exploration, NNs_) -> api.Node:
    """run_mcts(observation::Array{Float32,3}, legal_actions::Vector{Int}, to_play::Int, exploration::Bool, NNs)::Node (SelfPlay.jl:230)."""
    with _lock:
        e = _e(); e._mcts_calls += 1
        return api.run_mcts(e, observation, legal_actions, to_play, exploration, game_id=(1 << 31) + e._mcts_calls, move_idx=1)


def select_action(node: api.Node, temperature) -> int:
    """select_action(node, temperature) (SelfPlay.jl:293-306)."""
    with _lock:
        e = _e()
        return api.select_action(e, node, float(temperature), game_id=(1 << 31) + e._mcts_calls, move_idx=1)


visit_softmax_temperature_fn = api.visit_softmax_temperature_fn           # SelfPlay.jl:48-56


def _histories(h, n):
    W, H, Cc = conf.observation_shape
    out = []
    for j in range(n):
        T = int(h["T"][j])
        out.append(api.GameHistory(h["obs"][j, :T].reshape(T, Cc, H, W).copy(), h["actions"][j, :T].copy(), h["rewards"][j, :T].copy(),
                                   h["to_play"][j, :T].copy(), h["child_visits"][j, :T].copy(), h["root_values"][j, :T].copy()))
    return out


def play_game(env, temperature, render: bool, opponent: str, muzero_player: int, NNs_) -> api.GameHistory:
    """play_game(env, temperature, render, opponent, muzero_player, NNs)::GameHistory (SelfPlay.jl:330-382).  The history is returned, not
    saved (the caller passes it to save_game, SelfPlay.jl:414)."""
    if opponent == "human":
        raise NotImplementedError("a human opponent has no batched meaning; use the reference's own loop (SelfPlay.jl:312-316)")
    if opponent not in api._OPPONENTS:
        raise ValueError("Wrong argument: opponent argument should be self, human, expert or random")     # SelfPlay.jl:323
    with _lock:
        e = _e()
        h = e.ctx.play_games(e.next_game, 1, float(temperature), api._OPPONENTS[opponent], muzero_player)
        e.next_game += 1
    return _histories(h, 1)[0]


def save_game(history: api.GameHistory, remote_buffer, num_played_games, num_played_steps, total_samples):
    """save_game(history, remote_buffer, num_played_games, num_played_steps, total_samples) (ReplayBuffer.jl:133-161)."""
    with _lock:
        e = _e()
        before = e.ctx.replay_counters()
        api.save_game(e, history, game_id=int(before[0]))
        after = e.ctx.replay_counters()
    _bump(num_played_games, int(after[0] - before[0])); _bump(num_played_steps, int(after[1] - before[1]))   # :149-151
    total_samples.take(); total_samples.put(int(after[2]))                                                  # :152,158-160 (evictions included)


def self_play_(env, training_step, num_played_games, num_played_steps, total_samples, remote_NNs, remote_buffer) -> bool:
    """self_play!(env, training_step, num_played_games, num_played_steps, total_samples, remote_NNs, remote_buffer)::Bool (SelfPlay.jl:384-419),
    one wave of concurrent games per iteration."""
    fetch(remote_NNs)                                                       # :392 (the weights live on the device)
    training_step_ = 0
    while training_step_ <= conf.training_steps:                            # :394
        training_step_ = fetch(training_step)                               # :396 without the lock-step take!
        temperature = visit_softmax_temperature_fn(training_step_)          # :397
        with _lock:
            e = _e()
            before = e.ctx.replay_counters()
            e.ctx.self_play(e.next_game, e.ctx.cfg.num_slots, temperature)  # play_game + save_game for a wave (:405-417)
            e.next_game += e.ctx.cfg.num_slots
            after = e.ctx.replay_counters()
        _bump(num_played_games, int(after[0] - before[0])); _bump(num_played_steps, int(after[1] - before[1]))
        total_samples.take(); total_samples.put(int(after[2]))
    return True


def competitive_play_(NNs_=None):
    """competitive_play!(; NNs) (SelfPlay.jl:421-435): one game against conf.opponent at temperature 0; returns its GameHistory."""
    opponent = conf.opponent if len(conf.players) > 1 else "self"           # :428
    return play_game(None, 0.0, True, opponent, conf.muzero_player, NNs_)


# ---- ReplayBuffer.jl -------------------------------------------------------------------------------------------------------------------
def get_batch(buffer):
    """get_batch(buffer::Dict{Int,GameHistory}) (ReplayBuffer.jl:188-217) on the device ring `buffer` views."""
    with _lock:
        return api.get_batch(buffer.engine if hasattr(buffer, "engine") else _e())


# ---- Learning.jl -----------------------------------------------------------------------------------------------------------------------
def learning_(num_played_games, training_step, remote_NNs, remote_buffer, grad_mode=capi.GRAD_REFERENCE_L2) -> bool:
    """learning!(num_played_games, training_step, remote_NNs, remote_buffer)::Bool (Learning.jl:306-438)."""
    while fetch(num_played_games) < 1:                                      # :311-314
        threading.Event().wait(0.001)
    e = _e()
    training_step_ = fetch(training_step)
    while training_step_ <= conf.training_steps:                            # :327
        n = min(conf.checkpoint_interval, conf.training_steps - training_step_ + 1)
        with _lock:
            losses = e.ctx.learn_steps(training_step_ + 1, n, grad_mode)    # :329-404, n iterations queued back to back
            e.training_step = training_step_ + n
        training_step_ += n
        training_step.take(); training_step.put(training_step_)             # :411
        if remote_NNs.isready():                                            # :418-419 publish (the self-play side reads the same device weights)
            remote_NNs.take()
        remote_NNs.put(NNs(_callable("representation"), _callable("prediction"), _callable("dynamics")))
        e.last_losses = dict(l_representation=float(losses[0]), l_prediction=float(losses[1]), l_dynamics=float(losses[2]))   # :421-424
    return True
