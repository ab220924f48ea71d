"""muzero.jl_b200 -- B200-native self-play / learner hot path of deveshjawla/MuZero.jl.

The compute lives in ``libmuzero_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/muzero_b200.h``).  This package is the Python host-side mirror of the reference's Julia interface
(``Config``, ``FeedForwardHP``, ``GameHistory``, ``run_mcts``, ``play_game``, ``self_play!``, ``save_game``,
``get_batch``, ``learning!``); the Julia ``ccall`` wrapper with the same names is ``julia/MuZeroB200.jl``.
There is no CPU fallback: importing ``capi`` without the built library, or creating a context without a
CUDA device, raises.
"""
from . import capi  # noqa: F401
from .capi import Context, MzConfig, MuZeroB200Error, default_config, build_library  # noqa: F401
from .api import (Config, FeedForwardHP, ResNetHP, GameHistory, TicTacToe, init_networks, run_mcts, select_action, play_game,  # noqa: F401
                  self_play, competitive_play, save_game, get_batch, learning, ReplayBuffer, Engine)
from . import dropin  # noqa: F401  (the reference's entry points with the reference's own signatures)
