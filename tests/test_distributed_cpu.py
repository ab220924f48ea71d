"""CPU, gloo, world_size 2: the host-side multi-rank logic (game-id sharding, unique-id broadcast, gradient
averaging).  Per-game results must not depend on how games are sharded over ranks."""
import os
import pickle
import socket
import subprocess
import sys

import numpy as np

import common
from oracle import oracle as O


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_sharded_self_play_matches_single_rank(tmp_path):
    port = free_port(); out = str(tmp_path / "out.pkl")
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), MZ_DIST_OUT=out)
        procs.append(subprocess.Popen([sys.executable, os.path.join(common.ROOT, "tests", "dist_worker.py")], env=env))
    for p in procs:
        assert p.wait(timeout=240) == 0
    res = pickle.load(open(out, "rb"))
    assert [r["rank"] for r in res] == [0, 1] and (res[0]["lo"], res[0]["cnt"], res[1]["lo"], res[1]["cnt"]) == (500, 12, 512, 12)
    cfg = O.default_config(num_iters=8)
    ref = O.self_play(cfg, O.init_weights(cfg, 1337), 500, 24, 1.0, 1)
    assert sum(r["sims"] for r in res) == ref["sims"]
    for key, rk in (("T", "T"), ("actions", "actions"), ("child_visits", "cv"), ("root_values", "rv")):
        assert np.array_equal(np.concatenate([r[rk] for r in res]), ref[key]), key
    expect_uid = (np.arange(128) * 7 % 251).astype(np.uint8)
    assert all(np.array_equal(r["uid"], expect_uid) and r["grad_ok"] for r in res)


def test_shard_games_partitions_exactly():
    from muzero_jl_b200 import dist as mzdist
    for world in (1, 2, 3, 8):
        parts = [mzdist.shard_games(r, world, 7, 4099) for r in range(world)]
        assert parts[0][0] == 7 and sum(c for _, c in parts) == 4099
        for (lo, c), (lo2, _) in zip(parts, parts[1:]):
            assert lo + c == lo2
