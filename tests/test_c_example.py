"""The C ABI from plain C: examples/selfplay.c is compiled with gcc against include/muzero_b200.h and linked to the library.
CPU: it builds and refuses to run without a device (no fallback).  GPU: the whole self-play -> learner -> arena loop runs."""
import os
import subprocess

import pytest

import common

EXE = os.path.join(common.ROOT, "examples", "selfplay")


def build():
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(common.ROOT, "include"), "-o", EXE,
                           os.path.join(common.ROOT, "examples", "selfplay.c"), "-L", os.path.join(common.ROOT, "muzero.jl_b200"),
                           "-lmuzero_b200", "-Wl,-rpath," + os.path.join(common.ROOT, "muzero.jl_b200")])


def test_c_example_builds_and_fails_loudly_without_a_device():
    import torch
    from muzero_jl_b200 import capi
    capi.lib()
    build()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([EXE, "8", "5", "1"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_example_runs_on_the_gpu():
    build()
    r = subprocess.run([EXE, "512", "20", "10"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip().endswith("ok") and "arena vs expert" in r.stdout and "self-play: 512 games" in r.stdout
