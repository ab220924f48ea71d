"""GPU tests of the ResNet networks on the tensor cores (net_type = MZ_NET_RESNET, nn_mode = MZ_NN_BF16_TC).
bf16 operands cannot be bit-exact against the Float32 oracle; the kernels are held to the oracle's bf16 emulation
(same rounding points: every stored activation and every weight is bfloat16, sums are Float32) with a stated tolerance,
and the searches to a visit-count agreement rate."""
import numpy as np
import pytest

import common
from oracle import oracle as O
from test_oracle_resnet import _randomised_blob

pytestmark = pytest.mark.gpu
RN_ATOL = 2e-2      # one bf16 ulp of an O(1) activation is 4e-3; a few layers of re-rounding on top of tensor-core summation order


@pytest.fixture(scope="module")
def capi():
    from muzero_jl_b200 import capi
    return capi


def make(capi, **kw):
    kw.setdefault("num_slots", 128); kw.setdefault("replay_buffer_size", 512)
    cfg = capi.resnet_config(**kw)
    return capi.Context(cfg), common.oracle_config(cfg)


@pytest.mark.parametrize("kw", [dict(rn_kernel=1, rn_num_blocks=0), dict(rn_kernel=1, rn_num_blocks=1), dict(rn_kernel=1), dict(),
                                dict(rn_num_filters=32, depth_value=0, rn_second_head_filters=3)])
def test_resnet_networks_match_bf16_oracle(capi, kw):
    ctx, ocfg = make(capi, **kw)
    blob = _randomised_blob(ocfg, 7)
    assert ctx.num_params() == blob.shape[0]
    ctx.set_weights(blob)
    assert np.array_equal(ctx.get_weights(), blob)
    st, legal, tp = common.random_stacked(ocfg, 75, seed=4)
    O.set_bf16(True)
    try:
        h = ctx.representation(st)
        oh = np.stack([O.representation(ocfg, blob, x) for x in st])
        err = np.abs(h - oh)
        assert np.max(err) < RN_ATOL * max(1.0, np.max(np.abs(oh))) and np.median(err) < 1e-3, (np.max(err), np.median(err))
        v, p = ctx.prediction(oh)
        ov, op = zip(*[O.prediction(ocfg, blob, x) for x in oh])
        assert np.max(np.abs(v - np.array(ov))) < RN_ATOL and np.max(np.abs(p - np.stack(op))) < RN_ATOL
        sa = np.concatenate([2 * oh, np.repeat((np.arange(75) % 9 + 1)[:, None] / np.float32(9), 9, 1).astype(np.float32)], 1)
        nh, r = ctx.dynamics(sa)
        onh, orr = zip(*[O.dynamics(ocfg, blob, x) for x in sa])
        onh = np.stack(onh)
        assert np.max(np.abs(nh - onh)) < RN_ATOL * max(1.0, np.max(np.abs(onh))) and np.max(np.abs(r - np.array(orr))) < RN_ATOL
    finally:
        O.set_bf16(False)
    ctx.close()
