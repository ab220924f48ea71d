"""Oracle tests of the repaired ResNet networks (net_type = 1, src/Learning.jl:148-255): an independent numpy
(Float64) restatement of Conv (true convolution) / BatchNorm (test mode) / residual blocks / heads must agree with the C
oracle; bf16 emulation stays close to Float32; MCTS and self-play run on the (W,H,nf) hidden state."""
import numpy as np
import pytest

from oracle import oracle as O

f32 = np.float32


def _units(c):
    """(kind, k, cin, cout, act) in blob order for the three nets + offsets, restated independently."""
    nf, nb, K, cells = c.rn_num_filters, c.rn_num_blocks, c.rn_kernel, c.W * c.H
    planes = c.C * (c.stacked_observations + 1) + c.stacked_observations

    def tower(k, cin):
        return [("conv", k, cin, nf, "relu")] + [("conv", k, nf, nf, a) for _ in range(nb) for a in ("relu", "id")]

    def head(f, first_act, out, out_act):
        return [("conv", 1, nf, f, "relu"), ("dense", 1, cells * f, c.width_hidden, first_act)] + \
               [("dense", 1, c.width_hidden, c.width_hidden, "relu")] * c.depth_value + [("dense", 1, c.width_hidden, out, out_act)]
    rep = tower(K, planes)
    pred = tower(1, nf) + head(c.rn_first_head_filters, "relu", 1, "tanh") + head(c.rn_second_head_filters, "id", c.A, "id")
    dyn = tower(1, nf + 1) + tower(1, nf) + head(c.rn_first_head_filters, "relu", 1, "tanh")
    return rep, pred, dyn


class NpNet:
    def __init__(self, c, blob):
        self.c = c; self.W, self.H = c.W, c.H
        self.nets = []
        off = 0
        for units in _units(c):
            out = []
            for kind, k, cin, cout, act in units:
                u = dict(kind=kind, k=k, cin=cin, cout=cout, act=act)
                if kind == "conv":
                    u["w"] = blob[off:off + k * k * cin * cout].astype(np.float64).reshape((k, k, cin, cout), order="F"); off += k * k * cin * cout
                    for name in ("b", "beta", "gamma", "mu", "var"):
                        u[name] = blob[off:off + cout].astype(np.float64); off += cout
                else:
                    u["w"] = blob[off:off + cin * cout].astype(np.float64).reshape((cout, cin), order="F"); off += cin * cout
                    u["b"] = blob[off:off + cout].astype(np.float64); off += cout
                out.append(u)
            self.nets.append(out)
        self.n_params = off

    @staticmethod
    def act(x, a):
        return np.maximum(x, 0) if a == "relu" else np.tanh(x) if a == "tanh" else x

    def conv(self, u, x, skip=None):   # x: (W,H,cin)
        k, pad = u["k"], u["k"] // 2
        y = np.zeros((self.W, self.H, u["cout"]))
        for a in range(k):
            for b in range(k):
                for i in range(self.W):
                    for j in range(self.H):
                        si, sj = i + pad - a, j + pad - b
                        if 0 <= si < self.W and 0 <= sj < self.H:
                            y[i, j, :] += x[si, sj, :] @ u["w"][a, b, :, :]
        y = y + u["b"]
        y = u["gamma"] * (y - u["mu"]) / np.sqrt(u["var"] + np.float64(f32(1e-5))) + u["beta"]
        if skip is not None:
            y = y + skip
        return self.act(y, u["act"])

    def tower(self, units, x):
        x = self.conv(units[0], x)
        for b in range(self.c.rn_num_blocks):
            t = self.conv(units[1 + 2 * b], x)
            u2 = dict(units[2 + 2 * b]); u2["act"] = "relu"
            x = self.conv(u2, t, skip=x)
        return x

    def head(self, units, t):
        f = self.conv(units[0], t).reshape(-1, order="F")
        for u in units[1:]:
            f = self.act(u["w"] @ f + u["b"], u["act"])
        return f

    def representation(self, stacked):
        planes = stacked.size // (self.W * self.H)
        return self.tower(self.nets[0], stacked.astype(np.float64).reshape((self.W, self.H, planes), order="F")).reshape(-1, order="F")

    def prediction(self, h):
        nb2 = 1 + 2 * self.c.rn_num_blocks
        nh = 3 + self.c.depth_value
        t = self.tower(self.nets[1][:nb2], h.astype(np.float64).reshape((self.W, self.H, -1), order="F"))
        v = self.head(self.nets[1][nb2:nb2 + nh], t)
        z = self.head(self.nets[1][nb2 + nh:], t)
        e = np.exp(z - z.max())
        return v[0], e / e.sum()

    def dynamics(self, sa):
        nb2 = 1 + 2 * self.c.rn_num_blocks
        x = sa.astype(np.float64).reshape((self.W, self.H, -1), order="F")
        t = self.tower(self.nets[2][:nb2], x)
        s = self.tower(self.nets[2][nb2:2 * nb2], t)
        r = self.head(self.nets[2][2 * nb2:], t)
        return s.reshape(-1, order="F"), r[0]


def _randomised_blob(c, seed):
    """Glorot init + randomised conv biases and BatchNorm parameters / statistics (the defaults 0/1 would hide mistakes)."""
    blob = O.init_weights(c, seed)
    rng = np.random.default_rng(seed)
    off = 0
    for units in _units(c):
        for kind, k, cin, cout, act in units:
            nw = k * k * cin * cout
            off += nw
            if kind == "conv":
                blob[off:off + cout] = rng.normal(0, 0.1, cout); off += cout                       # b
                blob[off:off + cout] = rng.normal(0, 0.1, cout); off += cout                       # beta
                blob[off:off + cout] = rng.uniform(0.5, 1.5, cout); off += cout                    # gamma
                blob[off:off + cout] = rng.normal(0, 0.1, cout); off += cout                       # mu
                blob[off:off + cout] = rng.uniform(0.5, 2.0, cout); off += cout                    # var
            else:
                blob[off:off + cout] = rng.normal(0, 0.1, cout); off += cout
    assert off == blob.shape[0]
    return blob


@pytest.mark.parametrize("kw", [dict(), dict(rn_kernel=1, rn_num_blocks=1, rn_num_filters=16), dict(rn_num_filters=32, depth_value=0, rn_second_head_filters=3)])
def test_resnet_matches_numpy_restatement(kw):
    c = O.resnet_config(**kw)
    nr, np_, nd = (O.num_params(c, i) for i in range(3))
    blob = _randomised_blob(c, 5)
    ref = NpNet(c, blob)
    assert ref.n_params == nr + np_ + nd == O.num_params(c)
    rng = np.random.default_rng(1)
    s = O.sizes(c)
    stacked = rng.integers(0, 2, s["stack"]).astype(f32)
    h = O.representation(c, blob, stacked)
    assert h.shape[0] == c.W * c.H * c.rn_num_filters
    assert np.allclose(h, ref.representation(stacked), rtol=1e-4, atol=1e-5)
    v, p = O.prediction(c, blob, h)
    rv, rp = ref.prediction(h)
    assert abs(v - rv) < 1e-5 and np.allclose(p, rp, rtol=1e-4, atol=1e-6) and abs(p.sum() - 1) < 1e-5
    sa = np.concatenate([2 * h, np.full(c.W * c.H, f32(5.0 / 9.0), f32)]).astype(f32)
    nh, r = O.dynamics(c, blob, sa)
    rh, rr = ref.dynamics(sa)
    assert np.allclose(nh, rh, rtol=1e-4, atol=1e-5) and abs(r - rr) < 1e-5


def test_resnet_kat_parameter_counts():
    c = O.resnet_config()   # nf = 64, 2 blocks, 3x3 representation, TicTacToe: planes = 7, cells = 9, hs = 64, A = 9
    conv = lambda k, ci, co: k * k * ci * co + 5 * co
    dense = lambda i, o: i * o + o
    rep = conv(3, 7, 64) + 4 * conv(3, 64, 64)
    head = lambda f, out: conv(1, 64, f) + dense(9 * f, 64) + dense(64, 64) + dense(64, out)
    pred = conv(1, 64, 64) + 4 * conv(1, 64, 64) + head(1, 1) + head(2, 9)
    dyn = conv(1, 65, 64) + 4 * conv(1, 64, 64) + 5 * conv(1, 64, 64) + head(1, 1)
    assert (O.num_params(c, 0), O.num_params(c, 1), O.num_params(c, 2)) == (rep, pred, dyn)


def test_resnet_identity_batchnorm_defaults():
    """Freshly initialised BatchNorm (beta 0, gamma 1, mu 0, var 1) divides by sqrt(1 + 1f-5): a 1x1 ConvBN of an all-ones
    kernel on a one-hot input reproduces it scaled by 1/sqrt(1.00001)."""
    c = O.resnet_config(rn_kernel=1, rn_num_blocks=0, rn_num_filters=4)
    blob = O.init_weights(c, 1)
    planes = 7
    blob[:planes * 4] = 1.0                       # representation ConvBN: w[ci, co] = 1
    stacked = np.zeros(O.sizes(c)["stack"], f32); stacked[4] = 1.0   # cell 4 of plane 0
    h = O.representation(c, blob, stacked)
    expect = np.zeros_like(h); expect[4::9] = f32(1.0) / np.sqrt(f32(1.0) + f32(1e-5))
    assert np.allclose(h, expect, rtol=1e-7, atol=0)


def test_resnet_bf16_emulation_is_close_and_mcts_runs():
    c = O.resnet_config(num_iters=12, exploration_eps=0.0)
    blob = _randomised_blob(c, 2)
    rng = np.random.default_rng(3)
    stacked = rng.integers(0, 2, O.sizes(c)["stack"]).astype(f32)
    h = O.representation(c, blob, stacked); v, p = O.prediction(c, blob, h)
    O.set_bf16(True)
    try:
        hb = O.representation(c, blob, stacked); vb, pb = O.prediction(c, blob, hb)
    finally:
        O.set_bf16(False)
    assert np.max(np.abs(h - hb)) < 0.05 * max(1.0, np.max(np.abs(h))) and abs(v - vb) < 0.05 and np.max(np.abs(p - pb)) < 0.02
    vc, rv, pri = O.run_mcts(c, blob, stacked, 0b111101111, 1, False, 7, 1)
    assert vc.sum() == 12 and vc[4] == 0 and abs(pri.sum() - 1) < 1e-5
    hist = O.self_play(c, blob, 0, 3, 1.0, 1)
    assert hist["sims"] == int(hist["T"].sum()) * 12 and hist["T"].min() >= 5
