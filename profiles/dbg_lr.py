"""Debug / accuracy probe of the tensor-core BPTT path: gradients of one batch against the fp32 SIMT kernel and the oracle, per layer."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from muzero_jl_b200 import capi
import common
from oracle import oracle as O
B = int(os.environ.get("B", 64))
kw = dict(num_slots=256, replay_buffer_size=1024, batch_size=B, intermediate_rewards=int(os.environ.get("IR", 0)))
kw.update(eval(os.environ.get("KW", "{}")))
ctx = capi.Context(capi.default_config(nn_mode=capi.NN_SPLIT_MMA, **kw)); ocfg = common.oracle_config(ctx.cfg)
ctx.init_weights(5); blob = ctx.get_weights()
rng = np.random.default_rng(1); blob = (blob + np.where(blob == 0, rng.uniform(-0.1, 0.1, blob.shape), 0)).astype(np.float32); ctx.set_weights(blob)
ex = capi.Context(capi.default_config(**kw)); ex.set_weights(blob); ex.self_play(0, 300, 1.0); ctx.history_import(ex.history_export())
batch = ctx.get_batch(2)
g_tc, l_tc = ctx.learn_gradients(batch, capi.GRAD_BPTT)
os.environ["MUZERO_B200_BPTT_SIMT"] = "1"
g_s, l_s = ctx.learn_gradients(batch, capi.GRAD_BPTT)
del os.environ["MUZERO_B200_BPTT_SIMT"]
print("losses tc", l_tc, "simt", l_s)
# per layer report (blob order: per layer W (out*in) then b)
from muzero_jl_b200.capi import sizes
cfg = ctx.cfg
def layers():
    w = cfg.width_hidden; st = sizes(cfg)
    nets = [[(st["stack"], w)] + [(w, w)] * cfg.depth_representation + [(w, st["hidden"])],
            [(st["hidden"], w)] + [(w, w)] * cfg.depth_prediction + [(w, w)] * cfg.depth_value + [(w, 1)] + [(w, w)] * cfg.depth_policy + [(w, cfg.A)],
            [(st["sa"], w)] + [(w, w)] * cfg.depth_dynamics + [(w, w)] * cfg.depth_state_head + [(w, st["hidden"])] + [(w, w)] * cfg.depth_reward + [(w, 1)]]
    return [x for n in nets for x in n]
off = 0
gm = np.max(np.abs(g_s))
for i, (fin, fout) in enumerate(layers()):
    nw = fin * fout
    dw = g_tc[off:off + nw] - g_s[off:off + nw]; db = g_tc[off + nw:off + nw + fout] - g_s[off + nw:off + nw + fout]
    sw = np.max(np.abs(g_s[off:off + nw] - 2 * blob[off:off + nw]))
    print("L%2d in %2d out %2d: max|dW err| %.3e (data grad max %.3e)  max|db err| %.3e (db max %.3e)" % (i, fin, fout, np.max(np.abs(dw)), sw, np.max(np.abs(db)), np.max(np.abs(g_s[off + nw:off + nw + fout] - 2 * blob[off + nw:off + nw + fout]))))
    off += nw + fout
print("overall max err %.3e of max grad %.3e" % (np.max(np.abs(g_tc - g_s)), gm))
