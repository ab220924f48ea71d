"""Per-phase cycle breakdown of mz_k_search_rn's simulation loop (needs a -DMZ_PHASE_TIMERS build:
make -C muzero.jl_b200/csrc -B EXTRA=-DMZ_PHASE_TIMERS).  env: G, S"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi

G = int(os.environ.get("G", 8288)); S = int(os.environ.get("S", 50))
ctx = capi.Context(capi.resnet_config(num_slots=G, num_iters=S, replay_buffer_size=max(10000, G)))
ctx.init_weights(1337)
ctx.self_play(0, G, 1.0)
sims, moves = ctx.self_play(G, G, 1.0)
raw = ctx.phase_cycles().astype(float)
t = raw[28:36]; n = raw[36]
names = ["select", "stage hidden (x2 per sim)", "representation steps (root)", "prediction steps", "dynamics steps", "expand+backup", "-", "loop overhead"]
rounds = n * S
print("CTA launches", n, "cycles per simulation round: %.0f" % (t.sum() / rounds))
for i, nm in enumerate(names):
    if nm != "-":
        print("  %-32s %10.0f" % (nm, t[i] / rounds))
lab = ["wait weights", "issue MMAs", "wait MMA", "epilogue", "fence + CTA barrier", "between steps"]
for w, who in ((0, "thread 0 (warpgroup 0, issues the MMAs)"), (1, "thread 128 (warpgroup 1)")):
    v = raw[37 + 7 * w: 37 + 7 * w + 7]
    if v[6] > 0:
        print("per step, %s: %.0f cycles; " % (who, v[:6].sum() / v[6]) + ", ".join("%s %.0f" % (lab[i], v[i] / v[6]) for i in range(6)))
ctx.close()
