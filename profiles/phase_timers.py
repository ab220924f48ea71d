"""Per-phase cycle breakdown of the simulation loop (needs a -DMZ_PHASE_TIMERS build: make -C muzero.jl_b200/csrc -B EXTRA=-DMZ_PHASE_TIMERS)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi

G, S = 4096, 50
names = ["select", "stage", "-", "network(+wait)", "-", "hidden write(+wait)", "expand", "backup"]
for mode, label in ((capi.NN_FP32_EXACT, "fp32_exact"), (capi.NN_BF16_TC, "bf16_tcgen05"), (capi.NN_SPLIT_MMA, "split_mma")):
    ctx = capi.Context(capi.default_config(num_slots=G, num_iters=S, replay_buffer_size=10000, nn_mode=mode))
    ctx.init_weights(1337)
    ctx.self_play(0, G, 1.0)
    sims, moves = ctx.self_play(G, G, 1.0)
    raw = ctx.phase_cycles().astype(float)
    pc = raw[28:46].reshape(2, 9)
    for g, who in ((0, "thread 0 (tree lane / prediction group)"), (1, "thread 128 (dynamics group)")):
        n = pc[g, 8]; rounds = n * S
        if n == 0:
            print(label, who, "no timers in this build"); continue
        tot = pc[g, :8].sum()
        print("%s %s: %.0f cycles/round; " % (label, who, tot / rounds) + ", ".join("%s %.0f" % (names[i], pc[g, i] / rounds) for i in range(8) if names[i] != "-"))
    for o, who in ((12, 'issuing thread, group 0'), (20, 'epilogue thread, group 1')):
        q = raw[o + 6]
        if q > 0:
            lab = ['issue', 'mbarrier wait', 'tcgen05.ld', 'epilogue', 'fence.proxy.async', 'group barrier']
            if mode == capi.NN_SPLIT_MMA:      # per simulation of one dynamics warp (o = 12: trunk + state head, 20: trunk + reward head)
                lab = ['wait for weights', 'load input', 'MMAs', 'epilogue', 'release slot + refill', '-']
            print('%s TC round (%s): %.0f cycles; ' % (label, who, raw[o:o + 6].sum() / q) + ', '.join('%s %.0f' % (lab[i], raw[o + i] / q) for i in range(6)))
    ctx.close()
