"""Device time of the batched network callables (one network pass over 32 rows per CTA) for every network mode."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_jl_b200 import capi
B = 148 * 32
REP = int(os.environ.get("MZ_MMA_PROBE_REPEAT", "1"))
rng = np.random.default_rng(0)
for mode, name in ((capi.NN_SPLIT_MMA, "split_mma"), (capi.NN_BF16_TC, "bf16_tcgen05"), (capi.NN_FP32_EXACT, "fp32_exact")):
    ctx = capi.Context(capi.default_config(num_slots=64, nn_mode=mode)); ctx.init_weights(1337)
    for net, fn, dim in (("representation", ctx.representation, 63), ("prediction", ctx.prediction, 27), ("dynamics", ctx.dynamics, 36)):
        x = rng.uniform(-1, 1, (B, dim)).astype(np.float32)
        fn(x)
        ctx.kernel_time_reset(True)
        for _ in range(5):
            fn(x)
        ms, n = ctx.kernel_time(5)
        print("%-13s %-15s %.1f us per launch = %.0f cycles (one CTA per SM, 32 rows each)%s" % (name, net, 1e3 * ms / n, 1e3 * ms / n * 1965, ("; %.0f cycles per pass over the network" % (1e3 * ms / n * 1965 / REP)) if REP > 1 and mode == capi.NN_SPLIT_MMA else ""))
        ctx.kernel_time_reset(False)
    ctx.close()
