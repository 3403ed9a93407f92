"""Per-phase cycle breakdown of the control kernel (ARROY_B200_CTRL_TIMING=1) for few-tree builds."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ARROY_B200_CTRL_TIMING"] = "1"
import arroy_b200 as ab, bench
ctx = ab.Context(0)
n, d, metric, centre = 1_000_000, 768, "cosine", 0.5
items = torch.empty((n, d), dtype=torch.float32, device="cuda:0")
ctx.synth_device(bench.SEED, d, 0, n, centre, items.data_ptr())
ctx.stage_items_device(metric, np.arange(n, dtype=np.uint32), d, items.data_ptr())
for T in (1, 6):
    for cl in ("0", "8", "16"):
        os.environ["ARROY_B200_CLUSTER"] = cl
        seeds = bench.derive_seeds(ab, T)
        for rep in range(2):
            ctx.build_trees(seeds, list(range(T)), T, collect=False)
        st, bd = ctx.build_stats(), ctx.build_breakdown()
        attempts = st["create_split_calls"] / T
        print("T=%d cluster=%s: loop %.2f ms, %.1f us per attempt" % (T, cl, bd["loop_ms"], bd["loop_ms"] * 1e3 / attempts), flush=True)
