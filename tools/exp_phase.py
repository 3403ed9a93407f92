"""Per-phase cycle counts of the control CTAs (ARROY_B200_CTRL_TIMING) for T trees on the C2 items."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ARROY_B200_CTRL_TIMING"] = "1"
import arroy_b200 as ab, bench
ctx = ab.Context(0)
n, d, metric = 1_000_000, 768, "cosine"
items = torch.empty((n, d), dtype=torch.float32, device="cuda:0")
ctx.synth_device(bench.SEED, d, 0, n, 0.5, items.data_ptr())
ctx.stage_items_device(metric, np.arange(n, dtype=np.uint32), d, items.data_ptr())
for T in [int(x) for x in (sys.argv[1:] or ["1", "7"])]:
    seeds = bench.derive_seeds(ab, T)
    for rep in range(2):
        ctx.build_trees(seeds, list(range(T)), T, collect=False)
    bd = ctx.build_breakdown()
    print("T=%d loop %.2f ms" % (T, bd["loop_ms"]), flush=True)
