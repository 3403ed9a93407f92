"""Per-attempt chain latency: build few trees (async chains), loop time / device steps."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arroy_b200 as ab, bench
ctx = ab.Context(0)
for (n, d, metric, centre) in ((1_000_000, 768, "cosine", 0.5), (10_000, 64, "euclidean", 0.0)):
    items = torch.empty((n, d), dtype=torch.float32, device="cuda:0")
    ctx.synth_device(bench.SEED, d, 0, n, centre, items.data_ptr())
    ctx.stage_items_device(metric, np.arange(n, dtype=np.uint32), d, items.data_ptr())
    for T in (1, 6, 10, 50):
        seeds = bench.derive_seeds(ab, T)
        for rep in range(2):
            ctx.build_trees(seeds, list(range(T)), T, collect=False)
        st, bd = ctx.build_stats(), ctx.build_breakdown()
        attempts = st["create_split_calls"] / T
        print("n=%d d=%d T=%2d: loop %.2f ms, attempts/tree %.0f -> %.1f us per attempt; total build %.2f ms => %.0f vectors/s" %
              (n, d, T, bd["loop_ms"], attempts, bd["loop_ms"] * 1e3 / attempts, st["build_ms"], n / st["build_ms"] * 1e3), flush=True)
    del items
