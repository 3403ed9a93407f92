"""Chain latency of the forest build (C2 items): per-attempt time for few / many trees, speculative two_means on / off,
per-phase cycle counters of the control kernel for T = 1."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arroy_b200 as ab, bench
ctx = ab.Context(0)
n, d, metric, centre = 1_000_000, 768, "cosine", 0.5
items = torch.empty((n, d), dtype=torch.float32, device="cuda:0")
ctx.synth_device(bench.SEED, d, 0, n, centre, items.data_ptr())
ctx.stage_items_device(metric, np.arange(n, dtype=np.uint32), d, items.data_ptr())
variants = [v.split(",") for v in (sys.argv[1:] or ["SPEC=1", "SPEC=0"])]
for var in variants:
    env = {}
    for kv in var:
        k, v = kv.split("=")
        env["ARROY_B200_" + k] = v
    for T in (1, 7, 50):
        for k, v in env.items():
            os.environ[k] = v
        if T == 1:
            os.environ["ARROY_B200_CTRL_TIMING"] = "1"
        seeds = bench.derive_seeds(ab, T)
        for rep in range(2):
            ctx.build_trees(seeds, list(range(T)), T, collect=False)
        st, bd = ctx.build_stats(), ctx.build_breakdown()
        os.environ.pop("ARROY_B200_CTRL_TIMING", None)
        for k in env:
            os.environ.pop(k)
        attempts = st["create_split_calls"] / T
        print("%s T=%d: loop %.2f ms, %.1f us per attempt, misspeculated %d of %d" % (" ".join(var), T, bd["loop_ms"], bd["loop_ms"] * 1e3 / attempts, st["misspeculated_splits"], st["create_split_calls"]), flush=True)
