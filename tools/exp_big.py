"""Persistent vs per-attempt-launch schedule on a bandwidth-bound build (10M x 768 Cosine, 100 trees) and on C2."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arroy_b200 as ab, bench
ctx = ab.Context(0)
for n, T in ((1_000_000, 50), (10_000_000, 100)):
    d, metric = 768, "cosine"
    items = torch.empty((n, d), dtype=torch.float32, device="cuda:0")
    ctx.synth_device(bench.SEED, d, 0, n, 0.5, items.data_ptr())
    ctx.stage_items_device(metric, np.arange(n, dtype=np.uint32), d, items.data_ptr())
    seeds = bench.derive_seeds(ab, T)
    for var in sys.argv[1:] or ["PERSIST=1", "PERSIST=0"]:
        for kv in var.split(","):
            k, v = kv.split("=")
            os.environ["ARROY_B200_" + k] = v
        for rep in range(2):
            ctx.build_trees(seeds, list(range(T)), T, collect=False)
        st, bd = ctx.build_stats(), ctx.build_breakdown()
        for kv in var.split(","):
            os.environ.pop("ARROY_B200_" + kv.split("=")[0])
        print("n=%d T=%d %s: loop %.1f ms, %.0f GB/s" % (n, T, var, bd["loop_ms"], st["scanned_rows"] * d * 4 / bd["loop_ms"] / 1e6), flush=True)
    del items
    torch.cuda.empty_cache()
