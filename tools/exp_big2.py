"""10M x 768 Cosine, 100 trees in waves: loop time with / without the bf16 shadow (graph schedule)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arroy_b200 as ab, bench
ctx = ab.Context(0)
n, d, metric, T = int(os.environ.get("N", 10_000_000)), 768, "cosine", int(os.environ.get("T", 100))
items = torch.empty((n, d), dtype=torch.float32, device="cuda:0")
ctx.synth_device(bench.SEED, d, 0, n, 0.5, items.data_ptr())
ctx.stage_items_device(metric, np.arange(n, dtype=np.uint32), d, items.data_ptr())
del items
seeds = bench.derive_seeds(ab, T)
for var in sys.argv[1:] or ["SHADOW=1", "SHADOW=0"]:
    for kv in var.split(","):
        k, v = kv.split("="); os.environ["ARROY_B200_" + k] = v
    for rep in range(2):
        t0 = time.time(); ctx.build_trees(seeds, list(range(T)), T, collect=False); wall = time.time() - t0
    bd = ctx.build_breakdown()
    print(var, "loop %.1f ms wall %.1f ms" % (bd["loop_ms"], wall * 1e3), ctx.build_shadow_stats(), ctx.build_stats()["scanned_rows"], flush=True)
    for kv in var.split(","):
        os.environ.pop("ARROY_B200_" + kv.split("=")[0])
