"""Forest-build loop time on the C2 items for several forest sizes (trees per GPU), persistent vs per-attempt launches."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arroy_b200 as ab, bench
ctx = ab.Context(0)
n, d, metric = 1_000_000, 768, "cosine"
items = torch.empty((n, d), dtype=torch.float32, device="cuda:0")
ctx.synth_device(bench.SEED, d, 0, n, 0.5, items.data_ptr())
ctx.stage_items_device(metric, np.arange(n, dtype=np.uint32), d, items.data_ptr())
for var in sys.argv[1:] or ["PERSIST=1", "PERSIST=0"]:
    row = []
    for T in (1, 7, 13, 25, 50):
        for kv in var.split(","):
            k, v = kv.split("=")
            os.environ["ARROY_B200_" + k] = v
        seeds = bench.derive_seeds(ab, T)
        for rep in range(2):
            ctx.build_trees(seeds, list(range(T)), T, collect=False)
        bd = ctx.build_breakdown()
        for kv in var.split(","):
            os.environ.pop("ARROY_B200_" + kv.split("=")[0])
        row.append("T=%d %.1f ms" % (T, bd["loop_ms"]))
    print(var, " | ".join(row), flush=True)
