"""Forest-build loop time on the C2 items for several forest sizes and env settings: python tools/exp_chain6.py "A=1,B=2" "..." """
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arroy_b200 as ab, bench
ctx = ab.Context(0)
n, d, metric = 1_000_000, 768, "cosine"
items = torch.empty((n, d), dtype=torch.float32, device="cuda:0")
ctx.synth_device(bench.SEED, d, 0, n, 0.5, items.data_ptr())
ctx.stage_items_device(metric, np.arange(n, dtype=np.uint32), d, items.data_ptr())
TS = [int(x) for x in os.environ.get("TS", "1,7,13,25,50").split(",")]
for var in sys.argv[1:] or ["PERSIST=1"]:
    row = []
    for T in TS:
        for kv in var.split(","):
            k, v = kv.split("=")
            os.environ["ARROY_B200_" + k] = v
        seeds = bench.derive_seeds(ab, T)
        for rep in range(2):
            ctx.build_trees(seeds, list(range(T)), T, collect=False)
        bd = ctx.build_breakdown()
        for kv in var.split(","):
            os.environ.pop("ARROY_B200_" + kv.split("=")[0])
        row.append("T=%d %.1f" % (T, bd["loop_ms"]))
    print(var, " | ".join(row), flush=True)
