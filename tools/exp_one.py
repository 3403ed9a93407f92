"""One C2 forest build (for ncu captures): python tools/exp_one.py [T] [n]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arroy_b200 as ab, bench
T = int(sys.argv[1]) if len(sys.argv) > 1 else 50
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
d, metric = 768, "cosine"
ctx = ab.Context(0)
items = torch.empty((n, d), dtype=torch.float32, device="cuda:0")
ctx.synth_device(bench.SEED, d, 0, n, 0.5, items.data_ptr())
ctx.stage_items_device(metric, np.arange(n, dtype=np.uint32), d, items.data_ptr())
seeds = bench.derive_seeds(ab, T)
ctx.build_trees(seeds, list(range(T)), T, collect=False)
st, bd = ctx.build_stats(), ctx.build_breakdown()
print("T=%d n=%d loop %.1f ms kernel %.1f ms, %.0f GB/s" % (T, n, bd["loop_ms"], st["scan_ms"], st["scanned_rows"] * d * 4 / max(st["scan_ms"], 1e-9) / 1e6))
