"""One tree, 1M x 768 cosine (lockstep, for ncu source-level sampling of control_kernel)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arroy_b200 as ab, bench
ctx = ab.Context(0)
n, d = 1_000_000, 768
items = torch.empty((n, d), dtype=torch.float32, device="cuda:0")
ctx.synth_device(bench.SEED, d, 0, n, 0.5, items.data_ptr())
ctx.stage_items_device("cosine", np.arange(n, dtype=np.uint32), d, items.data_ptr())
seeds = bench.derive_seeds(ab, 1)
ctx.build_trees(seeds, [0], 1, collect=False)
print(ctx.build_breakdown())
