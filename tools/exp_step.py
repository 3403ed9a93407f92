"""Per-step wall time of the bench's value step (stage device-to-device + build) on C2, with the library's own breakdown."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arroy_b200 as ab, bench
ctx = ab.Context(0)
n, d, T, metric = 1_000_000, 768, 50, "cosine"
items = torch.empty((n, d), dtype=torch.float32, device="cuda:0")
ctx.synth_device(bench.SEED, d, 0, n, 0.5, items.data_ptr())
ids = np.arange(n, dtype=np.uint32)
seeds = bench.derive_seeds(ab, T)
rows = []
for i in range(12):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.stage_items_device(metric, ids, d, items.data_ptr())
    t1 = time.perf_counter()
    ctx.build_trees_begin(seeds)
    t2 = time.perf_counter()
    bd = ctx.build_breakdown()
    rows.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), bd["setup_ms"], bd["loop_ms"], bd["d2h_ms"]))
for r in rows[2:]:
    print("stage %.2f ms | build_begin %.2f ms (setup %.2f loop %.2f d2h %.2f, other %.2f)" % (r[0], r[1], r[2], r[3], r[4], r[1] - r[2] - r[3] - r[4]))
