"""Incremental update timing: 1M x 768 Cosine, 50 trees, then +10k new items and 1k deletions."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arroy_b200 as ab, bench
n, d, T, add = int(os.environ.get("N", 1_000_000)), 768, 50, 10_000
ctx = ab.Context(0)
items = torch.empty((n + add, d), dtype=torch.float32, device="cuda:0")
ctx.synth_device(bench.SEED, d, 0, n + add, 0.5, items.data_ptr())
host = items.cpu().numpy()
env = ab.Env(0); env._ctx = ctx
w = ab.Writer(env, 0, d, "cosine")
w.add_items(np.arange(n, dtype=np.uint32), host[:n])
rng = ab.StdRng.from_seed(bench.SEED)
t0 = time.perf_counter(); w.builder(rng).n_trees(T).build(); print("fresh build %.0f ms" % ((time.perf_counter() - t0) * 1e3), w.build_timings(), flush=True)
w.add_items(np.arange(n, n + add, dtype=np.uint32), host[n:])
for i in range(0, 1000):
    w.del_item(i * 7)
t0 = time.perf_counter(); w.builder(rng).n_trees(T).build(); dt = time.perf_counter() - t0
print("incremental build (+%d items, -1000): %.0f ms" % (add, dt * 1e3), w.build_timings(), flush=True)
r = ab.Reader.open(env, 0, "cosine")
print("items", r.n_items(), "trees", r.n_trees(), "query", r.nns(5).by_item(n + 5)[:2])
env._ctx = None
