"""Single-query latency on the C2 index: wall clock per query and the device-side breakdown of the latency path."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arroy_b200 as ab, bench
n, d, T, metric = 1_000_000, 768, 50, "cosine"
env = ab.Env(0)
ctx = env.ctx
items = torch.empty((n, d), dtype=torch.float32, device="cuda:0")
ctx.synth_device(bench.SEED, d, 0, n, 0.5, items.data_ptr())
host = items.cpu().numpy()
w = ab.Writer(env, 0, d, metric)
w.add_items(np.arange(n, dtype=np.uint32), host)
w.builder(ab.StdRng.from_seed(bench.SEED)).n_trees(T).build()
r = ab.Reader.open(env, 0, metric)
r.nns(100).by_item(0)
lat, bds = [], []
for i in range(200):
    t0 = time.perf_counter()
    r.nns(100).by_item(i)
    lat.append(time.perf_counter() - t0)
    bds.append(ctx.search_breakdown())
lat.sort()
keys = list(bds[0].keys())
print("by_item: p50 %.3f ms p99 %.3f ms, %.0f QPS" % (lat[100] * 1e3, lat[198] * 1e3, 200 / sum(lat)))
print("device breakdown (ms, mean):", {k: round(sum(b[k] for b in bds) / len(bds), 4) for k in keys})
q = host[5] * 0.5 + host[9] * 0.5
t0 = time.perf_counter()
for i in range(100):
    r.nns(100).by_vector(q)
print("by_vector: %.3f ms per query" % ((time.perf_counter() - t0) * 10))
