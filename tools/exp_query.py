import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arroy_b200 as ab, bench
n, d, T = int(os.environ.get("N", 200000)), 768, int(os.environ.get("T", 20))
ctx = ab.Context(0)
items = torch.empty((n, d), dtype=torch.float32, device="cuda:0")
ctx.synth_device(bench.SEED, d, 0, n, 0.5, items.data_ptr())
host = items.cpu().numpy()
env = ab.Env(0); env._ctx = ctx
w = ab.Writer(env, 0, d, "cosine")
w.add_items(np.arange(n, dtype=np.uint32), host)
w.builder(ab.StdRng.from_seed(bench.SEED)).n_trees(T).build()
print("built", w.build_timings())
r = ab.Reader.open(env, 0, "cosine")
q = np.arange(1000, dtype=np.uint32)
os.environ["ARROY_B200_TRACE"] = "1"
for rep in range(3):
    t0 = time.perf_counter()
    out = r.nns_batch_by_item(q, 100)
    print("batch 1000: %.1f ms" % ((time.perf_counter() - t0) * 1e3), out[3], ctx.search_breakdown(), flush=True)
q10 = np.arange(10000, dtype=np.uint32) % n
for rep in range(2):
    t0 = time.perf_counter()
    out = r.nns_batch_by_item(q10, 100)
    print("batch 10000: %.1f ms" % ((time.perf_counter() - t0) * 1e3), ctx.search_breakdown(), flush=True)
os.environ["ARROY_B200_HOST_WALK"] = "1"
for rep in range(2):
    t0 = time.perf_counter()
    out = r.nns_batch_by_item(q, 100)
    print("host-walk batch 1000: %.1f ms" % ((time.perf_counter() - t0) * 1e3), out[3], flush=True)
env._ctx = None
