import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arroy_b200 as ab, bench
n, d = 1_000_000, 768
ctx = ab.Context(0)
items = torch.empty((n, d), dtype=torch.float32, device="cuda:0")
ctx.synth_device(bench.SEED, d, 0, n, 0.5, items.data_ptr())
ids = np.arange(n, dtype=np.uint32)
ctx.stage_items_device("cosine", ids, d, items.data_ptr())
h0, _ = ctx.item_headers()
host = items[:2000].cpu().numpy()
r = np.random.default_rng(0)
for nq, nc in ((1, 5700), (100, 5700), (1000, 5700), (1000, 5700), (4096, 5700)):
    lists = [np.sort(r.choice(n, size=nc, replace=False)).astype(np.uint32) for _ in range(min(nq, 64))]
    rows = np.concatenate([lists[i % len(lists)] for i in range(nq)])
    offs = (np.arange(nq + 1, dtype=np.uint64) * np.uint64(nc))
    q = host[np.arange(nq) % 2000]
    for rep in range(2):
        t0 = time.perf_counter()
        ctx.timer_start()
        out = ctx.rerank_batch(q, h0[np.arange(nq) % 2000], rows, offs, 100)
        ms = ctx.timer_stop()
        print("rerank_batch nq=%d nc=%d: wall %.2f ms, stream-event %.2f ms" % (nq, nc, (time.perf_counter() - t0) * 1e3, ms), flush=True)
