"""BASELINE config 5: 4096 queries x 100k shared candidates, d = 768 (Cosine)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arroy_b200 as ab, bench
n, d, nq, nc, k = 200_000, 768, int(os.environ.get("NQ", 4096)), 100_000, 100
ctx = ab.Context(0)
items = torch.empty((n, d), dtype=torch.float32, device="cuda:0")
ctx.synth_device(bench.SEED, d, 0, n, 0.5, items.data_ptr())
ctx.stage_items_device("cosine", np.arange(n, dtype=np.uint32), d, items.data_ptr())
h0, _ = ctx.item_headers()
q = items[n - nq:].cpu().numpy()            # queries = fresh rows after the candidates
qh = h0[n - nq:]
rows = np.arange(nc, dtype=np.uint32)
for rep in range(6):
    os.environ["ARROY_B200_XRERANK"] = "exact" if rep < 2 else "filter"
    ctx.timer_start(); t0 = time.perf_counter()
    out = ctx.rerank_shared(q, qh, rows, k)
    ms = ctx.timer_stop()
    flop = 2.0 * nq * nc * d
    print(os.environ["ARROY_B200_XRERANK"], ctx.rerank_stats(), end=" ")
    print("rerank_shared %dx%d: wall %.1f ms, stream %.1f ms, %.1f TFLOP/s fp32 (%.1f TFMA/s)" % (nq, nc, (time.perf_counter() - t0) * 1e3, ms, flop / ms / 1e9, flop / 2 / ms / 1e9), flush=True)
m = 128
offs = (np.arange(m + 1, dtype=np.uint64) * np.uint64(nc))
t0 = time.perf_counter()
ref = ctx.rerank_batch(q[:m], qh[:m], np.tile(rows, m), offs, k)
print("rerank_batch (generic kernel) %dx%d: %.1f ms  -> extrapolated to %d queries: %.0f ms" % (m, nc, (time.perf_counter() - t0) * 1e3, nq, (time.perf_counter() - t0) * 1e3 * nq / m))
assert ref[0].tolist() == out[0][:m].tolist() and ref[1].tobytes() == out[1][:m].tobytes()
print("ids and distances identical to the per-pair kernel")
