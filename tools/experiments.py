"""Ad-hoc experiments on the GPU box (not part of the product): schedule variants and staging sweep."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arroy_b200 as ab  # noqa: E402
import bench  # noqa: E402

wl = bench.WORKLOADS[os.environ.get("WL", "c2")]
n, d, T, metric = wl["n"], wl["d"], wl["n_trees"], wl["metric"]
ctx = ab.Context(0)
items = torch.empty((n, d), dtype=torch.float32, device="cuda:0")
ctx.synth_device(bench.SEED, d, 0, n, wl["centre"], items.data_ptr())
ids = np.arange(n, dtype=np.uint32)
seeds = bench.derive_seeds(ab, T)


def build(env):
    for k, v in env.items():
        os.environ[k] = v
    ctx.stage_items_device(metric, ids, d, items.data_ptr())
    if metric == "dot-product":
        ctx.dot_preprocess()
    ctx.timer_start()
    ctx.build_trees(seeds, list(range(T)), T, collect=False)
    ms = ctx.timer_stop()
    st = ctx.build_stats()
    bd = ctx.build_breakdown()
    for k in env:
        os.environ.pop(k)
    return ms, st, bd


what = sys.argv[1:] or ["sched", "stage"]
if "sched" in what:
    for name, env in [("async", {}), ("async", {}), ("lockstep", {"ARROY_B200_LOCKSTEP": "1"}), ("lockstep", {"ARROY_B200_LOCKSTEP": "1"}),
                      ("lockstep+interleave", {"ARROY_B200_LOCKSTEP": "1", "ARROY_B200_INTERLEAVE": "1"}),
                      ("lockstep+interleave", {"ARROY_B200_LOCKSTEP": "1", "ARROY_B200_INTERLEAVE": "1"}),
                      ("profile", {"ARROY_B200_PROFILE": "1", "ARROY_B200_TRACE": "1"}),
                      ("profile+interleave", {"ARROY_B200_PROFILE": "1", "ARROY_B200_TRACE": "1", "ARROY_B200_INTERLEAVE": "1"})]:
        ms, st, bd = build(env)
        print("SCHED %-22s total %.1f ms loop %.1f scan_ms %.1f steps %d GB/s(whole) %.0f" % (name, ms, bd["loop_ms"], st["scan_ms"], st["steps"], st["scanned_rows"] * d * 4 / ms / 1e6), flush=True)
if "stage" in what:
    host = items.cpu().numpy()
    h0, h1 = ctx.item_headers()
    hf = 2 if metric == "dot-product" else 1
    stride = 1 + 4 * hf + 4 * d
    blob = np.zeros(n * stride, dtype=np.uint8)
    b2 = blob.reshape(n, stride)
    b2[:, 1:5] = h0.view(np.uint8).reshape(n, 4)
    b2[:, 1 + 4 * hf:] = host.view(np.uint8).reshape(n, 4 * d)
    ptrs = (blob.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(stride)).astype(np.uint64)
    for thr in (2, 4, 8, 12, 16, 24, 32):
        for mb in (4, 8, 32):
            os.environ["ARROY_B200_STAGE_THREADS"] = str(thr)
            os.environ["ARROY_B200_STAGE_CHUNK_MB"] = str(mb)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                ctx.stage_items_ptrs(metric, d, ids, ptrs)
                ts.append(time.perf_counter() - t0)
            print("STAGE threads %2d chunk %2d MB: best %.1f ms (%.1f GB/s) all %s" % (thr, mb, min(ts) * 1e3, n * d * 4 / min(ts) / 1e9, ["%.0f" % (t * 1e3) for t in ts]), flush=True)
    # pinned torch tensor through the flat path
    pinned = torch.from_numpy(host).pin_memory()
    os.environ["ARROY_B200_STAGE_THREADS"] = "12"
    os.environ["ARROY_B200_STAGE_CHUNK_MB"] = "8"
    for _ in range(3):
        t0 = time.perf_counter()
        ctx.stage_items_flat(metric, ids, pinned)
        print("STAGE flat(pinned src) %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
    t = torch.empty((n, d), dtype=torch.float32, device="cuda:0")
    torch.cuda.synchronize()
    t0 = time.perf_counter(); t.copy_(pinned); torch.cuda.synchronize()
    print("torch pinned->device copy %.1f ms (%.1f GB/s)" % ((time.perf_counter() - t0) * 1e3, n * d * 4 / (time.perf_counter() - t0) / 1e9))
